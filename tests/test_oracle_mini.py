"""Pin the oracle's mini-kernels bit-for-bit against the reference's known answers:
/root/reference/test/test_mini_kernels.jl:4-117 (values copied from there, not computed)."""
import numpy as np
import pytest

A2 = np.arange(1.0, 17.0).reshape((4, 4), order="F")
A3 = np.arange(1.0, 65.0).reshape((4, 4, 4), order="F")
i = j = k = 2
dx, dy, dz = 2.0, 3.0, 4.0


def harmonic(*x):
    return len(x) / sum(1.0 / v for v in x)


def test_accessors(oracle):
    m2 = lambda n: oracle.mini2(n, A2, 0.0, i, j)
    m3 = lambda n: oracle.mini3(n, A3, 0.0, i, j, k)
    assert m2("center") == 6.0 and m2("next") == 11.0 and m2("left") == 5.0 and m2("right") == 7.0
    assert m2("back") == 2.0 and m2("front") == 10.0
    assert m3("left") == 21.0 and m3("right") == 23.0 and m3("back") == 18.0 and m3("front") == 26.0
    assert m3("bot") == 6.0 and m3("top") == 38.0


def test_differences(oracle):
    assert oracle.mini2("_d_xa", A2, dx, i, j) == 2.0
    assert oracle.mini2("_d_ya", A2, dy, i, j) == 12.0
    assert oracle.mini3("_d_za", A3, dz, i, j, k) == 64.0
    assert oracle.mini2("_d_xi", A2, dx, i, j) == 2.0
    assert oracle.mini2("_d_yi", A2, dy, i, j) == 12.0
    assert oracle.mini3("_d_xi", A3, dx, i, j, k) == 2.0
    assert oracle.mini3("_d_yi", A3, dy, i, j, k) == 12.0
    assert oracle.mini3("_d_zi", A3, dz, i, j, k) == 64.0
    # div(Ax, Ay, dx, dy, i, j) == 26.0 ; div(Ax3, 2Ax3, 3Ax3, ...) == 218.0
    assert oracle.mini2("_d_xi", A2, dx, i, j) + oracle.mini2("_d_yi", 2.0 * A2, dy, i, j) == 26.0
    assert (oracle.mini3("_d_xi", A3, dx, i, j, k) + oracle.mini3("_d_yi", 2.0 * A3, dy, i, j, k)
            + oracle.mini3("_d_zi", 3.0 * A3, dz, i, j, k)) == 218.0


def test_averages(oracle):
    m2 = lambda n: oracle.mini2(n, A2, 0.0, i, j)
    m3 = lambda n, ii=i, jj=j, kk=k: oracle.mini3(n, A3, 0.0, ii, jj, kk)
    assert m2("_av") == 13.5 and m2("_av_a") == 8.5 and m2("_av_xa") == 6.5 and m2("_av_ya") == 8.0
    assert m2("_av_xi") == 10.5 and m2("_av_yi") == 9.0
    assert m2("_harm") == pytest.approx(harmonic(11.0, 12.0, 15.0, 16.0), rel=1e-15)
    assert m2("_harm_a") == pytest.approx(harmonic(6.0, 7.0, 10.0, 11.0), rel=1e-15)
    assert m2("_harm_xa") == pytest.approx(harmonic(6.0, 7.0), rel=1e-15)
    assert m2("_harm_ya") == pytest.approx(harmonic(6.0, 10.0), rel=1e-15)
    assert m3("_av") == 32.5 and m3("_av_x") == 22.5 and m3("_av_y") == 24.0 and m3("_av_z") == 30.0
    assert m3("_av_xy") == 24.5 and m3("_av_xz") == 30.5 and m3("_av_yz") == 32.0
    assert m3("_av_xyi") == 19.5 and m3("_av_xzi") == 13.5 and m3("_av_yzi") == 12.0
    assert m3("_harm_x") == pytest.approx(harmonic(22.0, 23.0), rel=1e-15)
    assert m3("_harm_y") == pytest.approx(harmonic(22.0, 26.0), rel=1e-15)
    assert m3("_harm_z") == pytest.approx(harmonic(22.0, 38.0), rel=1e-15)
    assert m3("_harm_xy") == pytest.approx(harmonic(22.0, 23.0, 26.0, 27.0), rel=1e-15)
    assert m3("_harm_xz") == pytest.approx(harmonic(22.0, 23.0, 38.0, 39.0), rel=1e-15)
    assert m3("_harm_yz") == pytest.approx(harmonic(22.0, 26.0, 38.0, 42.0), rel=1e-15)
    assert m3("_harm_xyi") == pytest.approx(harmonic(17.0, 18.0, 21.0, 22.0), rel=1e-15)
    assert m3("_harm_xzi") == pytest.approx(harmonic(5.0, 6.0, 21.0, 22.0), rel=1e-15)
    assert m3("_harm_yzi") == pytest.approx(harmonic(2.0, 6.0, 18.0, 22.0), rel=1e-15)


def test_clamped(oracle):
    m2 = lambda n, ii, jj: oracle.mini2(n, A2, 0.0, ii, jj)
    m3 = lambda n, ii=i, jj=j, kk=k: oracle.mini3(n, A3, 0.0, ii, jj, kk)
    assert m2("_av_ai_clamped", i, j) == m2("_av_a", i - 1, j - 1)
    assert m2("_av_ai_clamped", 1, 1) == A2[0, 0]
    assert m3("_av_xyi_clamped") == m3("_av_xyi") and m3("_av_xzi_clamped") == m3("_av_xzi")
    assert m3("_av_yzi_clamped") == m3("_av_yzi")
    assert m3("_harm_xyi_clamped") == pytest.approx(m3("_harm_xyi"), rel=1e-15)
    assert m3("_harm_xzi_clamped") == pytest.approx(m3("_harm_xzi"), rel=1e-15)
    assert m3("_harm_yzi_clamped") == pytest.approx(m3("_harm_yzi"), rel=1e-15)
    assert m3("_av_xyi_clamped", 1, j, k) == 0.5 * (A3[0, 0, k - 1] + A3[0, 1, k - 1])
    assert m3("_harm_xyi_clamped", 1, j, k) == pytest.approx(harmonic(A3[0, 0, k - 1], A3[0, 1, k - 1]), rel=1e-15)
    assert m3("_av_xzi_clamped", i, j, 1) == 0.5 * (A3[0, j - 1, 0] + A3[1, j - 1, 0])
    assert m3("_harm_yzi_clamped", i, 1, 1) == pytest.approx(A3[i - 1, 0, 0], rel=1e-15)


def test_mysum(oracle):
    v = np.arange(1.0, 6.0)
    assert oracle.mini2("mysum1", v, 0.0, 2, 4) == 9.0
    assert oracle.mini2("mysum1", v, 1.0, 2, 4) == 1.0833333333333333
    assert oracle.mini2("mysum", A2, 0.0, 2, 2) == 34.0
    assert oracle.mini2("mysum", A2, 1.0, 2, 2) == 0.5004329004329005
    assert oracle.mini3("mysum", A3, 0.0, 2, 2, 2) == 260.0
    assert oracle.mini3("mysum", A3, 1.0, 2, 2, 2) == 0.2634535347004082
