"""CPU oracle of update_phase_ratios_{2,3}D! (src/phases/PhaseRatios.jl) against the reference's own known answers (no GPU):
test/test_phase_ratios3D.jl:31-84 and test/test_rheology.jl:444-500."""
import numpy as np


def _grid(nd, nx=4):
    xv = [np.linspace(0.0, 1.0, nx + 1)] * nd           # range(0.0, 1.0; length = nx + 1)
    xc = [np.linspace(0.125, 0.875, nx)] * nd           # range(0.125, 0.875; length = nx)
    return xc, xv


def test_phase_ratios3d_reference_kats(oracle):
    nx = 4
    xc, xv = _grid(3)
    p1, p2 = np.zeros((nx,) * 3, order="F"), np.zeros((nx,) * 3, order="F")
    p1[:2], p2[2:] = 1.0, 1.0
    o = oracle.phase_ratios_from_arrays((p1, p2), xc, xv)
    assert o["center"].shape[:3] == (4, 4, 4) and o["vertex"].shape[:3] == (5, 5, 5)                       # :57-61
    assert [o[k].shape[:3] for k in ("Vx", "Vy", "Vz")] == [(5, 4, 4), (4, 5, 4), (4, 4, 5)]
    assert [o[k].shape[:3] for k in ("xy", "yz", "xz")] == [(5, 5, 4), (4, 5, 5), (5, 4, 5)]
    assert np.allclose(o["center"][:, 0, 0, 0], [1, 1, 0, 0]) and np.allclose(o["center"][:, 0, 0, 1], [0, 0, 1, 1])   # :64-65
    for A in o.values():                                                                                 # :68-70
        assert np.allclose(A.sum(-1), 1.0)
    v = o["vertex"][2, 1, 1]                                                                             # :73-74
    assert v[0] > 0 and v[1] > 0
    o3 = oracle.phase_ratios_from_arrays((np.full((nx,) * 3, 0.6), np.full((nx,) * 3, 0.4), np.full((nx,) * 3, 1.0e-6)), xc, xv)   # :77-83
    assert o3["center"][1, 1, 1, 2] == 0.0 and abs(o3["center"][1, 1, 1, 0] + o3["center"][1, 1, 1, 1] - 1.0) < 1e-15


def test_phase_ratios2d_reference_kats(oracle):
    nx = 4
    xc, xv = _grid(2)
    p1, p2 = np.zeros((nx, nx), order="F"), np.zeros((nx, nx), order="F")
    p1[:2], p2[2:] = 1.0, 1.0
    o = oracle.phase_ratios_from_arrays((p1, p2), xc, xv)                                                  # test_rheology.jl:463
    assert np.allclose(o["center"][:, 0, 0], [1, 1, 0, 0]) and np.allclose(o["center"][:, 0, 1], [0, 0, 1, 1])   # :473-476
    for k in ("center", "vertex", "Vx", "Vy"):                                                            # :479-491
        assert np.allclose(o[k].sum(-1), 1.0)
    v = o["vertex"][2, 1]                                                                                # :485-486
    assert v[0] > 0 and v[1] > 0
    o3 = oracle.phase_ratios_from_arrays((np.full((nx, nx), 0.6), np.full((nx, nx), 0.4), np.full((nx, nx), 1.0e-6)), xc, xv)
    assert o3["center"][1, 1, 2] == 0.0 and abs(o3["center"][1, 1, 0] + o3["center"][1, 1, 1] - 1.0) < 1e-15   # :498-499


def test_phase_ratios_weights_independent_check(oracle):
    """vertex ratios of a random one-hot field = trilinear (equal, on a uniform grid) average of the ≤ 8 surrounding cells; faces = 2-cell,
    midpoints = 4-cell averages restricted to existing cells"""
    rng = np.random.default_rng(1)
    ni = (5, 4, 3)
    lab = rng.integers(0, 3, size=ni)
    ph = [np.asfortranarray((lab == p).astype(float)) for p in range(3)]
    xv = [np.linspace(0, 1, n + 1) for n in ni]
    xc = [0.5 * (x[1:] + x[:-1]) for x in xv]
    o = oracle.phase_ratios_from_arrays(ph, xc, xv)

    def avg(axes):
        P = np.stack(ph, -1)
        cnt = np.ones(ni + (1,))
        for ax in axes:
            pad = [(0, 0)] * 4
            pad[ax] = (1, 1)
            P = np.pad(P, pad)
            cnt = np.pad(cnt, pad)
            sl0, sl1 = [slice(None)] * 4, [slice(None)] * 4
            sl0[ax], sl1[ax] = slice(0, -1), slice(1, None)
            P, cnt = P[tuple(sl0)] + P[tuple(sl1)], cnt[tuple(sl0)] + cnt[tuple(sl1)]
        r = P / cnt
        r[r < 1e-5] = 0
        return r / r.sum(-1, keepdims=True)

    for k, axes in dict(vertex=(0, 1, 2), Vx=(0,), Vy=(1,), Vz=(2,), xy=(0, 1), yz=(1, 2), xz=(0, 2)).items():
        assert np.allclose(o[k], avg(axes), rtol=0, atol=1e-14), k
    assert np.array_equal(o["center"], np.stack(ph, -1))
