"""Oracle pins for the per-time-step kernels around the loops, against the reference's own known-answer tests:
test/test_Utils.jl:498-561 (compute_lithostatic_pressure!), test/test_Interpolations.jl:150-192 (velocity2vertex!/velocity2center!),
test/test_shearheating2D.jl:246 (H_s ≥ 0)."""
import ctypes as C

import numpy as np

from justrelax_jl_b200 import rheology as R


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i3(v):
    return (C.c_int32 * 3)(*[int(x) for x in v])


def litho(oracle, P, rg, dz, above=None):
    nd = P.ndim
    n = list(P.shape) + [1] * (3 - nd)
    dzv = None if np.isscalar(dz) else np.ascontiguousarray(dz, dtype=np.float64)
    oracle.lib().orc_lithostatic_pressure(nd, _i3(n), _dp(P), _dp(rg), C.c_double(0.0 if dzv is not None else dz), _dp(dzv) if dzv is not None else None,
                                          _dp(above) if above is not None else None)
    return P


def test_lithostatic_pressure_reference_kats(oracle):
    nx, ny, dz = 3, 4, 0.5
    rg = np.full((nx, ny), 2.0, order="F")
    P = litho(oracle, np.zeros((nx, ny), order="F"), rg, dz)
    assert np.allclose(P[0, :], [(ny - j + 0.5) * 2.0 * dz for j in range(1, ny + 1)])          # test_Utils.jl:507
    assert np.all(P == P[0:1, :])                                                               # :508
    assert np.isclose(P[0, -1], 2.0 * dz / 2)                                                   # :511
    rg = np.asfortranarray([[float(i + 2 * j) for j in range(1, ny + 1)] for i in range(1, nx + 1)])
    P = litho(oracle, np.zeros((nx, ny), order="F"), rg, dz)
    for i in range(nx):
        for j in range(ny):
            assert np.isclose(P[i, j], rg[i, j + 1:].sum() * dz + rg[i, j] * dz / 2)          # :517-519
    dzs = np.array([0.25, 0.5, 1.0, 2.0])
    Pv = litho(oracle, np.zeros((nx, ny), order="F"), rg, dzs)
    for i in range(nx):
        for j in range(ny):
            assert np.isclose(Pv[i, j], (rg[i, j + 1:] * dzs[j + 1:]).sum() + rg[i, j] * dzs[j] / 2)   # :523-530
    assert np.allclose(litho(oracle, np.zeros((nx, ny), order="F"), rg, np.full(ny, dz)), P)           # :533-535
    rg3 = np.asfortranarray([[[float(i + j + k) for k in range(1, 4)] for j in range(1, 3)] for i in range(1, 3)])
    P3 = litho(oracle, np.zeros_like(rg3, order="F"), rg3, dz)
    for i in range(2):
        for j in range(2):
            for k in range(3):
                assert np.isclose(P3[i, j, k], rg3[i, j, k + 1:].sum() * dz + rg3[i, j, k] * dz / 2)   # :553-559


def vel2(oracle, fn, ni, e, V):
    nd = len(ni)
    out = [np.zeros(e, order="F") for _ in range(nd)]
    z = lambda seq: _dp(seq[2]) if nd == 3 else None
    getattr(oracle.lib(), fn)(nd, _i3(list(ni) + [1] * (3 - nd)), _i3(list(e) + [1] * (3 - nd)), _dp(out[0]), _dp(out[1]), z(out), _dp(V[0]), _dp(V[1]), z(V))
    return out


def test_velocity_interpolation_reference_kats(oracle):
    rng = np.random.default_rng(0)
    n = (3, 3, 3)
    Vx, Vy, Vz = (np.asfortranarray(rng.random(s)) for s in ((4, 5, 5), (5, 4, 5), (5, 5, 4)))
    Xv, Yv, Zv = vel2(oracle, "orc_velocity2vertex", n, n, (Vx, Vy, Vz))                          # test_Interpolations.jl:150-163
    Xc, Yc, Zc = vel2(oracle, "orc_velocity2center", n, n, (Vx, Vy, Vz))                          # :180-192
    for k in range(3):
        for j in range(3):
            for i in range(3):
                assert np.isclose(Xv[i, j, k], 0.25 * (Vx[i, j, k] + Vx[i, j + 1, k] + Vx[i, j, k + 1] + Vx[i, j + 1, k + 1]))
                assert np.isclose(Yv[i, j, k], 0.25 * (Vy[i, j, k] + Vy[i + 1, j, k] + Vy[i, j, k + 1] + Vy[i + 1, j, k + 1]))
                assert np.isclose(Zv[i, j, k], 0.25 * (Vz[i, j, k] + Vz[i, j + 1, k] + Vz[i + 1, j, k] + Vz[i + 1, j + 1, k]))
                assert np.isclose(Xc[i, j, k], (Vx[i, j + 1, k + 1] + Vx[i + 1, j + 1, k + 1]) / 2)
                assert np.isclose(Yc[i, j, k], (Vy[i + 1, j, k + 1] + Vy[i + 1, j + 1, k + 1]) / 2)
                assert np.isclose(Zc[i, j, k], (Vz[i + 1, j + 1, k] + Vz[i + 1, j + 1, k + 1]) / 2)
    n2 = (4, 3)
    Vx2, Vy2 = np.asfortranarray(rng.random((5, 5))), np.asfortranarray(rng.random((6, 4)))
    Xv2, Yv2 = vel2(oracle, "orc_velocity2vertex", n2, (5, 4), (Vx2, Vy2))                        # Interpolations.jl:244-248
    assert np.allclose(Xv2, (Vx2[:, :-1] + Vx2[:, 1:]) / 2) and np.allclose(Yv2, (Vy2[:-1, :] + Vy2[1:, :]) / 2)
    Xc2, Yc2 = vel2(oracle, "orc_velocity2center", n2, n2, (Vx2, Vy2))                            # :285-289
    assert np.allclose(Xc2, (Vx2[:-1, 1:-1] + Vx2[1:, 1:-1]) / 2) and np.allclose(Yc2, (Vy2[1:-1, :-1] + Vy2[1:-1, 1:]) / 2)


def shear_heating_case(oracle, ni, seed, nphase):
    """random τ, τ_o, ε; returns (host slot dict, vc, chi, dt, oracle output)"""
    rng = np.random.default_rng(seed)
    nd = len(ni)
    d = oracle.alloc_stokes(ni)
    for nm in list(d):
        if nm[0] in "te" and nm[1:3] in ("xx", "yy", "zz", "yz", "xz", "xy") and d[nm] is not None:
            d[nm][...] = rng.uniform(-1, 1, size=d[nm].shape)
    nc = int(np.prod(ni))
    ratios = None
    if nphase > 1:
        r = rng.dirichlet(np.ones(nphase), size=ni)
        r[rng.uniform(size=ni) < 0.3] = np.eye(nphase)[0]
        ratios = dict(center=np.asfortranarray(r))
    mats = tuple(R.SetMaterialParams(Phase=p + 1, Density=R.ConstantDensity(ρ=1.0), ShearHeat=R.ConstantShearheating(Χ=0.5 + 0.25 * p),
                                     CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0), R.ConstantElasticity(G=1.0 + p, ν=0.4))))
                 for p in range(max(nphase, 1)))
    vc = oracle.vc_inputs(R.lower_stokes(mats), (0.0, 0.0, 0.0), ratios or {})
    chi = np.array(R.shear_heating_coefficients(mats))
    dt = 0.3
    out = np.zeros(ni, order="F")
    fs = oracle.make_fields(d, ni)
    oracle.lib().orc_shear_heating(C.byref(fs), C.byref(vc), _dp(chi), C.c_double(dt), _dp(out))
    return d, mats, ratios, dt, out


def test_shear_heating_definition_and_sign(oracle):
    d, mats, ratios, dt, out = shear_heating_case(oracle, (7, 6), 3, 1)
    assert (out >= 0).all() and out.max() > 0                                                     # test_shearheating2D.jl:246
    exy_c = 0.25 * (d["exy"][:-1, :-1] + d["exy"][1:, :-1] + d["exy"][:-1, 1:] + d["exy"][1:, 1:])
    G, X = 1.0, 0.5
    el = lambda t, to: 0.5 * (t - to) / (G * dt)
    H = (d["txx"] * (d["exx"] - el(d["txx"], d["txx_o"])) + d["tyy"] * (d["eyy"] - el(d["tyy"], d["tyy_o"])) +
         2 * d["txy_c"] * (exy_c - el(d["txy_c"], d["txy_o_c"])))
    assert np.allclose(out, np.maximum(0.0, X * H), rtol=1e-12, atol=1e-14)
    d3, _, _, _, out3 = shear_heating_case(oracle, (5, 4, 6), 4, 3)
    assert (out3 >= 0).all() and out3.max() > 0
