"""Oracle (CPU restatement) for variant 3D-VA pinned on the reference's own goldens for that path:
 - test/test_stokes_solvi3D.jl:25-61 : solVi3D 16^3 -> iters.norm_Rx[end] < 1e-8
 - test/test_Utils.jl:387-397        : compute_maxloc! hotspot
 - test/test_boundary_conditions3D.jl: free-slip / no-slip ghost identities
"""
import ctypes as C

import numpy as np

from justrelax_jl_b200 import setups
from util import bc_flags


def _run_solvi(oracle, n, iterMax=5000, nout=100):
    s = setups.solvi3d(n, n, n)
    d = oracle.alloc_stokes(s.ni, s.fields)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=iterMax, nout=nout)
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)  # flow_bcs!(stokes, flow_bcs) SolVi3D.jl:100
    return s, d, oracle.solve3d_VA(d, s.ni, opts)


def test_solvi3d_reference_golden(oracle):
    # test/test_stokes_solvi3D.jl: nx=ny=nz=16, tol = 1e-8 on norm_Rx[end]
    s, d, out = _run_solvi(oracle, 16)
    assert out["status"] == 0
    assert out["norm_Rx"][-1] < 1.0e-8
    assert out["norm_Ry"][-1] < 1.0e-8 and out["norm_Rz"][-1] < 1.0e-8
    # pure-shear BC of the reference has ∇·V = εbg ≠ 0 with K = Inf, so RP ≡ −1 and the loop runs to iterMax+1
    assert out["iter"] == 5001
    assert np.allclose(out["norm_divV"], 1.0 / np.sqrt(16 ** 3), rtol=1e-3)  # mean(∇V) = εbg is fixed by the BC
    # τ_o ← τ at exit (Stokes3D.jl:172-173)
    for c in ("xx", "yy", "zz", "yz", "xz", "xy"):
        assert np.array_equal(d["t" + c], d["t" + c + "_o"])


def run_burstedde(oracle, n, setup=None):
    s = (setup or setups.burstedde3d)(n)
    d = oracle.alloc_stokes(s.ni, s.fields)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)  # flow_bcs! with no active face: leaves the prescribed values  Burstedde.jl:166
    out = oracle.solve3d_VA(d, s.ni, opts)
    return s, d, out


def test_burstedde_reference_criteria(oracle):
    """test/test_stokes_burstedde.jl:28-40: the manufactured solution of Burstedde et al. at 8³ and 16³ (3D-VA, variable η over 2.8 decades,
    spatially varying body force, velocity prescribed on every face): PT converges below 1e-8, velocity errors converge with order > 1.4,
    max L2 velocity error < 3e-2 and L2 pressure error < 2e-1 at 16³"""
    errs = []
    for n in (8, 16):
        s, d, out = run_burstedde(oracle, n)
        assert out["status"] == 0 and out["err_evo1"][-1] < 1.0e-8, (n, out["err_evo1"][-1])
        errs.append(s.error_norms(d["Vx"], d["Vy"], d["Vz"], d["P"]))
    order = np.log2(np.array(errs[0]) / np.array(errs[1]))
    L2_p, L2_vx, L2_vy, L2_vz = errs[1]
    assert np.all(order[1:] > 1.4), order
    assert max(L2_vx, L2_vy, L2_vz) < 3.0e-2 and L2_p < 2.0e-1, errs[1]


def test_taylor_green_reference_criteria(oracle):
    """test/test_stokes_taylor_green.jl:29-40: the FVCA8 Taylor-Green Stokes solution at 8³ and 16³ (3D-VA, η = 1, body force in x only):
    PT converges below 1e-8, all four errors converge with order > 1.7, max L2 velocity error < 5e-3 and L2 pressure error < 1.5e-1 at 16³"""
    errs = []
    for n in (8, 16):
        s, d, out = run_burstedde(oracle, n, setups.taylor_green3d)
        assert out["status"] == 0 and out["err_evo1"][-1] < 1.0e-8, (n, out["err_evo1"][-1])
        errs.append(s.error_norms(d["Vx"], d["Vy"], d["Vz"], d["P"]))
    order = np.log2(np.array(errs[0]) / np.array(errs[1]))
    L2_p, L2_vx, L2_vy, L2_vz = errs[1]
    assert np.all(order > 1.7), order
    assert max(L2_vx, L2_vy, L2_vz) < 5.0e-3 and L2_p < 1.5e-1, errs[1]


def test_norm_mpi_kats(oracle):
    """test/test_Utils.jl:172,237: norm_mpi(η) === 4.0 for η = ones(4, 4) and === 8.0 for ones(4, 4, 4) — the single-rank value of the
    residual-norm reduction (Utils.jl:698-701: √(Σ A²))"""
    assert np.sqrt(oracle.sumsq(np.ones((4, 4), order="F"), False)) == 4.0
    assert np.sqrt(oracle.sumsq(np.ones((4, 4, 4), order="F"), False)) == 8.0
    # the solvers' view A[2:end-1, …] (interior = true)
    assert oracle.sumsq(np.ones((6, 6, 6), order="F"), True) == 64.0


def test_maxloc_hotspot(oracle):
    # test_Utils.jl:387-397 analogue in 3D: a single hotspot spreads to its 3x3x3 neighbourhood
    A = np.ones((5, 5, 5), order="F")
    A[2, 2, 2] = 5.0
    B = np.zeros_like(A, order="F")
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    oracle.lib().orc_maxloc3(dp(B), dp(A), 5, 5, 5, 1, 1, 1)
    assert np.all(B[1:4, 1:4, 1:4] == 5.0)
    B[1:4, 1:4, 1:4] = 1.0
    assert np.all(B == 1.0)


def test_free_slip_and_no_slip_identities(oracle):
    # test/test_boundary_conditions3D.jl: free slip ghosts equal the adjacent interior layer, no-slip ghosts the
    # negated layer and normal components vanish.
    rng = np.random.default_rng(1)
    n = (6, 5, 4)
    nx, ny, nz = n
    mk = lambda: [np.asfortranarray(rng.uniform(size=s)) for s in ((nx + 1, ny + 2, nz + 2), (nx + 2, ny + 1, nz + 2), (nx + 2, ny + 2, nz + 1))]
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    i32 = lambda v: (C.c_int32 * len(v))(*v)
    Ax, Ay, Az = mk()
    oracle.lib().orc_free_slip3(dp(Ax), dp(Ay), dp(Az), i32(n), i32([1] * 6))
    assert np.array_equal(Ax[:, 0, :], Ax[:, 1, :]) and np.array_equal(Ax[:, -1, :], Ax[:, -2, :])
    assert np.array_equal(Ax[:, :, 0], Ax[:, :, 1]) and np.array_equal(Ax[:, :, -1], Ax[:, :, -2])
    assert np.array_equal(Ay[0, :, :], Ay[1, :, :]) and np.array_equal(Ay[-1, :, :], Ay[-2, :, :])
    assert np.array_equal(Ay[:, :, 0], Ay[:, :, 1]) and np.array_equal(Ay[:, :, -1], Ay[:, :, -2])
    assert np.array_equal(Az[0, :, :], Az[1, :, :]) and np.array_equal(Az[-1, :, :], Az[-2, :, :])
    assert np.array_equal(Az[:, 0, :], Az[:, 1, :]) and np.array_equal(Az[:, -1, :], Az[:, -2, :])
    # applying it twice changes nothing (the reference tests apply twice, test_boundary_conditions3D.jl:85-86)
    Bx, By, Bz = Ax.copy(order="F"), Ay.copy(order="F"), Az.copy(order="F")
    oracle.lib().orc_free_slip3(dp(Bx), dp(By), dp(Bz), i32(n), i32([1] * 6))
    assert np.array_equal(Ax, Bx) and np.array_equal(Ay, By) and np.array_equal(Az, Bz)
    Ax, Ay, Az = mk()
    oracle.lib().orc_no_slip3(dp(Ax), dp(Ay), dp(Az), i32(n), i32([1] * 6))
    assert np.all(Ax[0] == 0) and np.all(Ax[-1] == 0) and np.all(Ay[:, 0] == 0) and np.all(Ay[:, -1] == 0)
    assert np.all(Az[:, :, 0] == 0) and np.all(Az[:, :, -1] == 0)
    assert np.array_equal(Ay[0, 1:-1, 1:-1], -Ay[1, 1:-1, 1:-1]) and np.array_equal(Az[-1, 1:-1, 1:-1], -Az[-2, 1:-1, 1:-1])


def test_iteration_matches_numpy_restatement(oracle):
    """Independent cross-check of the C oracle: one PT iteration written directly with numpy slices
    from the formulas of SURVEY.md Appendix A (no fma: agreement to 1e-13)."""
    s = setups.random_stokes3d((7, 6, 5), seed=3)
    d = oracle.alloc_stokes(s.ni, s.fields)
    ref = {k: v.copy(order="F") for k, v in d.items()}
    pt = s.pt_stokes
    flags = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)
    opts = oracle.make_opts(pt, s.grid._di.center, s.dt, flags, s.ni, iterMax=1, nout=1)
    oracle.iterate3d_VA(d, s.ni, opts, 1)

    _dx, _dy, _dz = s.grid._di.center
    dt, r, th, edt = s.dt, pt.r, pt.θ_dτ, pt.ηdτ
    Vx, Vy, Vz = ref["Vx"], ref["Vy"], ref["Vz"]
    eta, G, K = ref["eta"], ref["G"], ref["K"]
    divV = (Vx[1:, 1:-1, 1:-1] - Vx[:-1, 1:-1, 1:-1]) * _dx + (Vy[1:-1, 1:, 1:-1] - Vy[1:-1, :-1, 1:-1]) * _dy + \
           (Vz[1:-1, 1:-1, 1:] - Vz[1:-1, 1:-1, :-1]) * _dz
    RP = -(ref["P"] - ref["P0"]) / (K * dt) - divV + ref["Q"] / dt
    psi = 1.0 / (1.0 / eta + 1.0 / (G * dt)) * r / th
    P = ((ref["P0"] / (K * dt) - divV + ref["Q"] / dt) * psi + ref["P"]) / (1 + psi / (K * dt))
    exx = (Vx[1:, 1:-1, 1:-1] - Vx[:-1, 1:-1, 1:-1]) * _dx - divV / 3
    eyz = 0.5 * (_dz * (Vy[1:-1, :, 1:] - Vy[1:-1, :, :-1]) + _dy * (Vz[1:-1, 1:, :] - Vz[1:-1, :-1, :]))
    pad = lambda A, ax: np.pad(A, [(1, 1) if a in ax else (0, 0) for a in range(3)], mode="edge")
    def av(A, ax):
        Ap = pad(A, ax)
        sl = lambda o0, o1: tuple(slice(o, o + A.shape[a] + 1) if a in ax else slice(None) for a, o in zip(range(3), (o0 if ax[0] == 0 else (o0 if ax[0] == 1 and False else 0), 0, 0)))
        # explicit 4-point average over the two averaged axes
        a0, a1 = ax
        idx = [slice(None)] * 3
        out = 0
        for o0 in (0, 1):
            for o1 in (0, 1):
                idx[a0] = slice(o0, o0 + A.shape[a0] + 1)
                idx[a1] = slice(o1, o1 + A.shape[a1] + 1)
                out = out + Ap[tuple(idx)]
        return 0.25 * out
    def upd(t, to, e, et, g):
        dtr = 1.0 / (th + et / (g * dt) + 1.0)
        return t + dtr * (2 * et * e - (t - to) * et / (g * dt) - t)
    txx = upd(ref["txx"], ref["txx_o"], exx, eta, G)
    tyz = upd(ref["tyz"], ref["tyz_o"], eyz, av(eta, (1, 2)), av(G, (1, 2)))
    assert np.allclose(d["divV"], divV, rtol=0, atol=1e-13 * np.abs(divV).max())
    assert np.allclose(d["RP"], RP, rtol=0, atol=1e-13 * np.abs(RP).max())
    assert np.allclose(d["P"], P, rtol=0, atol=1e-13 * np.abs(P).max())
    assert np.allclose(d["exx"], exx, rtol=0, atol=1e-13 * np.abs(exx).max())
    assert np.allclose(d["eyz"], eyz, rtol=0, atol=1e-13 * np.abs(eyz).max())
    assert np.allclose(d["txx"], txx, rtol=0, atol=1e-13 * np.abs(txx).max())
    assert np.allclose(d["tyz"], tyz, rtol=0, atol=1e-13 * np.abs(tyz).max())
    # x-momentum
    txy, txz = d["txy"], d["txz"]
    Rx = (d["txx"][1:] - d["txx"][:-1]) * _dx + _dy * (txy[1:-1, 1:, :] - txy[1:-1, :-1, :]) + _dz * (txz[1:-1, :, 1:] - txz[1:-1, :, :-1]) \
        - (d["P"][1:] - d["P"][:-1]) * _dx - 0.5 * (ref["rhogx"][1:] + ref["rhogx"][:-1])
    assert np.allclose(d["Rx"], Rx, rtol=0, atol=1e-12 * np.abs(Rx).max())
    ett = d["etatau"]
    Vx_new = ref["Vx"][1:-1, 1:-1, 1:-1] + Rx * edt / (0.5 * (ett[1:] + ett[:-1]))
    assert np.allclose(d["Vx"][1:-1, 1:-1, 1:-1], Vx_new, rtol=0, atol=1e-12 * np.abs(Vx_new).max())
    # U = V·dt is taken BEFORE flow_bcs! refreshes the ghosts (Stokes3D.jl:118-119)
    assert np.array_equal(d["Ux"][:, 1:-1, 1:-1], d["Vx"][:, 1:-1, 1:-1] * dt)
