"""CPU oracle of the 2D Stokes PT loops pinned on the reference's goldens (no GPU).

 - test/test_stokes_solcx.jl:26-43   (config 2 at 32²): 2D-V2 converges, err_evo1[end] < 1e-8
 - test/test_shearband2D.jl:194-202  (config 3 at 32²): 2D-VC + Drucker-Prager, 10 steps: err < 1e-6,
   extrema(τII) ≈ (1.5128689768248313, 1.6415759440014273) atol 1e-3, maximum(τxx) ≈ 1.6376258215356436 atol 1e-4
"""
import ctypes as C
import math

import numpy as np

from justrelax_jl_b200 import rheology as R, setups
from util import bc_flags


def run_shearband(oracle, s, *, strain_increment=0, displacement_bcs=0):
    d = oracle.alloc_stokes(s.ni, s.fields)
    rows = R.lower_stokes(s.rheology)
    vc = oracle.vc_inputs(rows, R.gravity_of(s.rheology), s.ratios)
    kw = s.kwargs
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=kw["iterMax"], nout=kw["nout"],
                            viscosity_cutoff=kw["viscosity_cutoff"], strain_increment=strain_increment, displacement_bcs=displacement_bcs)
    if displacement_bcs:   # the displacement form prescribes U = V·dt (pureshear_bc! on the displacement, then flow_bcs! on U)
        d["Ux"][...] = d["Vx"] * s.dt
        d["Uy"][...] = d["Vy"] * s.dt
    fs = oracle.make_fields(d, s.ni)
    # compute_viscosity!(stokes, phase_ratios, args, rheology, (-Inf, Inf)) with relaxation 1   test_shearband2D.jl:133-135
    oracle.lib().orc_viscosity2d(C.byref(fs), C.byref(opts), C.byref(vc), C.c_double(1.0))
    oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts), displacement_bcs)
    outs, txx_max = [], []
    for _ in range(s.nt):
        outs.append(oracle.solve2d_VC(d, s.ni, opts, vc))
        txx_max.append(d["txx"].max())
    return d, outs, txx_max


def test_shearband2d_reference_golden(oracle):
    s = setups.shearband2d(32)
    d, outs, txx_max = run_shearband(oracle, s)
    assert all(o["status"] == 0 for o in outs)
    assert outs[-1]["err_evo1"][-1] < 1.0e-6
    tII = oracle.tensor_invariant2d(d["txx"], d["tyy"], d["txy"])          # tensor_invariant!(stokes.τ)
    assert abs(tII.min() - 1.5128689768248313) < 1.0e-3
    assert abs(tII.max() - 1.6415759440014273) < 1.0e-3
    assert abs(txx_max[-1] - 1.6376258215356436) < 1.0e-4
    # plasticity was active
    assert d["EII_pl"].max() > 0 and d["lam"].max() > 0


def test_shearband2d_strain_increment_form_reproduces_golden(oracle):
    """kwarg strain_increment = true (Δε form: Stokes2D.jl:659-730, StressKernels.jl:1147-1302), with velocity and with displacement
    boundary conditions (types/displacement.jl:62-70): the same physics, so the reference's shear-band golden
    (test/test_shearband2D.jl:197-201) must be reproduced to its own tolerances"""
    s = setups.shearband2d(32)
    for dbc in (0, 1):
        d, outs, txx_max = run_shearband(oracle, s, strain_increment=1, displacement_bcs=dbc)
        assert all(o["status"] == 0 for o in outs) and outs[-1]["err_evo1"][-1] < 1.0e-6
        tII = oracle.tensor_invariant2d(d["txx"], d["tyy"], d["txy"])
        assert abs(tII.min() - 1.5128689768248313) < 1.0e-3, (dbc, tII.min())
        assert abs(tII.max() - 1.6415759440014273) < 1.0e-3, (dbc, tII.max())
        assert abs(txx_max[-1] - 1.6376258215356436) < 1.0e-4, (dbc, txx_max[-1])
        assert d["EII_pl"].max() > 0
        # Δε / dt is the strain rate
        assert np.allclose(d["dxx"] / s.dt, d["exx"], rtol=1e-12, atol=0) and np.abs(d["dxx"]).max() > 0


def run_sinking_block(oracle, s):
    d = oracle.alloc_stokes(s.ni, s.fields)
    rows = R.lower_stokes(s.rheology)
    vc = oracle.vc_inputs(rows, R.gravity_of(s.rheology), s.ratios)
    kw = s.kwargs
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=kw["iterMax"], nout=kw["nout"],
                            viscosity_cutoff=kw["viscosity_cutoff"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_viscosity2d(C.byref(fs), C.byref(opts), C.byref(vc), C.c_double(1.0))   # compute_viscosity!  test_sinking_block.jl:158
    oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts), 0)
    out = oracle.solve2d_VC(d, s.ni, opts, vc)
    return d, out


def vertex_speed(Vx, Vy):
    """velocity2vertex! (2D) + √(Vx_v² + Vy_v²)  test_sinking_block.jl:192-195"""
    Vx_v = 0.5 * (Vx[:, :-1] + Vx[:, 1:])
    Vy_v = 0.5 * (Vy[:-1, :] + Vy[1:, :])
    return np.sqrt(Vx_v ** 2 + Vy_v ** 2)


def test_sinking_block_reference_golden(oracle):
    """test/test_sinking_block.jl:202-208: 2D-VC with buoyancy in SI units, no elasticity (G = Kb = Inf): converges below 1e-5 and
    maximum(velocity) ≈ 4.841885609356093e-10 (atol 1e-6 in the reference; the restatement lands within a few per cent)"""
    s = setups.sinking_block2d(32)
    d, out = run_sinking_block(oracle, s)
    assert out["status"] == 0 and out["err_evo1"][-1] < 1.0e-5 and out["iter"] < s.kwargs["iterMax"]
    vmax = vertex_speed(d["Vx"], d["Vy"]).max()
    assert abs(vmax - 4.841885609356093e-10) < 1.0e-6
    assert abs(vmax / 4.841885609356093e-10 - 1) < 0.1, vmax
    # the block sinks: Vy < 0 at the block centre (x = 250 km, depth = 100 km)
    assert d["Vy"][17, 26] < 0


def run_elastic_buildup(oracle, s, solve=None):
    """the time loop of Elastic_BuildUp.jl:74-103 on the oracle; returns (fields, max|τyy| per step, analytic values, iterations per step)"""
    d = oracle.alloc_stokes(s.ni, s.fields)
    opts0 = oracle.make_opts(s.pt_stokes, s.grid._di.center, 1.0, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts0), 0)
    t, av, sol, iters = 0.0, [], [], []
    while t < s.ttot:
        dt = s.dt_of(t)
        opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
        out = oracle.solve2d_V2(d, s.ni, opts)
        assert out["status"] == 0
        t += dt
        av.append(np.abs(d["tyy"]).max()); sol.append(s.solution(t)); iters.append(out["iter"])
    return d, np.array(av), np.array(sol), iters


def test_elastic_buildup_reference_criterion(oracle):
    """test/test_stokes_elastic_buildup.jl:24-53: visco-elastic stress build-up under pure shear (2D-V2 with finite G and dt): the mean
    relative error of max|τyy| against 2 εbg η0 (1 − exp(−G t/η0)) over the 200 steps is ≤ 5e-3"""
    s = setups.elastic_buildup2d(32)
    d, av, sol, iters = run_elastic_buildup(oracle, s)
    err = np.mean(np.abs(np.abs(av) - sol) / sol)
    assert len(av) == 200 and err <= 5.0e-3, err


def test_continuation_linear_kat(oracle):
    """test/test_Utils.jl:150: continuation_linear(1.0, 0.8, 0.05) === 0.81 — the viscosity relaxation of update_viscosity_τII!
    (Viscosity.jl:382-418: η ← clamp((1 − ν)·η + ν·η_GP)) with η = 0.8, a single LinearViscous phase of η = 1, ν = 0.05"""
    ni = (4, 4)
    rheo = (R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=1.0), CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0),))),)
    ratios = dict(center=np.ones(ni + (1,), order="F"), vertex=np.ones((5, 5, 1), order="F"))
    d = oracle.alloc_stokes(ni, dict(eta=np.full(ni, 0.8, order="F")))
    vc = oracle.vc_inputs(R.lower_stokes(rheo), R.gravity_of(rheo), ratios)
    pt = setups.PTStokesCoeffs((1.0, 1.0), (0.25, 0.25))
    opts = oracle.make_opts(pt, (4.0, 4.0), 1.0, dict(free_slip=[1] * 6), ni, iterMax=1, nout=1)
    fs = oracle.make_fields(d, ni)
    oracle.lib().orc_viscosity2d(C.byref(fs), C.byref(opts), C.byref(vc), C.c_double(0.05))
    assert np.all(d["eta"] == 0.81)


def test_solcx_reference_golden(oracle):
    s = setups.solcx2d(32, 32)
    d = oracle.alloc_stokes(s.ni, s.fields)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts), 0)
    out = oracle.solve2d_V2(d, s.ni, opts)
    assert out["status"] == 0
    assert out["err_evo1"][-1] < 1.0e-8
    assert out["iter"] < s.kwargs["iterMax"]
    for c in ("xx", "yy", "xy"):
        assert np.array_equal(d["t" + c], d["t" + c + "_o"])


def run_solkz(oracle, s):
    d = oracle.alloc_stokes(s.ni, s.fields)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts), 0)
    return d, oracle.solve2d_V2(d, s.ni, opts)


def test_solkz_reference_criterion(oracle):
    """test/test_stokes_solkz.jl:26-37 (2D-V2, viscosity exp(B·y) over six decades, Re = 5π): err_evo1[end] < 1e-8"""
    s = setups.solkz2d(32, 32)
    d, out = run_solkz(oracle, s)
    assert out["status"] == 0 and out["err_evo1"][-1] < 1.0e-8 and out["iter"] < s.kwargs["iterMax"]


def test_v2_iteration_matches_numpy_restatement(oracle):
    """independent cross-check of the C oracle: one 2D-V2 iteration written with numpy slices"""
    rng = np.random.default_rng(4)
    nx, ny = 9, 7
    ni = (nx, ny)
    U = lambda *s: np.asfortranarray(rng.uniform(-1, 1, size=s))
    f = dict(Vx=U(nx + 1, ny + 2), Vy=U(nx + 2, ny + 1), P=U(*ni), P0=U(*ni), Q=U(*ni) * 0.1, txx=U(*ni), tyy=U(*ni), txy=U(nx + 1, ny + 1),
             txx_o=U(*ni), tyy_o=U(*ni), txy_o=U(nx + 1, ny + 1), eta=np.asfortranarray(10.0 ** rng.uniform(-2, 0, size=ni)),
             G=np.asfortranarray(rng.uniform(0.5, 2, size=ni)), K=np.asfortranarray(rng.uniform(1, 4, size=ni)), rhogx=U(*ni), rhogy=U(*ni))
    d = oracle.alloc_stokes(ni, f)
    ref = {k: v.copy(order="F") for k, v in d.items()}
    from justrelax_jl_b200.types import Geometry, PTStokesCoeffs
    li = (1.0, 1.2)
    grid = Geometry(ni, li)
    pt = PTStokesCoeffs(li, grid.di.center)
    dt = 0.6
    flags = dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6)
    opts = oracle.make_opts(pt, grid._di.center, dt, flags, ni, iterMax=1, nout=1)
    oracle.iterate2d_V2(d, ni, opts, 1)
    _dx, _dy = grid._di.center
    Vx, Vy, eta, G, K = ref["Vx"], ref["Vy"], ref["eta"], ref["G"], ref["K"]
    ett = np.zeros(ni)
    pe = np.pad(eta, 1, mode="edge")
    for a in range(3):
        for b in range(3):
            ett = np.maximum(ett, pe[a:a + nx, b:b + ny]) if (a or b) else pe[a:a + nx, b:b + ny].copy()
    divV = (Vx[1:, 1:-1] - Vx[:-1, 1:-1]) * _dx + (Vy[1:-1, 1:] - Vy[1:-1, :-1]) * _dy
    psi = 1.0 / (1.0 / ett + 1.0 / (G * dt)) * pt.r / pt.θ_dτ
    P = ((ref["P0"] / (K * dt) - divV + ref["Q"] / dt) * psi + ref["P"]) / (1 + psi / (K * dt))
    exx = (Vx[1:, 1:-1] - Vx[:-1, 1:-1]) * _dx - divV / 3
    exy = 0.5 * (_dy * (Vx[:, 1:] - Vx[:, :-1]) + _dx * (Vy[1:, :] - Vy[:-1, :]))
    av = lambda A: 0.25 * (np.pad(A, 1, mode="edge")[:-1, :-1] + np.pad(A, 1, mode="edge")[1:, :-1] + np.pad(A, 1, mode="edge")[:-1, 1:] + np.pad(A, 1, mode="edge")[1:, 1:])
    upd = lambda t, to, e, et, g: t + (1.0 / (pt.θ_dτ + et / (g * dt) + 1.0)) * (2 * et * e - (t - to) * et / (g * dt) - t)
    txx, txy = upd(ref["txx"], ref["txx_o"], exx, eta, G), upd(ref["txy"], ref["txy_o"], exy, av(eta), av(G))
    tol = lambda A: 1e-12 * np.abs(A).max()
    assert np.allclose(d["etatau"], ett, rtol=0, atol=0)
    assert np.allclose(d["P"], P, rtol=0, atol=tol(P)) and np.allclose(d["txx"], txx, rtol=0, atol=tol(txx)) and np.allclose(d["txy"], txy, rtol=0, atol=tol(txy))
    Rx = (d["txx"][1:] - d["txx"][:-1]) * _dx + (d["txy"][1:-1, 1:] - d["txy"][1:-1, :-1]) * _dy - (d["P"][1:] - d["P"][:-1]) * _dx - 0.5 * (ref["rhogx"][1:] + ref["rhogx"][:-1])
    assert np.allclose(d["Rx"], Rx, rtol=0, atol=tol(Rx))
    Vxn = ref["Vx"][1:-1, 1:-1] + Rx * pt.ηdτ / (0.5 * (ett[1:] + ett[:-1]))
    assert np.allclose(d["Vx"][1:-1, 1:-1], Vxn, rtol=0, atol=tol(Vxn))


def particle_equivalent_ratios(s, seed, nxcell=20):
    """JustPIC-style phase ratios for the sinking block: nxcell particles per cell at uniformly random positions inside 5–95 % of the cell
    (init_particles), phase by position (test_sinking_block.jl:60-79), ratios = bilinear-weighted particle fractions at the centres (own
    cell) and at the vertices (the four surrounding cells) — phase_ratio_weights / bilinear_weight of JustPIC (third party, not vendored)"""
    rng = np.random.default_rng(seed)
    n = s.ni[0]
    dx, dy = s.di
    (xc, yc), (xv, yv) = s.grid.xci, s.grid.xvi
    px = xv[:-1, None, None] + dx * (rng.random((n, n, nxcell)) * 0.9 + 0.05)
    py = yv[None, :-1, None] + dy * (rng.random((n, n, nxcell)) * 0.9 + 0.05)
    ph2 = (((px - 250e3) ** 2 <= 50e3 ** 2) & ((-py - 100e3) ** 2 <= 50e3 ** 2)).astype(float)
    w = (1 - np.abs(px - xc[:, None, None]) / dx) * (1 - np.abs(py - yc[None, :, None]) / dy)
    f2c = (w * ph2).sum(-1) / w.sum(-1)
    num, den = np.zeros((n + 1, n + 1)), np.zeros((n + 1, n + 1))
    for a in (0, 1):
        for b in (0, 1):
            wv = (1 - np.abs(px - xv[a:n + a, None, None]) / dx) * (1 - np.abs(py - yv[None, b:n + b, None]) / dy)
            num[a:n + a, b:n + b] += (wv * ph2).sum(-1)
            den[a:n + a, b:n + b] += wv.sum(-1)
    oh = lambda f: np.asfortranarray(np.stack([1 - f, f], axis=-1))
    return dict(center=oh(f2c), vertex=oh(num / den))


def test_sinking_block_particle_equivalent_phase_ratios(oracle):
    """What the 5.6 % between the restatement (5.11e-10, grid-sampled ratios) and the reference's golden (4.84e-10, JustPIC particles,
    unseeded RNG, atol 1e-6) can and cannot be: (a) NOT the PT tolerance — the maximum vertex speed is converged to 1e-9 relative after
    2000 of the 3000 iterations; (b) NOT the sampling noise of 20 random particles per cell — particle-equivalent ratios (two seeds here,
    six in the analysis of DESIGN.md) move the result by ≈ ±1 %: 4.97e-10 … 5.12e-10; the remaining ≈ 4 % is systematic and sits in the
    unpinned JustPIC particle / phase-ratio kernels (or GeoParams' phase-mixed viscosity), which only a Julia run can pin."""
    base = None
    for seed in (0, 5):
        s = setups.sinking_block2d(32)
        s.ratios = particle_equivalent_ratios(s, seed)
        rho = s.ratios["center"][..., 0] * 3.2e3 + s.ratios["center"][..., 1] * 3.3e3
        s.fields["rhogy"] = np.asfortranarray(rho * 9.81)
        s.fields["P"] = np.asfortranarray(s.fields["rhogy"] * np.abs(s.grid.xci[1])[None, :])
        d, out = run_sinking_block(oracle, s)
        assert out["status"] == 0 and out["err_evo1"][-1] < 1.0e-5
        v = vertex_speed(d["Vx"], d["Vy"]).max()
        assert abs(v / 5.1146e-10 - 1) < 0.04, (seed, v)          # within particle noise of the grid-sampled restatement
        assert abs(v / 4.841885609356093e-10 - 1) < 0.07, (seed, v)   # and within 7 % of the reference's golden
        base = v if base is None else base


def test_shearband2d_softening_reference_golden(oracle):
    """test/test_shearband2D_softening.jl:197-204: cohesion softening law present, five steps of dt = 0.05: err < 1e-6, maximum(τxx) ≈ 0.466
    (atol 1e-3) and the visco-elastic build-up 2 ε η (1 − exp(−G t/η)) ≈ 0.4423 (atol 1e-4)"""
    s = setups.shearband2d_softening(32)
    d, outs, txx_max = run_shearband(oracle, s)
    assert all(o["status"] == 0 for o in outs) and outs[-1]["err_evo1"][-1] < 1.0e-6
    assert abs(txx_max[-1] - 0.466) < 1.0e-3, txx_max[-1]
    assert abs(s.solution(5 * s.dt) - 0.4423) < 1.0e-4


def test_softening_laws(oracle):
    """GeoParams LinearSoftening / NonLinearSoftening as lowered into the flat table: limits and mid-points; and softening that engages
    (large accumulated plastic strain lowers the cohesion, so a stress state that is elastic with EII = 0 yields)"""
    lo, hi, mx, mn = 0.1, 0.5, 2.0, 1.0
    kind, p = R.LinearSoftening((mn, mx), (lo, hi)).params()
    lin = lambda x: mn if x >= p[1] else (mx if x <= p[0] else x * p[4] + p[5])
    assert kind == 1 and lin(0.0) == mx and lin(1.0) == mn and abs(lin(0.3) - 1.5) < 1e-15 and abs(lin(lo + 1e-12) - mx) < 1e-9
    kind, q = R.NonLinearSoftening(ξ0=1.6, Δ=0.8).params()
    nl = lambda x: q[0] - 0.5 * q[1] * math.erfc(-(x - q[2]) / q[3])
    assert kind == 2 and abs(nl(0.0) - (1.6 - 0.4 * math.erfc(2.0))) < 1e-15 and abs(nl(1.0) - 1.2) < 1e-15 and abs(nl(10.0) - 0.8) < 1e-12
    # engagement: one iteration from the same state with EII = 0 and EII = 3 (C: 1.6/cos30 → ≈ 0.8)
    s = setups.shearband2d_softening(16)
    res = {}
    for EII in (0.0, 3.0):
        d = oracle.alloc_stokes(s.ni, s.fields)
        d["txx"][...] = 1.2
        d["tyy"][...] = -1.2
        d["txx_o"][...] = 1.2
        d["tyy_o"][...] = -1.2
        d["EII_pl"][...] = EII
        vc = oracle.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), s.ratios)
        opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=1, nout=1)
        oracle.iterate2d_VC(d, s.ni, opts, vc, 1)
        res[EII] = d["lam"].max()
    assert res[0.0] == 0.0 and res[3.0] > 0.0, res
