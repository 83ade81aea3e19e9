"""GPU parity tests for the 2D Stokes variants (2D-V2: config 2 SolCx; 2D-VC: config 3 shear band) through the C ABI against
the CPU oracle.  Tolerances (north star): per-field max relative difference <= 1e-12 after a fixed number of PT iterations,
iteration count to convergence within ±1 %, converged fields within 1e-8; plus the reference's own golden values
(test/test_stokes_solcx.jl:26-43, test/test_shearband2D.jl:194-202).
"""
import ctypes as C
import math

import numpy as np
import pytest

from util import bc_flags, compare_slots, device_stokes, max_rel_diff

pytestmark = pytest.mark.gpu
TOL = 1.0e-12
V2_STATE = ["Vx", "Vy", "P", "txx", "tyy", "txy"]
V2_DIAG = ["divV", "RP", "exx", "eyy", "exy", "Rx", "Ry", "Ux", "Uy", "etatau"]
VC_STATE = V2_STATE + ["txy_c", "eta", "etav", "lam", "lamv"]
VC_DIAG = V2_DIAG + ["pxx", "pyy", "pxy", "tII", "eta_vep", "e_vol_pl", "P0", "rhogx", "rhogy"]
NAMES6 = ("left", "right", "front", "back", "top", "bot")


def _bcs(flags, displacement=False):
    from justrelax_jl_b200.types import DisplacementBoundaryConditions, VelocityBoundaryConditions

    pick = lambda nm: {k: bool(v) for k, v in zip(NAMES6, flags[nm]) if k in ("left", "right", "top", "bot")}
    T = DisplacementBoundaryConditions if displacement else VelocityBoundaryConditions
    return T(free_slip=pick("free_slip"), no_slip=pick("no_slip"), periodic=pick("periodic"))


def random_stokes2d(ni, seed, *, dt=0.6, finite=True):
    from justrelax_jl_b200.types import Geometry, PTStokesCoeffs

    rng = np.random.default_rng(seed)
    nx, ny = ni
    U = lambda *s: np.asfortranarray(rng.uniform(-1, 1, size=s))
    f = dict(Vx=U(nx + 1, ny + 2), Vy=U(nx + 2, ny + 1), P=np.asfortranarray(rng.uniform(0, 1, size=ni)), P0=np.asfortranarray(rng.uniform(0, 1, size=ni)),
             Q=U(*ni) * 0.1, txx=U(*ni), tyy=U(*ni), txy=U(nx + 1, ny + 1), txx_o=U(*ni), tyy_o=U(*ni), txy_o=U(nx + 1, ny + 1),
             eta=np.asfortranarray(10.0 ** rng.uniform(-3, 0, size=ni)),
             G=np.asfortranarray(rng.uniform(0.5, 2, size=ni)) if finite else np.full(ni, np.inf, order="F"),
             K=np.asfortranarray(rng.uniform(1, 4, size=ni)) if finite else np.full(ni, np.inf, order="F"), rhogx=U(*ni), rhogy=U(*ni))
    li = (1.0, 1.2)
    grid = Geometry(ni, li)
    return f, grid, PTStokesCoeffs(li, grid.di.center), dt


@pytest.mark.parametrize("ni", [(9, 7), (33, 17), (64, 64), (95, 130)])
@pytest.mark.parametrize("finite", [True, False])
def test_v2_fixed_iterations_random_state(oracle, ni, finite):
    from justrelax_jl_b200 import stokes as jst

    f, grid, pt, dt = random_stokes2d(ni, 20261017 + ni[0], finite=finite)
    for flags in (dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6),
                  dict(free_slip=[1, 0, 0, 0, 0, 1], no_slip=[0, 1, 0, 0, 1, 0], periodic=[0] * 6),
                  dict(free_slip=[0, 0, 0, 0, 1, 0], no_slip=[0, 0, 0, 0, 0, 1], periodic=[0] * 6)):
        for niter in (1, 2, 5):
            d = oracle.alloc_stokes(ni, f)
            st, extra = device_stokes(ni, d)
            opts = oracle.make_opts(pt, grid._di.center, dt, flags, ni, iterMax=niter, nout=niter)
            oracle.iterate2d_V2(d, ni, opts, niter)
            r = jst.iterate2d_V2_(st, pt, grid, _bcs(flags), (extra["rhogx"], extra["rhogy"]), extra["G"], extra["K"], dt, niter)
            assert r.kernel_launches > 0
            compare_slots(st.slots(), d, V2_STATE + V2_DIAG, TOL, f"V2 ni={ni} niter={niter} flags={flags}")


@pytest.mark.parametrize("ni", [(9, 7), (33, 17), (64, 64), (95, 130), (257, 131)])
@pytest.mark.parametrize("dt", [0.6, np.inf])
def test_v2_resident_batches(oracle, monkeypatch, ni, dt):
    """non-observable 2D-V2 iterations with the state resident in shared memory (one persistent CTA per tile, velocities exchanged through
    L2, neighbour flags): same result as the oracle and as the one-launch-per-iteration kernel, for every boundary-flag combination"""
    from justrelax_jl_b200 import stokes as jst

    f, grid, pt, _ = random_stokes2d(ni, 99 + ni[0], finite=False)
    if np.isfinite(dt):
        f["Q"] = np.zeros(ni, order="F")   # the resident kernel takes over when 1/(G dt) = 1/(K dt) = Q/dt = 0 everywhere
    for flags in (dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6),
                  dict(free_slip=[1, 0, 0, 0, 0, 1], no_slip=[0, 1, 0, 0, 1, 0], periodic=[0] * 6),
                  dict(free_slip=[0, 0, 0, 0, 1, 0], no_slip=[0, 0, 0, 0, 0, 1], periodic=[0] * 6),
                  dict(free_slip=[0] * 6, no_slip=[0] * 6, periodic=[0] * 6)):
        launches = {}
        for resident in ("1", "0"):
            monkeypatch.setenv("JRB200_2D_RESIDENT", resident)
            for niter in (3, 10):
                d = oracle.alloc_stokes(ni, f)
                st, extra = device_stokes(ni, d)
                opts = oracle.make_opts(pt, grid._di.center, dt, flags, ni, iterMax=niter, nout=niter)
                oracle.iterate2d_V2(d, ni, opts, niter)
                r = jst.iterate2d_V2_(st, pt, grid, _bcs(flags), (extra["rhogx"], extra["rhogy"]), extra["G"], extra["K"], dt, niter)
                compare_slots(st.slots(), d, V2_STATE + V2_DIAG, TOL, f"V2 resident={resident} ni={ni} dt={dt} niter={niter} flags={flags}")
            launches[resident] = r.kernel_launches
        assert launches["1"] < launches["0"] - 5, launches


def test_v2_periodic(oracle):
    from justrelax_jl_b200 import stokes as jst

    ni = (21, 18)
    f, grid, pt, dt = random_stokes2d(ni, 3)
    flags = dict(free_slip=[0, 0, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[1, 1, 0, 0, 0, 0])
    d = oracle.alloc_stokes(ni, f)
    st, extra = device_stokes(ni, d)
    opts = oracle.make_opts(pt, grid._di.center, dt, flags, ni, iterMax=3, nout=3)
    oracle.iterate2d_V2(d, ni, opts, 3)
    jst.iterate2d_V2_(st, pt, grid, _bcs(flags), (extra["rhogx"], extra["rhogy"]), extra["G"], extra["K"], dt, 3)
    compare_slots(st.slots(), d, V2_STATE + V2_DIAG, TOL, "V2 periodic")


def test_solcx_solve_matches_reference_test(oracle):
    """test/test_stokes_solcx.jl:26-43 at 32²: converges below 1e-8; iteration count, history and fields as the oracle"""
    from justrelax_jl_b200 import setups, stokes as jst

    s = setups.solcx2d(32, 32)
    d = oracle.alloc_stokes(s.ni, s.fields)
    st, extra = device_stokes(s.ni, d)
    jst.flow_bcs_(st, s.flow_bcs)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts), 0)
    ref = oracle.solve2d_V2(d, s.ni, opts)
    out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, (extra["rhogx"], extra["rhogy"]), extra["G"], extra["K"], s.dt, s.igg, kwargs=s.kwargs)
    assert out.err_evo1[-1] < 1.0e-8
    assert abs(out.iter - ref["iter"]) <= 0.01 * ref["iter"]
    assert np.array_equal(out.err_evo2, ref["err_evo2"])
    assert np.allclose(out.norm_divV, ref["norm_divV"], rtol=1e-8)
    compare_slots(st.slots(), d, V2_STATE + ["txx_o", "txy_o"], 1.0e-8, "solcx converged fields")


# ---------------------------------------------------------------------------------------------------------------------------
def random_vc2d(ni, seed, nphase=3, plastic=True, rho_var=False):
    """random multiphase VEP state: Dirichlet phase ratios with exact zeros / ones sprinkled in, yielding stresses"""
    from justrelax_jl_b200 import rheology as R

    rng = np.random.default_rng(seed)
    f, grid, pt, dt = random_stokes2d(ni, seed)
    nx, ny = ni
    f = {k: v for k, v in f.items() if k not in ("G", "K")}
    f["txy_c"], f["txy_o_c"] = np.asfortranarray(rng.uniform(-1, 1, size=ni)), np.asfortranarray(rng.uniform(-1, 1, size=ni))
    f["eta"] = np.asfortranarray(10.0 ** rng.uniform(-1, 0, size=ni))
    f["etav"] = np.asfortranarray(10.0 ** rng.uniform(-1, 0, size=(nx + 1, ny + 1)))
    f["T"] = np.asfortranarray(rng.uniform(300, 1500, size=(nx + 2, ny + 2)))
    f["Pargs"] = np.asfortranarray(rng.uniform(0, 1, size=ni))

    def ratios(shape):
        r = rng.dirichlet(np.ones(nphase), size=shape)
        pure = rng.uniform(size=shape) < 0.4
        one = rng.integers(0, nphase, size=shape)
        r[pure] = np.eye(nphase)[one[pure]]
        if nphase > 2:  # exact zero of one phase in a mixed node
            z = (~pure) & (rng.uniform(size=shape) < 0.3)
            r[z, 0] = 0.0
            r[z] /= r[z].sum(axis=-1, keepdims=True)
        return np.asfortranarray(r)

    rat = dict(center=ratios(ni), vertex=ratios((nx + 1, ny + 1)))
    rheo = []
    for p in range(nphase):
        el = R.ConstantElasticity(G=float(rng.uniform(0.5, 2)), Kb=(math.inf if p == 1 else float(rng.uniform(1, 4))))
        els = [R.LinearViscous(η=float(10.0 ** rng.uniform(-1, 0))), el]
        if plastic and p != 2:
            els.append(R.DruckerPrager_regularised(C=float(rng.uniform(0.05, 0.3)), ϕ=float(rng.uniform(10, 35)), η_vp=float(rng.uniform(1e-3, 1e-2)),
                                                   Ψ=float(rng.uniform(0, 10))))
        dens = R.PT_Density(ρ0=float(rng.uniform(1, 3)), α=1e-4, β=1e-2, T0=300.0, P0=0.1) if rho_var and p != 1 else R.ConstantDensity(ρ=float(rng.uniform(1, 3)))
        rheo.append(R.SetMaterialParams(Phase=p + 1, Density=dens, Gravity=R.ConstantGravity(g=1.5), CompositeRheology=R.CompositeRheology(els), Elasticity=el))
    return f, grid, pt, dt, rat, tuple(rheo)


def _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, niter, finish, free_surface=False, alias_P=False, dT=None, inc=False, dbc=False):
    from justrelax_jl_b200 import B200Backend, PhaseRatios, rheology as R, stokes as jst

    d = oracle.alloc_stokes(ni, f)
    if alias_P:
        d["Pargs"] = d["P"]
    if dT is not None:
        d["dTargs"] = dT
    st, extra = device_stokes(ni, d)
    rows = R.lower_stokes(rheo)
    vc = oracle.vc_inputs(rows, R.gravity_of(rheo), rat, free_surface=dt if free_surface else 0.0)
    opts = oracle.make_opts(pt, grid._di.center, dt, flags, ni, iterMax=niter, nout=niter, viscosity_relaxation=0.3, lambda_relaxation=0.2,
                            strain_increment=int(inc), displacement_bcs=int(dbc), dT_ghosted=int(dT is not None and dT.shape != tuple(ni)))
    oracle.iterate2d_VC(d, ni, opts, vc, niter, finish=finish)
    pr = PhaseRatios.from_arrays(B200Backend, **rat)
    args = dict(T=extra["T"], P=st.P if alias_P else extra["Pargs"])
    if dT is not None:
        args["ΔT"] = extra["dTargs"]
    jst.iterate2d_VC_(st, pt, grid, _bcs(flags, dbc), (extra["rhogx"], extra["rhogy"]), pr, rheo, args, dt, niter, finish=finish,
                      kwargs=dict(viscosity_relaxation=0.3, free_surface=free_surface, strain_increment=inc))
    return {**st.slots(), "rhogx": extra["rhogx"], "rhogy": extra["rhogy"]}, d


@pytest.mark.parametrize("ni", [(9, 7), (33, 17), (64, 64), (95, 130)])
@pytest.mark.parametrize("rho_var", [False, True])
def test_vc_fixed_iterations_random_state(oracle, ni, rho_var):
    f, grid, pt, dt, rat, rheo = random_vc2d(ni, 77 + ni[0], rho_var=rho_var)
    flags = dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6)
    for niter in (1, 2, 5):
        st, d = _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, niter, False, alias_P=rho_var)
        assert d["lam"].max() > 0 and d["lamv"].max() > 0, "the random state must yield somewhere"
        compare_slots(st, d, VC_STATE + VC_DIAG, TOL, f"VC ni={ni} niter={niter}")


def test_vc_thermal_stress_pressure_form(oracle):
    """args.ΔT given: thermal-stress form of compute_P! (PressureKernels.jl:128-149,197-206) in the fused 2D kernel"""
    from justrelax_jl_b200 import to_host

    ni = (45, 38)
    f, grid, pt, dt, rat, rheo = random_vc2d(ni, 12, rho_var=True)
    flags = dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6)
    dT = np.asfortranarray(np.random.default_rng(6).uniform(-2.0e3, 2.0e3, size=ni))
    for niter in (1, 4):
        st, d = _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, niter, False, alias_P=True, dT=dT)
        compare_slots(st, d, VC_STATE + VC_DIAG, TOL, f"2D-VC with ΔT niter={niter}")
    dTg = np.asfortranarray(np.random.default_rng(7).uniform(-2.0e3, 2.0e3, size=tuple(n + 2 for n in ni)))   # thermal.ΔT (ni .+ 2)
    stg, dg = _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, 4, False, alias_P=True, dT=dTg)
    compare_slots(stg, dg, VC_STATE + VC_DIAG, TOL, "2D-VC with ghosted ΔT")
    st0, d0 = _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, 4, False, alias_P=True)
    assert max_rel_diff(to_host(st["P"]), d0["P"]) > 1e-4, "ΔT must change the pressure"


def test_vc_cohesion_softening(oracle):
    """cohesion softening with the accumulated plastic strain (StressUpdate.jl:305-332; Linear / NonLinearSoftening) in the fused kernel:
    EII at the centres and interpolated to the vertices, both laws, engaged (random EII across the softening range)"""
    from justrelax_jl_b200 import rheology as R

    ni = (47, 38)
    f, grid, pt, dt, rat, rheo = random_vc2d(ni, 21, rho_var=False)
    f["EII_pl"] = np.asfortranarray(np.random.default_rng(2).uniform(0.0, 2.0, size=ni))
    laws = [R.LinearSoftening((0.02, 0.3), (0.2, 1.5)), R.NonLinearSoftening(ξ0=0.25, Δ=0.2)]
    mats = []
    for m in rheo:
        els = []
        for e in m.CompositeRheology.elements:
            if isinstance(e, R.DruckerPrager_regularised):
                e = R.DruckerPrager_regularised(C=e.C, ϕ=e.ϕ, η_vp=e.η_vp, Ψ=e.Ψ, softening_C=laws[len(mats) % 2])
            els.append(e)
        mats.append(R.SetMaterialParams(Phase=m.Phase, Density=m.Density, Gravity=m.Gravity, CompositeRheology=R.CompositeRheology(tuple(els)),
                                        Elasticity=m.Elasticity))
    flags = dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6)
    for niter in (1, 4):
        st, d = _run_vc(oracle, ni, f, grid, pt, dt, rat, tuple(mats), flags, niter, False)
        assert d["lam"].max() > 0 and d["lamv"].max() > 0
        compare_slots(st, d, VC_STATE + VC_DIAG, TOL, f"VC softening niter={niter}")
    st0, d0 = _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, 4, False)
    assert max_rel_diff(d["lam"], d0["lam"]) > 1e-3, "softening must change the plastic multiplier"


def test_softening_unsupported_in_3d():
    from justrelax_jl_b200 import rheology as R

    pl = R.DruckerPrager_regularised(C=1.0, ϕ=30, η_vp=1e-2, softening_C=R.NonLinearSoftening(ξ0=1.0, Δ=0.5))
    m = (R.SetMaterialParams(Phase=1, Density=R.ConstantDensity(ρ=1.0), CompositeRheology=R.CompositeRheology((R.LinearViscous(η=1.0), pl))),)
    assert R.lower_stokes(m, 2)[0]["soft_C_kind"] == 2
    with pytest.raises(R.UnsupportedRheology, match="2D multiphase solve only"):
        R.lower_stokes(m, 3)


@pytest.mark.parametrize("inc,dbc", [(True, False), (True, True), (False, True)])
@pytest.mark.parametrize("ni", [(9, 7), (64, 47), (95, 130)])
def test_vc_strain_increment_and_displacement_bcs(oracle, ni, inc, dbc):
    """kwarg strain_increment = true (Δε form, Stokes2D.jl:659-730, StressKernels.jl:1147-1302) and DisplacementBoundaryConditions
    (flow_bcs! on U, V = U/dt before the loop: types/displacement.jl:33-70) in the fused kernel, free-slip / no-slip mixes"""
    f, grid, pt, dt, rat, rheo = random_vc2d(ni, 300 + ni[0], rho_var=False)
    rng = np.random.default_rng(ni[1])
    f["Ux"] = np.asfortranarray(rng.uniform(-1, 1, size=f["Vx"].shape)) * dt
    f["Uy"] = np.asfortranarray(rng.uniform(-1, 1, size=f["Vy"].shape)) * dt
    extra_names = ["Ux", "Uy"] + (["dxx", "dyy", "dxy", "divU"] if inc else [])
    for flags in (dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6),
                  dict(free_slip=[1, 0, 0, 0, 0, 1], no_slip=[0, 1, 0, 0, 1, 0], periodic=[0] * 6)):
        for niter, finish in ((1, False), (4, False), (5, True)):
            st, d = _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, niter, finish, inc=inc, dbc=dbc)
            assert d["lam"].max() > 0 and d["lamv"].max() > 0
            names = VC_STATE + VC_DIAG + extra_names + (["dxy_c", "exy_c", "txx_o"] if finish else [])
            compare_slots(st, d, names, TOL, f"VC inc={inc} dbc={dbc} ni={ni} niter={niter} flags={flags}")


def test_vc_exit_kernels_free_surface_and_mixed_bcs(oracle):
    ni = (40, 27)
    f, grid, pt, dt, rat, rheo = random_vc2d(ni, 5, rho_var=True)
    flags = dict(free_slip=[1, 0, 0, 0, 0, 1], no_slip=[0, 1, 0, 0, 0, 0], periodic=[0] * 6)
    st, d = _run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, 4, True, free_surface=True)
    compare_slots(st, d, VC_STATE + VC_DIAG + ["wxy", "exy_c", "pxy_c", "EII_pl", "EVol_pl", "txx_o", "tyy_o", "txy_o", "txy_o_c"], TOL, "VC exit")


def test_shearband2d_reference_golden_on_gpu(oracle):
    """config 3 at the reference test's size (32², 10 steps) through the public API: the reference goldens
    (test/test_shearband2D.jl:194-202) and the oracle's iteration counts / fields."""
    from justrelax_jl_b200 import B200Backend, PhaseRatios, rheology as R, setups, stokes as jst, to_host
    from test_oracle_stokes2d import run_shearband

    s = setups.shearband2d(32)
    d, outs, txx_max = run_shearband(oracle, s)
    d0 = oracle.alloc_stokes(s.ni, s.fields)
    st, extra = device_stokes(s.ni, d0)
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    args = dict(T=extra["T"], P=st.P)
    jst.compute_viscosity_(st, pr, args, s.rheology, (-math.inf, math.inf))
    jst.flow_bcs_(st, s.flow_bcs)
    ρg = (extra["rhogx"], extra["rhogy"])
    iters = []
    for _ in range(s.nt):
        out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, args, s.dt, s.igg, kwargs=s.kwargs)
        iters.append(out.iter)
    assert out.err_evo1[-1] < 1.0e-6
    jst.tensor_invariant_(st.τ, s.ni)
    tII = to_host(st.τ.II)
    assert abs(tII.min() - 1.5128689768248313) < 1.0e-3
    assert abs(tII.max() - 1.6415759440014273) < 1.0e-3
    assert abs(to_host(st.τ.xx).max() - 1.6376258215356436) < 1.0e-4
    for a, b in zip(iters, [o["iter"] for o in outs]):
        assert abs(a - b) <= max(0.01 * b, 1)
    compare_slots(st.slots(), d, ["Vx", "Vy", "P", "txx", "tyy", "txy", "EII_pl"], 1.0e-6, "shear band final fields")


def test_sinking_block_reference_golden_on_gpu(oracle):
    """test/test_sinking_block.jl through the public API: 2D-VC with buoyancy in SI units (η = 1e21 | 1e23, G = Kb = Inf, P ~ 1e10):
    the oracle's iteration count and fields, and the reference's golden maximum velocity."""
    from justrelax_jl_b200 import B200Backend, PhaseRatios, setups, stokes as jst, to_host
    from test_oracle_stokes2d import run_sinking_block, vertex_speed

    s = setups.sinking_block2d(32)
    d, out_o = run_sinking_block(oracle, s)
    d0 = oracle.alloc_stokes(s.ni, s.fields)
    st, extra = device_stokes(s.ni, d0)
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    args = dict(T=extra["T"], P=st.P)
    jst.compute_viscosity_(st, pr, args, s.rheology, (-math.inf, math.inf))
    jst.flow_bcs_(st, s.flow_bcs)
    out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, (extra["rhogx"], extra["rhogy"]), pr, s.rheology, args, s.dt, s.igg, kwargs=s.kwargs)
    assert out.err_evo1[-1] < 1.0e-5
    assert out.iter == out_o["iter"]
    vmax = vertex_speed(to_host(st.V.Vx), to_host(st.V.Vy)).max()
    assert abs(vmax - 4.841885609356093e-10) < 1.0e-6 and abs(vmax / 4.841885609356093e-10 - 1) < 0.1
    compare_slots(st.slots(), d, ["Vx", "Vy", "P", "txx", "tyy", "txy"], 1.0e-8, "sinking block converged fields")


def test_compute_dt_reference_kat():
    """test/test_Utils.jl:146-148: ni = (4, 4), di = (0.25, 0.25), Vx = 0, Vy = 10 → compute_dt === 0.022500000000000003"""
    from justrelax_jl_b200 import B200Backend, StokesArrays, stokes as jst

    st = StokesArrays(B200Backend, 4, 4)
    st.V.Vy.fill_(10.0)
    di = (0.25, 0.25)
    assert jst.compute_dt_(st, di, 0.1) == 0.022500000000000003
    assert jst.compute_dt_(st, di) == 0.022500000000000003


def test_standalone_2d_kernels(oracle):
    from justrelax_jl_b200 import B200Backend, PTArray, StokesArrays, stokes as jst, to_host

    rng = np.random.default_rng(2)
    ni = (13, 11)
    st = StokesArrays(B200Backend, *ni)
    h = {k: np.asfortranarray(rng.uniform(size=tuple(v.shape))) for k, v in (("Vx", st.V.Vx), ("Vy", st.V.Vy))}
    for flags in (dict(free_slip=[1, 0, 0, 0, 0, 1], no_slip=[0, 1, 0, 0, 1, 0], periodic=[0] * 6),
                  dict(free_slip=[0] * 6, no_slip=[0] * 6, periodic=[1, 1, 0, 0, 1, 1])):
        for k in h:
            st.slots()[k].copy_(PTArray(B200Backend)(h[k]))
        jst.flow_bcs_(st, _bcs(flags))
        d = oracle.alloc_stokes(ni, h)
        opts = oracle.make_opts(type("pt", (), dict(r=1, θ_dτ=1, ηdτ=1, ϵ_rel=1, ϵ_abs=1)), (1, 1), 1.0, flags, ni, iterMax=1, nout=1)
        fs = oracle.make_fields(d, ni)
        oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts), 0)
        for k in h:
            assert np.array_equal(to_host(st.slots()[k]), d[k]), (k, flags)
    xx, yy, xy = (np.asfortranarray(rng.uniform(-1, 1, size=s)) for s in (ni, ni, (ni[0] + 1, ni[1] + 1)))
    st.τ.xx.copy_(PTArray(B200Backend)(xx)); st.τ.yy.copy_(PTArray(B200Backend)(yy)); st.τ.xy.copy_(PTArray(B200Backend)(xy))
    jst.tensor_invariant_(st.τ, ni)
    assert max_rel_diff(to_host(st.τ.II), oracle.tensor_invariant2d(xx, yy, xy)) <= 1e-15
