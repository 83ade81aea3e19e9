"""GPU parity tests for the 3D multiphase visco-elasto-plastic Stokes solve (variant 3D-VC, config 5's Stokes half) through
the C ABI against the CPU oracle.  Tolerances (north star): per-field max relative difference <= 1e-12 after a fixed number of
PT iterations, iteration count to convergence within ±1 %, converged fields within 1e-8.
"""
import ctypes as C
import math

import numpy as np
import pytest

from util import bc_flags, compare_slots, device_stokes, max_rel_diff

pytestmark = pytest.mark.gpu
TOL = 1.0e-12
STATE = ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "tyz_c", "txz_c", "txy_c", "eta", "lam"]
DIAG = ["divV", "RP", "exx", "eyy", "ezz", "eyz", "exz", "exy", "pxx", "pyy", "pzz", "pyz", "pxz", "pxy", "tII", "eta_vep", "e_vol_pl",
        "Rx", "Ry", "Rz", "Ux", "Uy", "Uz", "etatau", "P0", "rhogx", "rhogy", "rhogz"]
EXIT = ["wyz", "wxz", "wxy", "eyz_c", "exz_c", "exy_c", "pyz_c", "pxz_c", "pxy_c", "EII_pl", "EVol_pl",
        "txx_o", "tyy_o", "tzz_o", "tyz_o", "txz_o", "txy_o", "tyz_o_c", "txz_o_c", "txy_o_c"]
NAMES6 = ("left", "right", "front", "back", "top", "bot")


def _bcs(flags):
    from justrelax_jl_b200.types import VelocityBoundaryConditions

    pick = lambda nm: {k: bool(v) for k, v in zip(NAMES6, flags[nm])}
    return VelocityBoundaryConditions(free_slip=pick("free_slip"), no_slip=pick("no_slip"), periodic=pick("periodic"))


def _run(oracle, s, flags, niter, finish, *, alias_P=True, kw=None, dT=None):
    from justrelax_jl_b200 import B200Backend, PhaseRatios, rheology as R
    from justrelax_jl_b200.stokes3d_vc import iterate3d_VC_

    kw = dict(viscosity_relaxation=0.3, λ_relaxation=0.2, viscosity_cutoff=s.kwargs["viscosity_cutoff"], **(kw or {}))
    d = oracle.alloc_stokes(s.ni, s.fields)
    if alias_P:
        d["Pargs"] = d["P"]
    if dT is not None:
        d["dTargs"] = dT
    st, extra = device_stokes(s.ni, d)
    vc = oracle.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), s.ratios)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, flags, s.ni, iterMax=niter, nout=niter, viscosity_relaxation=kw["viscosity_relaxation"],
                            lambda_relaxation=kw["λ_relaxation"], viscosity_cutoff=kw["viscosity_cutoff"],
                            dT_ghosted=int(dT is not None and dT.shape != tuple(s.ni)))
    oracle.iterate3d_VC(d, s.ni, opts, vc, niter, finish=finish)
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    args = dict(T=extra["T"], P=st.P if alias_P else extra["Pargs"], dt=s.dt)
    if dT is not None:
        args["ΔT"] = extra["dTargs"]
    ρg = (extra["rhogx"], extra["rhogy"], extra["rhogz"])
    r = iterate3d_VC_(st, s.pt_stokes, s.grid, _bcs(flags), ρg, pr, s.rheology, args, s.dt, niter, finish=finish, kwargs=kw)
    assert r.kernel_launches >= 3 * niter
    return {**st.slots(), "rhogx": ρg[0], "rhogy": ρg[1], "rhogz": ρg[2]}, d


@pytest.mark.parametrize("ni", [(7, 6, 5), (17, 9, 12), (33, 20, 15), (40, 40, 40)])
def test_vc3_fixed_iterations_random_state(oracle, ni):
    from justrelax_jl_b200 import setups

    s = setups.random_vc3d(ni, seed=100 + ni[0])
    flags = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)
    for niter in (1, 2, 5):
        st, d = _run(oracle, s, flags, niter, False)
        assert d["lam"].max() > 0 and np.abs(d["pyz"]).max() > 0, "the random state must yield somewhere"
        compare_slots(st, d, STATE + DIAG, TOL, f"3D-VC ni={ni} niter={niter}")


@pytest.mark.parametrize("rows,chunk", [("0", None), ("8", None), ("8", "3"), ("12", None), ("12", "5")])
@pytest.mark.parametrize("ni", [(70, 29, 21), (40, 40, 40), (33, 13, 3)])
def test_vc3_stress_kernel_variants(oracle, monkeypatch, ni, rows, chunk):
    """the z-marching stress kernel (interior) + the per-plane kernel on the rim slabs, for both tile heights and for z-chunks that do not
    divide the plane count, against the oracle — and the per-plane kernel alone (JRB200_VC3_ZM=0): all the same bits"""
    from justrelax_jl_b200 import setups

    monkeypatch.setenv("JRB200_VC3_ZM", rows)
    if chunk is not None:
        monkeypatch.setenv("JRB200_VC3_ZM_CHUNK", chunk)
    s = setups.random_vc3d(ni, seed=300 + ni[0])
    flags = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)
    st, d = _run(oracle, s, flags, 4, False)
    assert d["lam"].max() > 0 and np.abs(d["pyz"]).max() > 0, "the random state must yield somewhere"
    compare_slots(st, d, STATE + DIAG, TOL, f"3D-VC stress variants ni={ni} rows={rows} chunk={chunk}")


def test_vc3_thermal_stress_pressure_form(oracle):
    """args.ΔT given: compute_P! takes the thermal-stress form (PressureKernels.jl:128-149,197-206) with α = fn_ratio(get_thermal_expansion, …);
    parity with the oracle, and the pressure really differs from the run without ΔT"""
    from justrelax_jl_b200 import setups, to_host

    ni = (21, 18, 13)
    s = setups.random_vc3d(ni, seed=31)
    flags = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)
    dT = np.asfortranarray(np.random.default_rng(5).uniform(-50.0, 50.0, size=ni))
    for niter in (1, 3):
        st, d = _run(oracle, s, flags, niter, False, dT=dT)
        compare_slots(st, d, STATE + DIAG, TOL, f"3D-VC with ΔT niter={niter}")
    # ΔT passed as thermal.ΔT (ni .+ 2), the way the reference's scripts do (Blob3D.jl:280): indexed ΔT[I...] without offset
    dTg = np.asfortranarray(np.random.default_rng(6).uniform(-50.0, 50.0, size=tuple(n + 2 for n in ni)))
    st, d = _run(oracle, s, flags, 3, False, dT=dTg)
    compare_slots(st, d, STATE + DIAG, TOL, "3D-VC with ghosted ΔT")
    assert max_rel_diff(d["RP"], _run(oracle, s, flags, 3, False, dT=np.asfortranarray(dTg[1:-1, 1:-1, 1:-1]))[1]["RP"]) > 1e-6, "the quirk is an offset"

    st0, d0 = _run(oracle, s, flags, 3, False)
    assert max_rel_diff(to_host(st["P"]), d0["P"]) > 1e-3, "ΔT must change the pressure"


def test_vc3_exit_kernels_mixed_bcs_one_hot_phases(oracle):
    from justrelax_jl_b200 import setups

    s = setups.random_vc3d((19, 14, 11), seed=9, mixed=False)
    flags = dict(free_slip=[1, 0, 1, 1, 0, 1], no_slip=[0, 1, 0, 0, 1, 0], periodic=[0] * 6)
    st, d = _run(oracle, s, flags, 4, True, alias_P=False)
    compare_slots(st, d, STATE + DIAG + EXIT, TOL, "3D-VC exit")


def test_shearband3d_solve_matches_oracle(oracle):
    """the reference's 3D shear-band test setup (test/test_shearband3D_MPI.jl) at 12³ through the public API: iteration counts within ±1 %,
    converged fields within 1e-8 (relative to the field scale), plasticity active"""
    from justrelax_jl_b200 import B200Backend, PhaseRatios, rheology as R, setups, stokes as jst, to_host

    s = setups.shearband3d(12)
    d = oracle.alloc_stokes(s.ni, s.fields)
    st, extra = device_stokes(s.ni, d)
    vc = oracle.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), s.ratios)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"])
    fs = oracle.make_fields(d, s.ni)
    oracle.lib().orc_viscosity3d(C.byref(fs), C.byref(opts), C.byref(vc), C.c_double(1.0))
    oracle.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    args = dict(T=extra["T"], P=st.P)
    jst.compute_viscosity_(st, pr, args, s.rheology, (-math.inf, math.inf))
    jst.flow_bcs_(st, s.flow_bcs)
    ρg = (extra["rhogx"], extra["rhogy"], extra["rhogz"])
    for step in range(8):
        ref = oracle.solve3d_VC(d, s.ni, opts, vc)
        out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, args, s.dt, s.igg, kwargs=s.kwargs)
        assert abs(out.iter - ref["iter"]) <= max(0.01 * ref["iter"], 1), (step, out.iter, ref["iter"])
        assert np.allclose(out.err_evo1, ref["err_evo1"], rtol=1e-6)
    assert d["lam"].max() > 0
    compare_slots(st.slots(), d, ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "EII_pl", "tII"], 1.0e-8, "3D shear band final fields")
    jst.tensor_invariant_(st.ε, s.ni)          # tensor_invariant!(stokes.ε)  test_shearband3D_MPI.jl:209
    assert np.isfinite(to_host(st.ε.II)).all() and to_host(st.ε.II).max() > 0


def test_extruded_shearband_2d_reference_golden_through_3d_vc_on_gpu(oracle):
    """the external pin of 3D-VC (tests/test_oracle_stokes3d_vc.py) on the device: the reference's 2D shear-band test extruded along y,
    ten solve! calls through the public API, lands on the 2D golden of test/test_shearband2D.jl:197-201 to its own tolerances"""
    from justrelax_jl_b200 import B200Backend, PhaseRatios, setups, stokes as jst, to_host
    from test_oracle_stokes3d_vc import in_plane_invariant

    s = setups.shearband3d_extruded(32, 4)
    d = oracle.alloc_stokes(s.ni, s.fields)
    st, extra = device_stokes(s.ni, d)
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    args = dict(T=extra["T"], P=st.P)
    jst.compute_viscosity_(st, pr, args, s.rheology, (-math.inf, math.inf))
    jst.flow_bcs_(st, s.flow_bcs)
    ρg = (extra["rhogx"], extra["rhogy"], extra["rhogz"])
    txx_max = []
    for _ in range(s.nt):
        out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, args, s.dt, s.igg, kwargs=s.kwargs)
        assert out.err_evo1[-1] < 1.0e-6
        txx_max.append(float(to_host(st.τ.xx).max()))
    tII = in_plane_invariant(oracle, to_host(st.τ.xx), to_host(st.τ.zz), to_host(st.τ.xz))
    assert abs(tII.min() - 1.5128689768248313) < 1.0e-3 and abs(tII.max() - 1.6415759440014273) < 1.0e-3, (tII.min(), tII.max())
    assert abs(txx_max[-1] - 1.6376258215356436) < 1.0e-4, txx_max[-1]
    assert float(to_host(st.EII_pl).max()) > 0


def test_convection3d_stokes_solve(oracle):
    """config 5's Stokes half at 16³: three phases (plastic crust, blob, weak layer), PT_Density with args.P aliasing stokes.P,
    gravity — solve to tolerance, compare with the oracle"""
    from justrelax_jl_b200 import B200Backend, PhaseRatios, rheology as R, setups, stokes as jst

    s = setups.convection3d(16, 16, 16)
    d = oracle.alloc_stokes(s.ni, s.fields)
    d["Pargs"] = d["P"]
    st, extra = device_stokes(s.ni, d)
    ratios = {k: v for k, v in s.ratios.items() if k in ("center", "xy", "yz", "xz")}
    vc = oracle.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), ratios)
    kw = dict(s.kwargs, iterMax=3000, nout=500)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=kw["iterMax"], nout=kw["nout"],
                            viscosity_cutoff=kw["viscosity_cutoff"])
    ref = oracle.solve3d_VC(d, s.ni, opts, vc)
    pr = PhaseRatios.from_arrays(B200Backend, **ratios)
    args = dict(T=extra["T"], P=st.P)
    ρg = (extra["rhogx"], extra["rhogy"], extra["rhogz"])
    out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, args, s.dt, s.igg, kwargs=kw)
    assert out.iter == ref["iter"]
    assert np.allclose(out.err_evo1, ref["err_evo1"], rtol=1e-6)
    assert np.abs(d["Vz"]).max() > 0
    got = {**st.slots(), "rhogz": ρg[2]}
    compare_slots(got, d, ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "rhogz", "eta"], 1.0e-8, "convection Stokes fields")


def test_standalone_3d_kernels(oracle):
    from justrelax_jl_b200 import B200Backend, PhaseRatios, PTArray, StokesArrays, rheology as R, setups, stokes as jst, to_host

    s = setups.random_vc3d((11, 9, 8), seed=4)
    d = oracle.alloc_stokes(s.ni, s.fields)
    st, extra = device_stokes(s.ni, d)
    ratios = s.ratios
    vc = oracle.vc_inputs(R.lower_stokes(s.rheology), R.gravity_of(s.rheology), ratios)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, dict(free_slip=[1] * 6), s.ni, iterMax=1, nout=1, viscosity_cutoff=(1e-2, 0.5))
    fs = oracle.make_fields(d, s.ni)
    pr = PhaseRatios.from_arrays(B200Backend, **ratios)
    args = dict(T=extra["T"], P=extra["Pargs"])
    oracle.lib().orc_viscosity3d(C.byref(fs), C.byref(opts), C.byref(vc), C.c_double(0.25))
    jst.compute_viscosity_(st, pr, args, s.rheology, (1e-2, 0.5), relaxation=0.25)
    assert max_rel_diff(to_host(st.viscosity.η), d["eta"]) <= 1e-15
    oracle.lib().orc_rhog3d(C.byref(fs), C.byref(vc))
    ρg = (extra["rhogx"], extra["rhogy"], extra["rhogz"])
    jst.compute_ρg_(ρg, pr, s.rheology, args, st)
    for a, nm in zip(ρg, ("rhogx", "rhogy", "rhogz")):
        assert max_rel_diff(to_host(a), d[nm]) <= 1e-15, nm


def test_compute_dt_and_accumulate_entry_points(oracle):
    """compute_dt (src/Utils.jl:492-519), accumulate_tensor! / accumulate_vol! (src/stokes/StressKernels.jl:364-438), flow_bcs! and
    displacement2velocity! with DisplacementBoundaryConditions (src/types/displacement.jl:33-70) through the public API"""
    from justrelax_jl_b200 import B200Backend, PTArray, StokesArrays, stokes as jst, to_host
    from justrelax_jl_b200.types import DisplacementBoundaryConditions

    rng = np.random.default_rng(12)
    ni = (10, 9, 8)
    st = StokesArrays(B200Backend, *ni)
    host = {}
    for nm, a in (("Vx", st.V.Vx), ("Vy", st.V.Vy), ("Vz", st.V.Vz), ("Ux", st.U.Ux), ("Uy", st.U.Uy), ("Uz", st.U.Uz)):
        host[nm] = np.asfortranarray(rng.uniform(-2, 2, size=tuple(a.shape)))
        a.copy_(PTArray(B200Backend)(host[nm]))
    di = (0.1, 0.2, 0.3)
    want = min(0.5, 0.9 * min(d * (1.0 / np.abs(host[k]).max()) for d, k in zip(di, ("Vx", "Vy", "Vz"))))
    assert jst.compute_dt_(st, di, 0.5) == want
    assert jst.compute_dt_(st, di) == 0.9 * min(d * (1.0 / np.abs(host[k]).max()) for d, k in zip(di, ("Vx", "Vy", "Vz")))
    # accumulate_tensor! 3D: II += second_invariant_staggered(ε_pl) dt
    T = st.ε_pl
    h = {}
    for nm in ("xx", "yy", "zz", "yz", "xz", "xy"):
        a = getattr(T, nm)
        h[nm] = np.asfortranarray(rng.uniform(-1, 1, size=tuple(a.shape)))
        a.copy_(PTArray(B200Backend)(h[nm]))
    E0 = np.asfortranarray(rng.uniform(0, 1, size=ni))
    st.EII_pl.copy_(PTArray(B200Backend)(E0))
    jst.accumulate_tensor_(st.EII_pl, T, 0.37, ni)
    g = lambda A, ax: sum(np.take(A, range(o[0], o[0] + A.shape[ax[0]] - 1), ax[0]).take(range(o[1], o[1] + A.shape[ax[1]] - 1), ax[1]) ** 2
                          for o in ((0, 0), (1, 0), (0, 1), (1, 1))) / 4
    II = np.sqrt(0.5 * (h["xx"] ** 2 + h["yy"] ** 2 + h["zz"] ** 2) + g(h["yz"], (1, 2)) + g(h["xz"], (0, 2)) + g(h["xy"], (0, 1)))
    assert max_rel_diff(to_host(st.EII_pl), E0 + II * 0.37) <= 1e-14
    ev = np.asfortranarray(rng.uniform(-1, 1, size=ni))
    st.ε_vol_pl.copy_(PTArray(B200Backend)(ev))
    jst.accumulate_vol_(st.EVol_pl, st.ε_vol_pl, 0.37)
    assert max_rel_diff(to_host(st.EVol_pl), 0.37 * ev) <= 1e-15
    # DisplacementBoundaryConditions: flow_bcs! acts on U; displacement2velocity!: V = U · inv(dt)
    bcs = DisplacementBoundaryConditions(free_slip=dict(left=True, right=True, front=True, back=True, top=True, bot=True))
    jst.flow_bcs_(st, bcs)
    d = oracle.alloc_stokes(ni, dict(Vx=host["Ux"], Vy=host["Uy"], Vz=host["Uz"]))
    opts = oracle.make_opts(type("pt", (), dict(r=1, θ_dτ=1, ηdτ=1, ϵ_rel=1, ϵ_abs=1)), (1, 1, 1), 1.0, dict(free_slip=[1] * 6), ni, iterMax=1, nout=1)
    fs = oracle.make_fields(d, ni)
    oracle.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)
    for u, v in (("Ux", "Vx"), ("Uy", "Vy"), ("Uz", "Vz")):
        assert np.array_equal(to_host(getattr(st.U, u)), d[v]), u
    assert np.array_equal(to_host(st.V.Vx), host["Vx"])
    jst.displacement2velocity_(st, 0.25)
    assert np.array_equal(to_host(st.V.Vz), d["Vz"] * (1.0 / 0.25))
