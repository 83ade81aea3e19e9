"""GPU parity tests of heatdiffusion_PT! (through the C ABI) against the CPU oracle (oracle/thermal.c).

Tolerance: per-field max relative difference <= 1e-12 after a fixed number of PT iterations (north star); ghost
edges/corners of T are excluded — the reference's values there depend on its thread order and nothing reads them
(DESIGN.md §3.6).  Config 1 (test/test_diffusion2D.jl) additionally checks the reference's golden temperatures."""
import ctypes as C

import numpy as np
import pytest

from util import max_rel_diff

pytestmark = pytest.mark.gpu
TOL = 1.0e-12


def face_only(T):
    """copy of ghosted T with ghost edges/corners (>= 2 ghost indices) zeroed"""
    T = np.array(T, order="F", copy=True)
    g = [np.zeros(s, dtype=int) for s in T.shape]
    for a in g:
        a[0] = a[-1] = 1
    cnt = sum(np.reshape(a, [-1 if q == d else 1 for q in range(T.ndim)]) for d, a in enumerate(g))
    T[cnt >= 2] = 0.0
    return T


def to_device(ni, host):
    from justrelax_jl_b200 import B200Backend, PTArray, ThermalArrays

    th = ThermalArrays(B200Backend, *ni)
    names = dict(T="T", Told="Told", dT="ΔT", qTx="qTx", qTy="qTy", qTz="qTz", qTx2="qTx2", qTy2="qTy2", qTz2="qTz2", H="H",
                 shear_heating="shear_heating", adiabatic="adiabatic", ResT="ResT")
    for k, attr in names.items():
        if k in host and getattr(th, attr) is not None:
            getattr(th, attr).copy_(PTArray(B200Backend)(host[k]))
    extra = {k: PTArray(B200Backend)(v) for k, v in host.items() if k not in names}
    return th, extra


def compare(th, host, names, label):
    from justrelax_jl_b200 import to_host

    attr = dict(dT="ΔT")
    bad = {}
    for nm in names:
        a, b = to_host(getattr(th, attr.get(nm, nm))), host[nm]
        if nm in ("T", "Told", "dT"):
            a, b = face_only(a), face_only(b)
        r = max_rel_diff(a, b)
        if not r <= TOL:
            bad[nm] = r
    assert not bad, f"{label}: {bad}"


def random_thermal(ni, seed, nphase=0):
    rng = np.random.default_rng(seed)
    g = tuple(n + 2 for n in ni)
    U = lambda lo, hi, s: np.asfortranarray(rng.uniform(lo, hi, size=s))
    f = dict(T=U(1000, 2000, g), Told=U(1000, 2000, g), H=U(0, 1e-6, ni), shear_heating=U(0, 1e-7, ni), adiabatic=U(-1e-9, 1e-9, ni),
             theta_r_dtau=U(0.5, 3, ni), dtau_rho=U(1e-3, 1e-2, ni), K=U(2, 4, ni), rhoCp=U(3e6, 4e6, ni), P=U(0, 1e9, ni))
    for a, nm in enumerate(("qTx", "qTy", "qTz")[:len(ni)]):
        e = tuple(n + (1 if b == a else 0) for b, n in enumerate(ni))
        f[nm] = U(-1e-2, 1e-2, e)
    if nphase:
        def ratios(shape):
            r = rng.dirichlet(np.ones(nphase), size=shape)            # (..., nphase)
            pick = rng.uniform(size=shape)
            one_hot = np.eye(nphase)[rng.integers(0, nphase, size=shape)]
            r = np.where((pick < 0.4)[..., None], one_hot, r)          # many pure cells: exact 0 and 1 ratios
            return np.asfortranarray(np.moveaxis(r, -1, 0).reshape(nphase, -1).T.reshape(*shape, nphase))
        f["phase_c"] = ratios(ni)
        for a, nm in enumerate(("phase_x", "phase_y", "phase_z")[:len(ni)]):
            f[nm] = ratios(tuple(n + (1 if b == a else 0) for b, n in enumerate(ni)))
    return f


PHASES = [dict(rho_kind=1, has_Hr=1, rho0=3.1e3, alpha=1.5e-5, beta=1e-11, T0=273.0, P0=1e5, Cp=1.2e3, k=3.0, Hr=2e-7),
          dict(rho_kind=0, has_Hr=0, rho0=2.7e3, alpha=0.0, beta=0.0, T0=0.0, P0=0.0, Cp=1.0e3, k=2.2, Hr=0.0),
          dict(rho_kind=2, has_Hr=1, rho0=3.3e3, alpha=3e-5, beta=0.0, T0=300.0, P0=0.0, Cp=1.1e3, k=4.1, Hr=5e-8)]


# the same table with GeoParams' TP_Conductivity k(T, P) = (a + b / (T + c)) (1 + d P) on two of the three phases (the parameters of
# miniapps/convection/Particles3D/Layered_rheology.jl:45-57; d per Pa)
PHASES_TP = [dict(PHASES[0], k=0.0, k_kind=1, k_a=0.64, k_b=807.0, k_c=0.77, k_d=0.00004e-6),
             dict(PHASES[1]),
             dict(PHASES[2], k=0.0, k_kind=1, k_a=0.73, k_b=1293.0, k_c=0.77, k_d=0.00004e-6)]


def rheology_of(rows):
    from justrelax_jl_b200 import rheology as R

    out = []
    for i, r in enumerate(rows):
        ρ = (R.ConstantDensity(ρ=r["rho0"]) if r["rho_kind"] == 0 else
             R.PT_Density(ρ0=r["rho0"], α=r["alpha"], β=r["beta"], T0=r["T0"], P0=r["P0"]) if r["rho_kind"] == 1 else
             R.T_Density(ρ0=r["rho0"], α=r["alpha"], T0=r["T0"]))
        out.append(R.SetMaterialParams(Phase=i + 1, Density=ρ, HeatCapacity=R.ConstantHeatCapacity(Cp=r["Cp"]),
                                       Conductivity=(R.TP_Conductivity(a=r["k_a"], b=r["k_b"], c=r["k_c"], d=r["k_d"]) if r.get("k_kind", 0) == 1
                                                     else R.ConstantConductivity(k=r["k"])),
                                       RadioactiveHeat=R.ConstantRadioactiveHeat(H_r=r["Hr"]) if r["has_Hr"] else None))
    return tuple(out)


def bc_variants(nd):
    from justrelax_jl_b200.types import TemperatureBoundaryConditions as TBC

    if nd == 2:
        return [TBC(no_flux=dict(left=True, right=True, top=False, bot=False), constant_value=dict(left=True, right=True, top=300.0, bot=3500.0)),
                TBC(no_flux=dict(left=False, right=False, top=True, bot=False), periodic=dict(left=True, right=True, top=False, bot=False),
                    constant_flux=dict(left=False, right=False, top=False, bot=0.03))]
    return [TBC(no_flux=dict(left=True, right=True, front=True, back=True, top=False, bot=False),
                constant_value=dict(left=False, right=False, front=False, back=False, top=300.0, bot=1800.0)),
            TBC(no_flux=dict(left=False, right=False, front=True, back=False, top=False, bot=False),
                periodic=dict(left=True, right=True, front=False, back=False, top=False, bot=False),
                constant_value=dict(left=False, right=False, front=False, back=500.0, top=False, bot=False),
                constant_flux=dict(left=False, right=False, front=False, back=False, top=0.02, bot=-0.01))]


class _Phase:
    pass


def _run_case(oracle, ni, form, nphase, vb, bc, niter, seed_shift=0, table=None):
    """`niter` PT iterations (the last one sampled) from a random state on the B200 and in the oracle; returns the result"""
    from justrelax_jl_b200 import thermal as jth
    from justrelax_jl_b200.types import Geometry

    li = tuple(1.0e5 * (1 + 0.1 * d) for d in range(len(ni)))
    grid = Geometry(ni, li)
    rows = (table or PHASES)[:max(nphase, 1)]
    host = random_thermal(ni, 77 + ni[0] + vb + seed_shift, nphase if nphase > 1 else 0)
    if vb == 1:  # Dirichlet mask on a block of nodes
        m = np.zeros(tuple(n + 2 for n in ni), order="F")
        m[(slice(2, 5),) * len(ni)] = 1.0
        host["dir_mask"] = m
        bc.dirichlet = (1234.5, m)
    full = oracle.alloc_thermal(ni, host)
    pt = type("PT", (), {})()
    pt.ϵ, pt.max_lxyz, pt.Vpdτ = 1e-8, max(li), min(grid.di.center) * 0.5
    dt = 1.0e11
    o = oracle.thermal_opts(_di=grid._di.center, dt=dt, eps=1e-8, iterMax=10, nout=niter, max_lxyz=pt.max_lxyz, Vpdtau=pt.Vpdτ,
                            form=form, phases=rows, bc=bc, dir_const=1234.5)
    th, extra = to_device(ni, full)
    fs = oracle.thermal_fields(full, ni)
    for _ in range(niter):
        oracle.lib().orc_thermal_iterate_once(C.byref(fs), C.byref(o))
    oracle.lib().orc_thermal_check_res(C.byref(fs), C.byref(o))
    pt.θr_dτ, pt.dτ_ρ = extra["theta_r_dtau"], extra["dtau_rho"]
    kw = dict(verbose=False)
    if nphase > 1:
        ph = _Phase()
        ph.center, ph.Vx, ph.Vy = extra["phase_c"], extra["phase_x"], extra["phase_y"]
        ph.Vz = extra.get("phase_z")
        kw["phase"] = ph
    if form == 0:
        r = jth.thermal_iterate_(th, pt, bc, extra["K"], extra["rhoCp"], dt, grid, niter, kwargs=kw)
    else:
        rheo = rheology_of(rows)
        r = jth.thermal_iterate_(th, pt, bc, rheo if nphase > 1 else rheo[0], dict(P=extra["P"], T=th.T), dt, grid, niter, kwargs=kw)
    assert r.kernel_launches > 0
    names = ["T", "qTx", "qTy", "qTx2", "qTy2", "ResT"] + (["qTz", "qTz2"] if len(ni) == 3 else [])
    compare(th, full, names, f"ni={ni} form={form} nphase={nphase} bc={vb} niter={niter}")
    from justrelax_jl_b200 import to_host
    if nphase > 1:
        assert max_rel_diff(to_host(pt.θr_dτ), full["theta_r_dtau"]) <= TOL and max_rel_diff(to_host(pt.dτ_ρ), full["dtau_rho"]) <= TOL
    bc.dirichlet = None
    return r


@pytest.mark.parametrize("ni", [(17, 12), (33, 40), (9, 8, 7), (34, 17, 21)])
@pytest.mark.parametrize("form,nphase", [(0, 0), (1, 1), (1, 3)])
def test_fixed_iterations_random_state(oracle, ni, form, nphase):
    for vb, bc in enumerate(bc_variants(len(ni))):
        for niter in (1, 3):
            _run_case(oracle, ni, form, nphase, vb, bc, niter)


@pytest.mark.parametrize("ni", [(17, 12), (33, 40), (9, 8, 7), (34, 17, 21)])
@pytest.mark.parametrize("nphase", [1, 3])
def test_tp_conductivity_fixed_iterations(oracle, ni, nphase):
    """TP_Conductivity: K̄ at the faces from the mean of the two adjacent temperatures and the pressure of the cell on either
    side, every iteration; the PT coefficients from T, P at the centres — single MaterialParams and three phases (two TP, one constant)"""
    for vb, bc in enumerate(bc_variants(len(ni))):
        for niter in (1, 3, 4):
            _run_case(oracle, ni, 1, nphase, vb, bc, niter, table=PHASES_TP)


@pytest.mark.parametrize("ni", [(9, 8, 7), (34, 17, 21), (40, 70, 37)])
@pytest.mark.parametrize("form,nphase", [(0, 0), (1, 1), (1, 3)])
@pytest.mark.parametrize("kchunk", [32, 5])
def test_fused_flux_update_3d(oracle, monkeypatch, ni, form, nphase, kchunk):
    """3D, no Dirichlet mask / constant-flux face: pairs of unsampled iterations run the fused flux + update kernel on ping-pong
    sets (z-marching, chunk seams at kchunk): same fields as the oracle, fewer launches than the two-kernel path"""
    monkeypatch.setenv("JRB200_TH_KCHUNK", str(kchunk))
    bc = bc_variants(3)[0]
    launches = {}
    for fused in ("1", "0"):
        monkeypatch.setenv("JRB200_TH_FUSED", fused)
        for niter in (4, 7, 8):
            r = _run_case(oracle, ni, form, nphase, 0, bc, niter, seed_shift=niter)
        launches[fused] = r.kernel_launches
    assert launches["1"] < launches["0"], launches


def test_config1_diffusion2d_golden_and_parity(oracle):
    """BASELINE config 1: test/test_diffusion2D.jl — 20 × 50 kyr, rheology form, single MaterialParams."""
    from justrelax_jl_b200 import B200Backend, PTArray, setups, thermal as jth, to_host
    from test_oracle_thermal import run_diffusion2d

    s = setups.diffusion2d()
    f = oracle.alloc_thermal(s.ni, dict(T=s.T, H=s.H, P=s.P, theta_r_dtau=s.pt.θr_dτ, dtau_rho=s.pt.dτ_ρ))
    th, extra = to_device(s.ni, f)
    outs = run_diffusion2d(oracle, s, f)
    # device: same script as the reference test
    jth.thermal_bcs_(th, s.bc)
    Tin = th.T[1:-1, 1:-1]
    Tin += PTArray(B200Backend)(s.perturbation.astype(np.float64) * s.δT)
    pt = jth.PTThermalCoeffs(B200Backend, extra.get("K", PTArray(B200Backend)(s.K)), PTArray(B200Backend)(s.ρCp), s.dt, s.di, s.li,
                             CFL=0.95 / np.sqrt(2.1))
    assert max_rel_diff(to_host(pt.θr_dτ), s.pt.θr_dτ) <= 1e-14
    rheo = rheology_of(s.phases)[0]
    iters = []
    for _ in range(s.nt):
        out = jth.heatdiffusion_PT_(th, pt, s.bc, rheo, dict(P=extra["P"], T=th.T), s.dt, s.grid, kwargs=dict(verbose=False))
        iters.append(int(out.iter_count[-1]))
    T = to_host(th.T)
    assert abs(T[17, 17] - 1817.9448461176817) < 1.0e-1 and abs(T[16, 16] - 1827.4674313638786) < 1.0e-1
    assert iters == [int(o["iter_count"][-1]) for o in outs]
    compare(th, f, ["T", "Told", "dT", "qTx", "qTy", "ResT"], "config 1 after 20 steps")


def test_diffusion3d_reference_golden_to_the_last_digit(oracle):
    """test/test_diffusion3D.jl (assertions commented out in the reference, golden numbers kept): 10 × 50 kyr at 32³, rheology form, single
    MaterialParams, through the public API — the solve loop runs the fused flux + update kernel in pairs between the samples.  The CPU
    restatement reproduces the Julia golden to the last digit (tests/test_oracle_thermal.py); the B200 must land on the same numbers
    (≤ 1e-12 relative against the golden itself), with the oracle's iteration counts and fields."""
    from justrelax_jl_b200 import B200Backend, PTArray, setups, thermal as jth, to_host

    s = setups.diffusion3d()
    init = dict(T=s.T.copy(order="F"), H=s.H, P=s.P, theta_r_dtau=s.pt.θr_dτ, dtau_rho=s.pt.dτ_ρ)
    f = oracle.alloc_thermal(s.ni, init)
    th, extra = to_device(s.ni, oracle.alloc_thermal(s.ni, init))
    o = oracle.thermal_opts(_di=s.grid._di.center, dt=s.dt, eps=s.pt.ϵ, iterMax=s.kwargs["iterMax"], nout=s.kwargs["nout"],
                            max_lxyz=s.pt.max_lxyz, Vpdtau=s.pt.Vpdτ, form=1, phases=s.phases, bc=s.bc)
    f["T"][1:-1, 1:-1, 1:-1][s.perturbation] += s.δT
    outs = [oracle.heatdiffusion_PT(f, s.ni, o) for _ in range(s.nt)]
    Tin = th.T[1:-1, 1:-1, 1:-1]
    Tin += PTArray(B200Backend)(s.perturbation.astype(np.float64) * s.δT)
    pt = type("PT", (), {})()
    pt.ϵ, pt.max_lxyz, pt.Vpdτ = s.pt.ϵ, s.pt.max_lxyz, s.pt.Vpdτ
    pt.θr_dτ, pt.dτ_ρ = extra["theta_r_dtau"], extra["dtau_rho"]
    rheo = rheology_of(s.phases)[0]
    iters = []
    for _ in range(s.nt):
        out = jth.heatdiffusion_PT_(th, pt, s.bc, rheo, dict(P=extra["P"], T=th.T), s.dt, s.grid, kwargs=dict(s.kwargs))
        iters.append(int(out.iter_count[-1]))
    assert iters == [int(o_["iter_count"][-1]) for o_ in outs]
    T = to_host(th.T)
    assert abs(T[15, 15, 15] / 1813.2470160788096 - 1) < 1.0e-12, repr(T[15, 15, 15])
    assert abs(T[1:-1, 1:-1, 1:-1][15, 15, 15] / 1831.2568044653274 - 1) < 1.0e-12
    compare(th, f, ["T", "Told", "dT", "qTx", "qTy", "qTz", "ResT"], "3D single-phase diffusion after 10 steps")


def test_diffusion3d_multiphase_solve_parity(oracle):
    """test/test_diffusion3D_multiphase.jl through the public API (heatdiffusion_PT!, rheology form, two phases with ratios, nout = 100):
    the solve loop runs the fused flux + update kernel in pairs between the samples.  Same PT iteration counts as the oracle and the same
    fields (≤ 1e-12; the arithmetic is the oracle's), and the reference's golden temperatures (rtol 1e-3) after all 10 steps."""
    from justrelax_jl_b200 import B200Backend, PTArray, setups, thermal as jth, to_host
    from test_oracle_thermal import run_diffusion_multiphase

    s = setups.diffusion_multiphase(3)
    f, outs = run_diffusion_multiphase(oracle, s)
    init = dict(T=s.T.copy(order="F"), H=s.H, P=s.P, theta_r_dtau=s.pt.θr_dτ, dtau_rho=s.pt.dτ_ρ)
    th, extra = to_device(s.ni, oracle.alloc_thermal(s.ni, init))
    Tin = th.T[1:-1, 1:-1, 1:-1]
    Tin += PTArray(B200Backend)(s.perturbation.astype(np.float64) * s.δT)
    pt = type("PT", (), {})()
    pt.ϵ, pt.max_lxyz, pt.Vpdτ = s.pt.ϵ, s.pt.max_lxyz, s.pt.Vpdτ
    pt.θr_dτ, pt.dτ_ρ = extra["theta_r_dtau"], extra["dtau_rho"]
    ph = _Phase()
    ph.center, ph.Vx, ph.Vy, ph.Vz = (PTArray(B200Backend)(s.phase[k]) for k in ("center", "Vx", "Vy", "Vz"))
    rheo = rheology_of(s.phases)
    iters = []
    for _ in range(s.nt):
        out = jth.heatdiffusion_PT_(th, pt, s.bc, rheo, dict(P=extra["P"], T=th.T), s.dt, s.grid, kwargs=dict(s.kwargs, phase=ph))
        iters.append(int(out.iter_count[-1]))
    assert iters == [int(o["iter_count"][-1]) for o in outs]
    T = to_host(th.T)
    assert abs(T[15, 15, 15] / 1816.8262937737384 - 1) < 1.0e-3
    assert abs(T[1:-1, 1:-1, 1:-1][15, 15, 15] / 1834.4197141500213 - 1) < 1.0e-3
    compare(th, f, ["T", "Told", "dT", "qTx", "qTy", "qTz", "ResT"], "3D multiphase diffusion after 10 steps")


@pytest.mark.parametrize("ni", [(12, 9), (10, 9, 8)])
def test_thermal_bcs_standalone(oracle, ni):
    from justrelax_jl_b200 import B200Backend, PTArray, thermal as jth, to_host

    for vb, bc in enumerate(bc_variants(len(ni))):
        host = random_thermal(ni, 5 + vb)
        full = oracle.alloc_thermal(ni, host)
        o = oracle.thermal_opts(_di=(1,) * len(ni), dt=1, eps=1e-8, iterMax=1, nout=1, max_lxyz=1, Vpdtau=1, form=0, bc=bc)
        fs = oracle.thermal_fields(full, ni)
        T = PTArray(B200Backend)(host["T"])
        oracle.lib().orc_thermal_bcs(C.byref(fs), C.byref(o))
        jth.thermal_bcs_(T, bc)
        assert np.array_equal(face_only(to_host(T)), face_only(full["T"])), (ni, vb)


def test_unsupported_law_fails_loudly():
    from justrelax_jl_b200 import rheology as R

    class T_Conductivity_Whittington:
        pass

    p = R.SetMaterialParams(Density=R.ConstantDensity(), HeatCapacity=R.ConstantHeatCapacity(), Conductivity=T_Conductivity_Whittington())
    with pytest.raises(R.UnsupportedRheology):
        R.lower_thermal(p)
