"""init_global_grid(...; periodx, periody, periodz) on ONE rank: the rank is its own neighbour in a periodic dimension
(ImplicitGlobalGrid; the reference uses it in test/test_periodic_boundary_conditions_MPI.jl:12-19 and
miniapps/DYREL2D/shear_band/ShearBand2D_DYREL_SimpleShearPeriodic.jl:214).  update_halo_, the 3D-VA iteration (fused and
unfused), the thermal iteration and the 3D multiphase iteration on such a grid against the oracle + the literal exchange of
tests/mrank.py.  The multi-rank version of the same checks runs in tests/mgpu_worker.py (section 8)."""
import numpy as np
import pytest

import mrank
from util import device_stokes, max_rel_diff

pytestmark = pytest.mark.gpu

NAMES = ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "Rx", "Ry", "Rz", "RP", "etatau"]
DIMS = (1, 1, 1)


@pytest.fixture()
def periodic_grid():
    from justrelax_jl_b200 import comm

    made = []

    def make(ni, periods):
        igg = comm.init_global_grid(*ni, init_dist=False, periodx=periods[0], periody=periods[1], periodz=periods[2])
        made.append(igg)
        return igg

    yield make
    comm.finalize_global_grid()


@pytest.mark.parametrize("periods", [(1, 0, 0), (1, 0, 1), (1, 1, 1), (0, 1, 0)])
def test_update_halo_on_a_periodic_single_rank(periodic_grid, periods):
    from justrelax_jl_b200 import B200Backend, PTArray, comm, to_host

    ni = (12, 9, 10)
    igg = periodic_grid(ni, periods)
    assert tuple(igg.n_g(ni)) == mrank.n_g(ni, DIMS, periods)
    for grow in [(0, 0, 0), (1, 2, 2), (2, 1, 2), (2, 2, 1), (2, 2, 2), (1, 1, 0)]:
        ext = tuple(ni[d] + grow[d] for d in range(3))
        host = [np.asfortranarray(np.random.default_rng(3).uniform(size=ext))]
        A = PTArray(B200Backend)(host[0])
        for _ in range(2):
            comm.update_halo_(A)
            mrank.update_halo(host, DIMS, ni, periods)
        assert np.array_equal(to_host(A), host[0]), (periods, grow)


@pytest.mark.parametrize("unfused", [True, False], ids=["unfused", "fused"])
@pytest.mark.parametrize("dt,finite_K", [(0.7, True), (np.inf, False)])
def test_va_iterations_on_a_periodic_single_rank(oracle, periodic_grid, dt, finite_K, unfused):
    from justrelax_jl_b200 import _abi, setups, stokes as jst, to_host
    from justrelax_jl_b200.types import VelocityBoundaryConditions

    ni, periods = (20, 17, 15), (1, 0, 1)
    igg = periodic_grid(ni, periods)
    s = setups.random_stokes3d(ni, seed=77, dt=dt, finite_K=finite_K)
    blocks = [oracle.alloc_stokes(ni, s.fields)]
    flags = dict(free_slip=[0, 0, 1, 1, 0, 0], no_slip=[0] * 6, periodic=[0] * 6)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, dt, flags, mrank.n_g(ni, DIMS, periods), iterMax=100, nout=100)
    st, extra = device_stokes(ni, blocks[0])
    mrank.va_pre(oracle, blocks, DIMS, ni, periods)
    mrank.va_iterate(oracle, blocks, opts, DIMS, ni, 5, periods)
    bcs = VelocityBoundaryConditions(free_slip=dict(left=False, right=False, front=True, back=True, top=False, bot=False),
                                     no_slip=dict(left=False, right=False, front=False, back=False, top=False, bot=False))
    jst.set_flags(_abi.JR_FLAG_UNFUSED if unfused else 0)
    try:
        jst.iterate_(st, s.pt_stokes, s.grid, bcs, (extra["rhogx"], extra["rhogy"], extra["rhogz"]), extra["K"], extra["G"], dt, 5, igg)
    finally:
        jst.set_flags(0)
    worst = {nm: max_rel_diff(to_host(st.slots()[nm]), blocks[0][nm]) for nm in NAMES}
    assert max(worst.values()) <= 1e-12, worst


def test_thermal_iterations_on_a_periodic_single_rank(oracle, periodic_grid):
    from justrelax_jl_b200 import thermal as jth
    from justrelax_jl_b200.types import Geometry, TemperatureBoundaryConditions
    import test_gpu_thermal as tg

    nit, periods = (13, 12, 10), (1, 0, 0)
    igg = periodic_grid(nit, periods)
    li = (1.0e5, 1.1e5, 1.2e5)
    grid = Geometry(nit, li)
    bc = TemperatureBoundaryConditions(no_flux=dict(left=False, right=False, front=True, back=True, top=True, bot=True))
    blocks = [oracle.alloc_thermal(nit, tg.random_thermal(nit, 41, 3))]
    pt = type("PT", (), {})()
    pt.ϵ, pt.max_lxyz, pt.Vpdτ = 1e-8, max(li), min(grid.di.center) * 0.5
    ot = oracle.thermal_opts(_di=grid._di.center, dt=1.0e11, eps=1e-8, iterMax=10, nout=4, max_lxyz=pt.max_lxyz, Vpdtau=pt.Vpdτ, form=1,
                             phases=tg.PHASES, bc=bc)
    th, extra = tg.to_device(nit, blocks[0])
    mrank.thermal_iterate(oracle, blocks, ot, DIMS, nit, 4, periods)
    pt.θr_dτ, pt.dτ_ρ = extra["theta_r_dtau"], extra["dtau_rho"]
    ph = tg._Phase()
    ph.center, ph.Vx, ph.Vy, ph.Vz = extra["phase_c"], extra["phase_x"], extra["phase_y"], extra["phase_z"]
    jth.thermal_iterate_(th, pt, bc, tg.rheology_of(tg.PHASES), dict(P=extra["P"], T=th.T), 1.0e11, grid, 4,
                         kwargs=dict(verbose=False, phase=ph, igg=igg))
    tg.compare(th, blocks[0], ["T", "qTx", "qTy", "qTz", "qTx2", "qTy2", "qTz2", "ResT"], "thermal 3D on a periodic single rank")


def test_vc_iterations_on_a_periodic_single_rank(oracle, periodic_grid):
    from justrelax_jl_b200 import B200Backend, PhaseRatios, rheology as R, setups, to_host
    from justrelax_jl_b200.stokes3d_vc import iterate3d_VC_
    from justrelax_jl_b200.types import VelocityBoundaryConditions

    niv, periods = (14, 12, 11), (0, 1, 1)
    igg = periodic_grid(niv, periods)
    sv = setups.random_vc3d(niv, seed=901)
    d = oracle.alloc_stokes(niv, sv.fields)
    d["Pargs"] = d["P"]
    blocks = [d]
    vcs = [oracle.vc_inputs(R.lower_stokes(sv.rheology), R.gravity_of(sv.rheology), sv.ratios)]
    flags = dict(free_slip=[1, 1, 0, 0, 0, 0], no_slip=[0] * 6, periodic=[0] * 6)
    opts = oracle.make_opts(sv.pt_stokes, sv.grid._di.center, sv.dt, flags, mrank.n_g(niv, DIMS, periods), iterMax=100, nout=100,
                            viscosity_relaxation=0.3, viscosity_cutoff=sv.kwargs["viscosity_cutoff"])
    st, extra = device_stokes(niv, blocks[0])
    mrank.vc_iterate(oracle, blocks, opts, vcs, DIMS, niv, 4, finish=True, periods=periods)
    pr = PhaseRatios.from_arrays(B200Backend, **sv.ratios)
    bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, front=False, back=False, top=False, bot=False),
                                     no_slip=dict(left=False, right=False, front=False, back=False, top=False, bot=False))
    iterate3d_VC_(st, sv.pt_stokes, sv.grid, bcs, (extra["rhogx"], extra["rhogy"], extra["rhogz"]), pr, sv.rheology, dict(T=extra["T"], P=st.P),
                  sv.dt, 4, igg, finish=True, kwargs=dict(viscosity_relaxation=0.3, viscosity_cutoff=sv.kwargs["viscosity_cutoff"]))
    names = ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "tyz_c", "txz_c", "txy_c", "eta", "etatau", "lam", "Rx", "Ry", "Rz", "RP"]
    worst = {nm: max_rel_diff(to_host(st.slots()[nm]), blocks[0][nm]) for nm in names}
    assert max(worst.values()) <= 1e-12, worst
