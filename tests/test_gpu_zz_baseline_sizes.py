"""GPU parity at the BASELINE.json sizes (configs 2-5): the CUDA path through the C ABI against the CPU oracle on the same
inputs, per-field max relative difference <= 1e-12 after a fixed number of PT iterations — with the plans / tilings the
benchmarks actually run (255^3: BY = 10, one z-chunk, 288 persistent CTAs; 257^3: remainder-column packing; 1023^2: the
per-grid tile height), which the small-grid tests do not reach.

Sizes: 3D-VA 255^3 (config 4), 2D-V2 511^2 (config 2), 2D-VC 1023^2 (config 3), 3D-VC 257^3 and thermal 257^3 with three
phases (config 5, per GPU).  The oracle needs a few seconds per case on the GPU box's host cores.
"""
import numpy as np
import pytest

from util import compare_slots

pytestmark = pytest.mark.gpu
TOL = 1.0e-12
FS6 = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)


def test_va3d_255_solvi_default_plan(oracle):
    """config 4: SolVi 255^3, dt = Inf, 5 iterations with the default plan of the benchmark"""
    import test_gpu_stokes3d as t3
    from justrelax_jl_b200 import setups, stokes as jst

    s = setups.solvi3d(255, 255, 255)
    st, d = t3._run_both(oracle, s, 5, FS6, False)
    info = jst.plan_info()
    assert (info["BY"], info["nchunk"], info["rhog_const"]) == (10, 1, 1), info
    compare_slots(st.slots(), d, t3.FIELDS_STATE + t3.FIELDS_DIAG, TOL, "SolVi3D 255^3 x5")


def test_va3d_255_random_finite_dt(oracle):
    """the general visco-elastic form (finite dt, K, G, P0, τ_o streamed: 264 B/cell) at 255^3 from a seeded random state"""
    import test_gpu_stokes3d as t3
    from justrelax_jl_b200 import setups

    s = setups.random_stokes3d((255, 255, 255), seed=4, dt=0.7, finite_K=True)
    st, d = t3._run_both(oracle, s, 2, FS6, False)
    compare_slots(st.slots(), d, t3.FIELDS_STATE + t3.FIELDS_DIAG, TOL, "random 3D-VA 255^3 x2")


def test_v2_511_solcx(oracle):
    """config 2: SolCx 511^2, 50 iterations"""
    import test_gpu_stokes2d as t2
    from justrelax_jl_b200 import setups, stokes as jst
    from util import bc_flags, device_stokes

    s = setups.solcx2d(511, 511)
    niter = 50
    d = oracle.alloc_stokes(s.ni, s.fields)
    st, extra = device_stokes(s.ni, d)
    flags = bc_flags(s.flow_bcs)
    opts = oracle.make_opts(s.pt_stokes, s.grid._di.center, s.dt, flags, s.ni, iterMax=niter, nout=niter)
    oracle.iterate2d_V2(d, s.ni, opts, niter)
    jst.iterate2d_V2_(st, s.pt_stokes, s.grid, s.flow_bcs, (extra["rhogx"], extra["rhogy"]), extra["G"], extra["K"], s.dt, niter)
    compare_slots(st.slots(), d, t2.V2_STATE + t2.V2_DIAG, TOL, "SolCx 511^2 x50")


def test_vc2d_1023_two_phases_plastic(oracle):
    """config 3's kernel at 1023^2: two phases, Drucker-Prager active at centres and vertices, 3 iterations from a random state"""
    import test_gpu_stokes2d as t2

    ni = (1023, 1023)
    f, grid, pt, dt, rat, rheo = t2.random_vc2d(ni, 1023, rho_var=False)
    flags = dict(free_slip=[1, 1, 0, 0, 1, 1], no_slip=[0] * 6, periodic=[0] * 6)
    st, d = t2._run_vc(oracle, ni, f, grid, pt, dt, rat, rheo, flags, 3, False, alias_P=False)
    assert d["lam"].max() > 0 and d["lamv"].max() > 0, "the random state must yield somewhere"
    compare_slots(st, d, t2.VC_STATE + t2.VC_DIAG, TOL, "2D-VC 1023^2 x3")


def test_vc3d_257_three_phases(oracle):
    """config 5's Stokes half per GPU: 257^3, three phases, plasticity active, 2 iterations from a random state"""
    import test_gpu_stokes3d_vc as t3
    from justrelax_jl_b200 import setups

    ni = (257, 257, 257)
    s = setups.random_vc3d(ni, seed=257)
    st, d = t3._run(oracle, s, FS6, 2, False)
    assert d["lam"].max() > 0 and np.abs(d["pyz"]).max() > 0
    compare_slots(st, d, t3.STATE + t3.DIAG, TOL, "3D-VC 257^3 x2")


def test_thermal3d_257_three_phases(oracle):
    """config 5's thermal half per GPU: 257^3, rheology form with three phases, 4 iterations (fused flux + update pairs)"""
    import test_gpu_thermal as tt

    bc = tt.bc_variants(3)[0]
    tt._run_case(oracle, (257, 257, 257), 1, 3, 0, bc, 4)
