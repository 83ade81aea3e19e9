"""GPU runs of further reference tests through the public API, with the reference's own quantitative criteria:
test/test_stokes_elastic_buildup.jl (2D-V2, visco-elastic) and test/test_stokes_burstedde.jl (3D-VA, manufactured solution).
Collected after the other GPU suites."""
import numpy as np
import pytest

from util import compare_slots, device_stokes

pytestmark = pytest.mark.gpu

V2_STATE = ["Vx", "Vy", "P", "txx", "tyy", "txy"]


def test_elastic_buildup_reference_criterion_on_gpu(oracle):
    """test/test_stokes_elastic_buildup.jl:24-53 through the public API: 200 visco-elastic time steps of 2D-V2 (finite G and dt, SI units):
    the reference's analytic criterion (mean relative error of max|τyy| ≤ 5e-3), and iteration counts / final fields as the oracle"""
    from justrelax_jl_b200 import setups, stokes as jst, to_host
    from test_oracle_stokes2d import run_elastic_buildup

    s = setups.elastic_buildup2d(32)
    d, av_o, sol, iters_o = run_elastic_buildup(oracle, s)
    st, extra = device_stokes(s.ni, oracle.alloc_stokes(s.ni, s.fields))
    jst.flow_bcs_(st, s.flow_bcs)
    t, av, iters = 0.0, [], []
    while t < s.ttot:
        dt = s.dt_of(t)
        out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, (extra["rhogx"], extra["rhogy"]), extra["G"], extra["K"], dt, s.igg, kwargs=s.kwargs)
        t += dt
        av.append(np.abs(to_host(st.τ.yy)).max()); iters.append(out.iter)
    err = np.mean(np.abs(np.abs(np.array(av)) - sol) / sol)
    assert err <= 5.0e-3, err
    assert iters == iters_o
    compare_slots(st.slots(), d, V2_STATE + ["txx_o", "txy_o"], 1.0e-10, "elastic build-up after 200 steps")


@pytest.mark.parametrize("case", ["burstedde", "taylor_green"])
def test_manufactured_solutions_on_gpu(oracle, case):
    """test/test_stokes_burstedde.jl:28-40 and test/test_stokes_taylor_green.jl:29-40 through the public API (3D-VA, streamed body force,
    velocity prescribed on every face — no free-slip / no-slip flag, so the boundary kernel only carries the prescribed layers over):
    the reference's convergence-order and error criteria at 8³ and 16³, and iteration counts / fields as the oracle"""
    from justrelax_jl_b200 import setups, stokes as jst, to_host
    from test_oracle_stokes3d import run_burstedde

    setup = setups.burstedde3d if case == "burstedde" else setups.taylor_green3d
    errs = []
    for n in (8, 16):
        s, d, ref = run_burstedde(oracle, n, setup)
        st, extra = device_stokes(s.ni, oracle.alloc_stokes(s.ni, s.fields))
        jst.flow_bcs_(st, s.flow_bcs)
        out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, (extra["rhogx"], extra["rhogy"], extra["rhogz"]), extra["K"], extra["G"], s.dt, s.igg,
                         kwargs=s.kwargs)
        assert out.err_evo1[-1] < 1.0e-8 and out.iter == ref["iter"]
        errs.append(s.error_norms(to_host(st.V.Vx), to_host(st.V.Vy), to_host(st.V.Vz), to_host(st.P)))
        compare_slots(st.slots(), d, ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy"], 1.0e-10, f"{case} {n}^3")
    order = np.log2(np.array(errs[0]) / np.array(errs[1]))
    L2_p, L2_vx, L2_vy, L2_vz = errs[1]
    if case == "burstedde":
        assert np.all(order[1:] > 1.4), order
        assert max(L2_vx, L2_vy, L2_vz) < 3.0e-2 and L2_p < 2.0e-1
    else:
        assert np.all(order > 1.7), order
        assert max(L2_vx, L2_vy, L2_vz) < 5.0e-3 and L2_p < 1.5e-1


def test_solkz_on_gpu(oracle):
    """test/test_stokes_solkz.jl:26-37 through the public API (2D-V2, η over six decades, Re = 5π): converges below 1e-8 with the oracle's
    iteration count and fields"""
    from justrelax_jl_b200 import setups, stokes as jst
    from test_oracle_stokes2d import run_solkz

    s = setups.solkz2d(32, 32)
    d, ref = run_solkz(oracle, s)
    st, extra = device_stokes(s.ni, oracle.alloc_stokes(s.ni, s.fields))
    jst.flow_bcs_(st, s.flow_bcs)
    out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, (extra["rhogx"], extra["rhogy"]), extra["G"], extra["K"], s.dt, s.igg, kwargs=s.kwargs)
    assert out.err_evo1[-1] < 1.0e-8 and out.iter == ref["iter"]
    compare_slots(st.slots(), d, V2_STATE, 1.0e-8, "solkz converged fields")
