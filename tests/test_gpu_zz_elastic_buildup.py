"""GPU run of the reference's visco-elastic build-up test (test/test_stokes_elastic_buildup.jl) through the public API.
Kept in its own file, collected after the other GPU suites."""
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
