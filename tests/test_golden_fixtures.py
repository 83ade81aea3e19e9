"""The committed golden vectors of tests/golden/ (written by tests/golden/make_fixtures.py after the oracle passed the reference's own
known answers): the oracle still reproduces them (no GPU) and the CUDA path matches them through the C ABI (-m gpu)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from util import device_stokes, max_rel_diff  # noqa: E402

CASES = ("oracle_va3d", "oracle_vc3d", "oracle_stokes2d", "oracle_thermal")


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_golden_fixtures(oracle, case):
    import make_fixtures as mf

    want = np.load(os.path.join(HERE, "golden", case + ".npz"))
    got = mf.CASES[case](oracle)
    assert sorted(want.files) == sorted(got)
    for k in want.files:
        assert max_rel_diff(got[k], want[k]) <= 1e-14, (case, k)


@pytest.mark.gpu
def test_cuda_va3d_matches_golden_fixture():
    from justrelax_jl_b200 import setups, stokes as jst, to_host
    from justrelax_jl_b200.types import VelocityBoundaryConditions

    want = np.load(os.path.join(HERE, "golden", "oracle_va3d.npz"))
    s = setups.random_stokes3d((9, 8, 7), seed=20261017)
    z = {k: np.zeros(sh, order="F") for k, sh in (("rhogx", s.ni), ("rhogy", s.ni), ("rhogz", s.ni), ("K", s.ni), ("G", s.ni))}
    st, extra = device_stokes(s.ni, {**z, **s.fields})
    bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, front=True, back=True, top=True, bot=True))
    jst.iterate_(st, s.pt_stokes, s.grid, bcs, (extra["rhogx"], extra["rhogy"], extra["rhogz"]), extra["K"], extra["G"], s.dt, 4)
    for k in want.files:
        assert max_rel_diff(to_host(st.slots()[k]), want[k]) <= 1e-12, k


@pytest.mark.gpu
def test_cuda_vc3d_matches_golden_fixture():
    from justrelax_jl_b200 import B200Backend, PhaseRatios, setups, to_host
    from justrelax_jl_b200.stokes3d_vc import iterate3d_VC_
    from justrelax_jl_b200.types import VelocityBoundaryConditions

    want = np.load(os.path.join(HERE, "golden", "oracle_vc3d.npz"))
    s = setups.random_vc3d((9, 8, 7), seed=20261017)
    st, extra = device_stokes(s.ni, s.fields)
    bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, front=True, back=True, top=True, bot=True))
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    ρg = (extra["rhogx"], extra["rhogy"], extra["rhogz"])
    iterate3d_VC_(st, s.pt_stokes, s.grid, bcs, ρg, pr, s.rheology, dict(T=extra["T"], P=st.P), s.dt, 4, finish=True,
                  kwargs=dict(viscosity_relaxation=0.3, viscosity_cutoff=s.kwargs["viscosity_cutoff"]))
    got = {**st.slots(), "rhogz": ρg[2]}
    for k in want.files:
        assert max_rel_diff(to_host(got[k]), want[k]) <= 1e-12, k


@pytest.mark.gpu
def test_cuda_phase_ratios_bit_exact(oracle):
    """update_phase_ratios_3D!/2D! on the B200 vs the oracle: bit-exact (north star: bit-exact phase arrays)"""
    from justrelax_jl_b200 import B200Backend, PhaseRatios, PTArray, to_host
    from justrelax_jl_b200.types import update_phase_ratios_

    rng = np.random.default_rng(7)
    for ni in ((9, 8, 7), (33, 5, 12), (17, 13)):
        N = 3
        raw = rng.dirichlet(np.ones(N), size=ni)
        raw[rng.uniform(size=ni) < 0.3] = np.eye(N)[0]
        raw[..., 2] *= (rng.uniform(size=ni) > 0.2)            # exact zeros; sums no longer one → exercises the normalisation
        raw[..., 1] = np.where(rng.uniform(size=ni) < 0.1, 1.0e-6, raw[..., 1])   # below-threshold values
        ph = [np.asfortranarray(raw[..., p]) for p in range(N)]
        xv = [np.linspace(-0.3, 1.1 + 0.1 * d, n + 1) for d, n in enumerate(ni)]
        xc = [0.5 * (x[1:] + x[:-1]) for x in xv]
        want = oracle.phase_ratios_from_arrays(ph, xc, xv)
        pr = PhaseRatios(B200Backend, N, ni)
        update_phase_ratios_(pr, [PTArray(B200Backend)(a) for a in ph], xc, xv)
        for k, w in want.items():
            assert np.array_equal(to_host(getattr(pr, k)), w), (ni, k)
