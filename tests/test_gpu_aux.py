"""GPU parity (through the C ABI) of the per-time-step kernels around the loops against the oracle: bit-exact for the
interpolations and the column integration, 1e-12 for shear heating; plus the reference's own assertions on the device results."""
import ctypes as C

import numpy as np
import pytest

from test_oracle_aux import litho, shear_heating_case, vel2
from util import device_stokes, max_rel_diff

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ni", [(9, 7), (33, 18), (6, 5, 4), (37, 20, 23)])
def test_velocity_interpolations(oracle, ni):
    from justrelax_jl_b200 import B200Backend, PTArray, stokes as jst, to_host, zeros

    rng = np.random.default_rng(5)
    nd = len(ni)
    shp = [tuple(n + (1 if a == b else 2) for b, n in enumerate(ni)) for a in range(nd)]
    V = [np.asfortranarray(rng.uniform(-1, 1, size=s)) for s in shp]
    Vd = [PTArray(B200Backend)(v) for v in V]
    ev = tuple(n + 1 for n in ni)
    for fn, e in (("velocity2vertex", ev), ("velocity2center", tuple(ni))):
        ref = vel2(oracle, "orc_" + fn, ni, e, V)
        out = [zeros(B200Backend, *e) for _ in range(nd)]
        getattr(jst, fn + "_")(out, Vd, ni)
        for a, b in zip(out, ref):
            assert np.array_equal(to_host(a), b), fn


@pytest.mark.parametrize("shape", [(3, 4), (130, 97), (5, 4, 6), (65, 33, 40)])
def test_lithostatic_pressure(oracle, shape):
    from justrelax_jl_b200 import B200Backend, PTArray, stokes as jst, to_host, zeros

    rng = np.random.default_rng(7)
    rg = np.asfortranarray(rng.uniform(1.0, 3.0, size=shape))
    dzs = rng.uniform(0.2, 2.0, size=shape[-1])
    for dz in (0.5, dzs):
        ref = litho(oracle, np.zeros(shape, order="F"), rg, dz)
        P = zeros(B200Backend, *shape)
        jst.compute_lithostatic_pressure_(P, PTArray(B200Backend)(rg), dz if np.isscalar(dz) else PTArray(B200Backend)(dz))
        assert np.array_equal(to_host(P), ref)
    # the reference's own identities (test/test_Utils.jl:507-519) on the device result
    P = to_host(P)
    top = rg[..., -1] * dzs[-1] / 2
    assert np.allclose(P[..., -1], top)
    with pytest.raises(ValueError, match="must span the same cells"):
        jst.compute_lithostatic_pressure_(zeros(B200Backend, *shape), zeros(B200Backend, *[s + 1 for s in shape]), 0.5)
    with pytest.raises(ValueError, match="one height per cell"):
        jst.compute_lithostatic_pressure_(zeros(B200Backend, *shape), zeros(B200Backend, *shape), zeros(B200Backend, shape[-1] + 1))


@pytest.mark.parametrize("ni,nphase", [((17, 12), 1), ((33, 20), 3), ((9, 8, 7), 1), ((24, 17, 19), 3)])
def test_shear_heating(oracle, ni, nphase):
    from justrelax_jl_b200 import B200Backend, PhaseRatios, ThermalArrays, stokes as jst, to_host

    d, mats, ratios, dt, ref = shear_heating_case(oracle, ni, 11 + ni[0], nphase)
    st, extra = device_stokes(ni, d)
    th = ThermalArrays(B200Backend, *ni)
    if nphase > 1:
        pr = PhaseRatios.from_arrays(B200Backend, **ratios)
        jst.compute_shear_heating_(th, st, pr, mats, dt)
    else:
        jst.compute_shear_heating_(th, st, mats[0], dt)
    got = to_host(th.shear_heating)
    assert (got >= 0).all()                                                                       # test_shearheating2D.jl:246
    assert max_rel_diff(got, ref) <= 1e-12
