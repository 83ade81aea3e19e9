"""test/test_boundary_conditions2D.jl:67-180 restated at the two places the path has them (no GPU): the host mirror of the BC types
(same exceptions, same messages) and the oracle's flow_bcs! (free slip / periodic / no slip ghost identities, incl. the reshape(1:42) KAT)."""
import ctypes as C

import numpy as np
import pytest

from justrelax_jl_b200 import setups
from justrelax_jl_b200.types import AbstractFlowBoundaryConditions, VelocityBoundaryConditions

F4 = lambda **kw: dict(left=False, right=False, top=False, bot=False) | kw


def test_velocity_bc_validation_errors():
    # test_boundary_conditions2D.jl:70-83, 114-129
    with pytest.raises(RuntimeError):
        VelocityBoundaryConditions(no_slip=F4(left=True), free_slip=F4(left=True, right=True, top=True, bot=True))
    with pytest.raises(RuntimeError):
        VelocityBoundaryConditions(no_slip=F4(bot=True), free_slip=F4(left=True, right=True, top=True, bot=True))
    bcs = VelocityBoundaryConditions(no_slip=F4(), free_slip=F4())   # neither: prescribed velocity
    assert isinstance(bcs, VelocityBoundaryConditions) and isinstance(bcs, AbstractFlowBoundaryConditions)
    with pytest.raises(RuntimeError, match="Periodic boundary conditions must be paired"):
        VelocityBoundaryConditions(no_slip=F4(), free_slip=F4(), periodic=F4(left=True))
    with pytest.raises(RuntimeError, match="Incompatible boundary conditions on the left boundary"):
        VelocityBoundaryConditions(no_slip=F4(), free_slip=F4(left=True), periodic=F4(left=True, right=True))
    with pytest.raises(RuntimeError, match="top can't be both periodic and free_surface"):
        VelocityBoundaryConditions(no_slip=F4(), free_slip=F4(), periodic=F4(top=True, bot=True), free_surface=True)


def _apply(oracle, Vx, Vy, bcs):
    n = Vx.shape[0] - 1
    ni = (n, n)
    d = oracle.alloc_stokes(ni, dict(Vx=np.asfortranarray(Vx), Vy=np.asfortranarray(Vy)))
    pt = setups.PTStokesCoeffs((1.0, 1.0), (1.0 / n, 1.0 / n))
    flags = dict(free_slip=bcs.flags("free_slip"), no_slip=bcs.flags("no_slip"), periodic=bcs.flags("periodic"))
    opts = oracle.make_opts(pt, (float(n), float(n)), 1.0, flags, ni, iterMax=1, nout=1)
    fs = oracle.make_fields(d, ni)
    oracle.lib().orc_flow_bcs2(C.byref(fs), C.byref(opts), 0)
    return d["Vx"], d["Vy"]


def test_flow_bcs2_reference_identities(oracle):
    rng = np.random.default_rng(5)
    n = 5
    # free slip  :86-96
    bcs = VelocityBoundaryConditions(no_slip=F4(), free_slip=F4(left=True, right=True, top=True, bot=True))
    Vx, Vy = _apply(oracle, rng.uniform(size=(n + 1, n + 2)), rng.uniform(size=(n + 2, n + 1)), bcs)
    assert np.array_equal(Vx[:, 0], Vx[:, 1]) and np.array_equal(Vx[:, -1], Vx[:, -2])
    assert np.array_equal(Vy[0, :], Vy[1, :]) and np.array_equal(Vy[-1, :], Vy[-2, :])
    # periodic left/right on the 1:42 arrays  :100-112
    Vx0 = np.arange(1.0, 43.0).reshape((6, 7), order="F")
    Vy0 = np.arange(1.0, 43.0).reshape((7, 6), order="F")
    bcs = VelocityBoundaryConditions(no_slip=F4(), free_slip=F4(), periodic=F4(left=True, right=True))
    Vx, Vy = _apply(oracle, Vx0.copy(order="F"), Vy0.copy(order="F"), bcs)
    assert np.array_equal(Vx[0, :], Vx0[-1, :]) and np.array_equal(Vx[-1, :], Vx0[-1, :])
    assert np.array_equal(Vy[0, :], Vy0[-2, :]) and np.array_equal(Vy[-1, :], Vy0[1, :])
    # no slip  :130-144
    bcs = VelocityBoundaryConditions(no_slip=F4(left=True, right=True, top=True, bot=True), free_slip=F4())
    Vx, Vy = _apply(oracle, rng.uniform(size=(n + 1, n + 2)), rng.uniform(size=(n + 2, n + 1)), bcs)
    assert not Vx[0, :].any() and not Vx[-1, :].any() and not Vy[:, 0].any() and not Vy[:, -1].any()
    assert np.array_equal(Vy[0, :], -Vy[1, :]) and np.array_equal(Vy[-1, :], -Vy[-2, :])
    assert np.array_equal(Vx[:, 0], -Vx[:, 1]) and np.array_equal(Vx[:, -1], -Vx[:, -2])
