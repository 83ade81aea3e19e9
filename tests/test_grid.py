"""Geometry — the reference's grid known-answer tests on the host mirror (test/test_grid2D.jl:9-74, test/test_grid3D.jl:9-66):
the uniform constructor (vertex / centre / staggered velocity grids), the constructor from explicit vertex coordinates
(non-uniform meshes: midpoints, vector spacings, ghost-extended velocity grids), legacy_uniform_grid, and the loud refusal of
non-uniform grids by the solvers (the B200 kernels take scalar spacings)."""
import numpy as np
import pytest

from justrelax_jl_b200.types import Geometry, IGG, legacy_uniform_grid


@pytest.mark.parametrize("nD", [2, 3])
def test_uniform_geometry_reference_kats(nD):
    n = 4
    ni, li = (n,) * nD, (1.0,) * nD
    origin = (0.0, -1.0) if nD == 2 else (0.0, 0.0, -1.0)
    di = tuple(l / n for l in li)
    grid = Geometry(ni, li, origin=origin)
    assert grid.origin == origin                                          # test_grid2D.jl:26
    for i in range(nD):
        assert grid.xvi[i][0] == origin[i]                                # :29
        assert grid.xci[i][0] == origin[i] + di[i] / 2                    # :31
    assert grid.xi_vel[0][1][0] == origin[1] - di[0] / 2                  # :34  (Vx grid: ghost point below the first centre in y)
    assert grid.xi_vel[1][0][0] == origin[0] - di[1] / 2                  # :35
    if nD == 3:
        assert grid.xi_vel[2][0][0] == origin[0] - di[2] / 2              # test_grid3D.jl:35
    for c in range(nD):
        for d in range(nD):
            assert len(grid.xi_vel[c][d]) == (n + 1 if d == c else n + 2)
    assert grid.uniform


def test_nonuniform_geometry_2d_reference_kats():
    xv1 = np.linspace(0.0, 1.0, 5)
    xv2 = np.array([0.0, 0.4, 0.7, 0.9, 1.0])                             # non-uniform along y
    g = Geometry.from_vertices(xv1, xv2)
    assert g.ni == (4, 4) and g.li == (1.0, 1.0) and g.origin == (0.0, 0.0) and g.max_li == 1.0     # test_grid2D.jl:41-44
    assert len(g.xci[0]) == 4 and len(g.xvi[0]) == 5                                                # :45
    assert np.array_equal(g.xci[1], (xv2[:-1] + xv2[1:]) / 2)                                       # :47
    assert np.array_equal(g.di.vertex[1], np.diff(xv2))                                             # :49
    assert isinstance(g.di.center, tuple) and len(g.di.center) == 2 and all(isinstance(a, np.ndarray) for a in g.di.center)   # :50
    assert len(g.xi_vel[0][1]) == len(g.xci[1]) + 2 and len(g.xi_vel[1][0]) == len(g.xci[0]) + 2    # :53-54
    # ghost points continue the first / last centre spacing (velocity_grids, Grid.jl:171-182)
    assert g.xi_vel[0][1][0] == g.xci[1][0] - (g.xci[1][1] - g.xci[1][0])
    assert g.xi_vel[0][1][-1] == g.xci[1][-1] + (g.xci[1][-1] - g.xci[1][-2])
    assert np.array_equal(g._di.vertex[1], 1.0 / np.diff(xv2))
    g2 = Geometry.from_vertices((xv1, xv2))                                                          # tuple dispatcher :57-59
    assert g2.ni == g.ni and g2.li == g.li
    assert not g.uniform


def test_nonuniform_geometry_3d_reference_kats():
    xv1, xv2, xv3 = np.linspace(0.0, 1.0, 5), np.array([0.0, 0.4, 0.7, 0.9, 1.0]), np.linspace(0.0, 2.0, 5)
    g = Geometry.from_vertices(xv1, xv2, xv3)
    assert g.ni == (4, 4, 4) and g.li == (1.0, 1.0, 2.0) and g.origin == (0.0, 0.0, 0.0) and g.max_li == 2.0   # test_grid3D.jl:42-45
    assert np.array_equal(g.xci[1], (xv2[:-1] + xv2[1:]) / 2)                                                  # :46
    assert np.array_equal(g.di.vertex[1], np.diff(xv2))                                                        # :47
    assert len(g.xi_vel) == 3                                                                                  # :48
    assert len(g.xi_vel[0][1]) == len(g.xci[1]) + 2 and len(g.xi_vel[2][0]) == len(g.xci[0]) + 2               # :50-51
    g2 = Geometry.from_vertices((xv1, xv2, xv3))
    assert g2.ni == g.ni and g2.li == g.li


@pytest.mark.parametrize("nD", [2, 3])
def test_legacy_uniform_grid(nD):
    n = 4
    ni, di = (n,) * nD, (0.25,) * nD
    leg = legacy_uniform_grid(ni, di)                                      # test_grid2D.jl:62-64
    assert leg.ni == ni and leg.li == (1.0,) * nD
    grid = Geometry(ni, (1.0,) * nD)
    leg_nt = legacy_uniform_grid(ni, grid.di.center)                       # NamedTuple variant forwards to .center  :67-69
    assert leg_nt.ni == ni and leg_nt.li == (1.0,) * nD


def test_solvers_refuse_nonuniform_grids():
    from justrelax_jl_b200 import stokes as jst, thermal as jth

    g = Geometry.from_vertices(np.linspace(0, 1, 5), np.array([0.0, 0.4, 0.7, 0.9, 1.0]))
    with pytest.raises(NotImplementedError, match="non-uniform"):
        jst._grid_of(None, g, IGG())
    with pytest.raises(NotImplementedError, match="non-uniform"):
        jth._grid_of(None, g, IGG())
    with pytest.raises(NotImplementedError, match="non-uniform"):
        jst._grid_of(type("S", (), {"ni": (4, 4)})(), g.di, IGG())
