"""Host logic of the multi-GPU layer (no GPU): the one-shot source map of libjrb200's halo pull (jr_halo_source,
the index arithmetic k_halo_pull runs on the device) equals a literal x → y → z ImplicitGlobalGrid exchange, and the
same exchange done with real messages between 2 gloo ranks."""
import os
import subprocess
import sys

import numpy as np
import pytest

import mrank

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("dims", [(2, 1, 1), (2, 2, 1), (2, 2, 2), (3, 2, 1), (1, 1, 2)])
@pytest.mark.parametrize("grow", [(0, 0, 0), (1, 2, 2), (2, 1, 2), (2, 2, 1), (-1, 0, 0), (1, 1, 0)])
def test_halo_source_equals_sequential_exchange(dims, grow):
    from justrelax_jl_b200 import comm

    ncell = (7, 6, 8)
    ext = tuple(ncell[d] + grow[d] for d in range(3))
    nr = dims[0] * dims[1] * dims[2]
    rng = np.random.default_rng(7)
    before = [np.asfortranarray(rng.uniform(size=ext)) for _ in range(nr)]
    after = [a.copy(order="F") for a in before]
    mrank.update_halo(after, dims, ncell)
    for c in mrank.all_coords(dims):
        r = mrank.cart_rank(c, dims)
        got = before[r].copy(order="F")
        for idx in np.ndindex(*ext):
            if all(0 < idx[d] < ext[d] - 1 for d in range(3)):
                continue
            moved, sc, si = comm.halo_source(dims, c, ext, ncell, idx)
            if moved:
                got[idx] = before[mrank.cart_rank(sc, dims)][si]
        assert np.array_equal(got, after[r]), (dims, grow, c)


@pytest.mark.parametrize("dims,periods", [((2, 1, 1), (1, 0, 0)), ((1, 1, 1), (1, 0, 1)), ((2, 2, 1), (1, 1, 0)), ((3, 1, 2), (1, 1, 1)),
                                          ((2, 2, 2), (0, 1, 0))])
@pytest.mark.parametrize("grow", [(0, 0, 0), (1, 2, 2), (2, 1, 2), (2, 2, 1)])
def test_halo_source_periodic_equals_sequential_exchange(dims, periods, grow):
    """init_global_grid(...; periodx, periody, periodz): the grid of ranks wraps around, a rank alone in a periodic dimension
    exchanges with itself (ImplicitGlobalGrid; test/test_periodic_boundary_conditions_MPI.jl:12-19)."""
    from justrelax_jl_b200 import comm

    ncell = (7, 6, 8)
    ext = tuple(ncell[d] + grow[d] for d in range(3))
    nr = dims[0] * dims[1] * dims[2]
    rng = np.random.default_rng(11)
    before = [np.asfortranarray(rng.uniform(size=ext)) for _ in range(nr)]
    after = [a.copy(order="F") for a in before]
    mrank.update_halo(after, dims, ncell, periods)
    for c in mrank.all_coords(dims):
        r = mrank.cart_rank(c, dims)
        got = before[r].copy(order="F")
        for idx in np.ndindex(*ext):
            if all(0 < idx[d] < ext[d] - 1 for d in range(3)):
                continue
            moved, sc, si = comm.halo_source(dims, c, ext, ncell, idx, periods)
            if moved:
                got[idx] = before[mrank.cart_rank(sc, dims)][si]
        assert np.array_equal(got, after[r]), (dims, periods, grow, c)


def test_periodic_global_sizes():
    """nx_g = dims·(nx − overlap) + overlap·(period == 0)  (ImplicitGlobalGrid init_global_grid; used by src/grid/Utils.jl:29-83)."""
    from justrelax_jl_b200.types import IGG

    assert IGG(dims=(2, 1, 1), periods=(1, 0, 0)).n_g((8, 6, 1 + 0)) [:2] == (12, 6)
    assert IGG(dims=(2, 1, 1)).n_g((8, 6)) == (14, 6)
    assert IGG(dims=(1, 1, 1), periods=(1, 0, 1)).n_g((8, 6, 5)) == (6, 6, 3)


def test_2d_solvers_refuse_a_periodic_grid_of_ranks():
    """the 2D solvers do no halo exchange: a periodic grid of ranks must not be ignored silently (host check; the library returns
    JR_ERR_UNSUPPORTED for a periodic communicator as well)."""
    from justrelax_jl_b200 import stokes as jst
    from justrelax_jl_b200.types import IGG

    with pytest.raises(NotImplementedError, match="one non-periodic rank"):
        jst._single_rank2d(IGG(periods=(1, 0, 0)))
    jst._single_rank2d(IGG())


def test_lithostatic_pressure_rejects_vertically_periodic_columns():
    """test/test_lithostatic_pressure2D_MPI.jl:131-143: the vertical direction split across ranks AND periodic → error"""
    from justrelax_jl_b200 import stokes as jst
    from justrelax_jl_b200.types import IGG

    class A:
        shape = (6, 5)

        def dim(self):
            return 2

    with pytest.raises(RuntimeError, match="periodic along the vertical"):
        jst.compute_lithostatic_pressure_(A(), A(), 0.5, IGG(dims=(1, 2, 1), nprocs=2, periods=(0, 1, 0)))


def test_dims_create_and_cart_coords():
    from justrelax_jl_b200 import comm

    assert comm.dims_create(8) == (2, 2, 2)
    assert comm.dims_create(4) == (2, 2, 1)
    assert comm.dims_create(2) == (2, 1, 1)
    assert comm.dims_create(1) == (1, 1, 1)
    assert comm.dims_create(4, 2) == (2, 2, 1)
    assert comm.dims_create(6) == (3, 2, 1)
    for dims in [(2, 2, 2), (3, 2, 1)]:
        for r in range(dims[0] * dims[1] * dims[2]):
            assert mrank.cart_rank(comm.cart_coords(r, dims), dims) == r


def test_gloo_two_ranks_exchange_matches_source_map(tmp_path):
    """world_size-2 gloo run: real send/recv of the IGG planes vs the source map applied to the gathered arrays."""
    script = os.path.join(ROOT, "tests", "gloo_halo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", script], env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("HALO_OK") == 2, p.stdout[-2000:]
