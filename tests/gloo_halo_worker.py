"""2-rank gloo worker of tests/test_halo_host.py: literal IGG plane exchange with torch.distributed send/recv (CPU),
compared with libjrb200's source map (jr_halo_source) applied to the all-gathered pre-exchange arrays; also the
deterministic rank-order sum that the device all-reduce implements."""
import numpy as np
import torch
import torch.distributed as dist

from justrelax_jl_b200 import comm


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dims = comm.dims_create(world)
    coords = comm.cart_coords(rank, dims)
    ncell = (6, 5, 7)
    # periods: init_global_grid's periodx / periody / periodz — the grid of ranks wraps around; a rank alone in a periodic dimension
    # exchanges with itself (ImplicitGlobalGrid)
    for periods in [(0, 0, 0), (1, 0, 1), (1, 1, 1)]:
        for grow in [(1, 2, 2), (2, 1, 2), (0, 0, 0)]:
            ext = tuple(ncell[d] + grow[d] for d in range(3))
            rng = np.random.default_rng(100 + rank)
            A = rng.uniform(size=ext)
            before = [torch.empty(ext, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(before, torch.from_numpy(A.copy()))
            # literal exchange, dimension by dimension
            B = torch.from_numpy(A.copy())
            for d in range(3):
                if dims[d] == 1 and not periods[d]:
                    continue
                ol = 2 + ext[d] - ncell[d]
                n = ext[d]
                lo = (coords[d] - 1) % dims[d] if (coords[d] > 0 or periods[d]) else None
                hi = (coords[d] + 1) % dims[d] if (coords[d] < dims[d] - 1 or periods[d]) else None
                rk = lambda cd: int(np.ravel_multi_index(tuple(cd if e == d else coords[e] for e in range(3)), dims))
                send_lo, send_hi = B.select(d, ol - 1).contiguous().clone(), B.select(d, n - ol).contiguous().clone()
                reqs, bufs = [], {}
                if lo is not None:
                    if rk(lo) == rank:
                        bufs["lo"] = send_hi          # my own high send plane arrives in my low ghost plane
                    else:   # tag 0: a message travelling towards lower coordinates, tag 1: towards higher ones
                        reqs.append(dist.isend(send_lo, rk(lo), tag=0))
                        bufs["lo"] = torch.empty_like(send_lo)
                        reqs.append(dist.irecv(bufs["lo"], rk(lo), tag=1))
                if hi is not None:
                    if rk(hi) == rank:
                        bufs["hi"] = send_lo
                    else:
                        reqs.append(dist.isend(send_hi, rk(hi), tag=1))
                        bufs["hi"] = torch.empty_like(send_hi)
                        reqs.append(dist.irecv(bufs["hi"], rk(hi), tag=0))
                for q in reqs:
                    q.wait()
                if "lo" in bufs:
                    B.select(d, 0).copy_(bufs["lo"])
                if "hi" in bufs:
                    B.select(d, n - 1).copy_(bufs["hi"])
            got = A.copy()
            for idx in np.ndindex(*ext):
                if all(0 < idx[d] < ext[d] - 1 for d in range(3)):
                    continue
                moved, sc, si = comm.halo_source(dims, coords, ext, ncell, idx, periods)
                if moved:
                    got[idx] = before[int(np.ravel_multi_index(sc, dims))][si].item()
            assert np.array_equal(got, B.numpy()), (rank, periods, grow)
    # rank-order deterministic sum == what every rank computes
    parts = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(parts, torch.tensor([0.1 * (rank + 1), 1e16, -1e16 + rank, 3.0], dtype=torch.float64))
    acc = parts[0].clone()
    for r in range(1, world):
        acc = acc + parts[r]
    ref = [acc.clone() for _ in range(world)]
    dist.all_gather(ref, acc)
    assert all(torch.equal(ref[0], x) for x in ref)
    print("HALO_OK", rank, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
