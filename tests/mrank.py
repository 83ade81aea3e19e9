"""Multi-rank emulation for the parity tests (test infrastructure).

`update_halo` is a LITERAL restatement of ImplicitGlobalGrid's update_halo! on a list of per-rank numpy arrays
(SURVEY.md §5): for dim in x, y, z: every rank sends plane `ol` (1-based) to its low neighbour and plane
`size-ol+1` to its high neighbour, receives into planes 1 and `size`; ol = 2 + (size(A,d) − n_d); arrays with
ol < 2 are skipped in that dimension; `periods[d]` wraps the grid of ranks around (a rank alone in a periodic dimension
exchanges with itself).  Ranks are numbered like MPI_Cart (last dim fastest).
"""
import itertools

import numpy as np


def cart_rank(c, dims):
    return (c[0] * dims[1] + c[1]) * dims[2] + c[2]


def all_coords(dims):
    return list(itertools.product(range(dims[0]), range(dims[1]), range(dims[2])))


def update_halo(per_rank, dims, ncell, periods=(0, 0, 0)):
    """per_rank: list (MPI_Cart rank order) of numpy arrays of identical shape — exchanged in place."""
    nd = per_rank[0].ndim
    shp = per_rank[0].shape
    for d in range(nd):
        if dims[d] == 1 and not periods[d]:
            continue
        ol = 2 + (shp[d] - ncell[d])
        if ol < 2:
            continue
        n = shp[d]
        take = lambda A, i: np.take(A, i, axis=d).copy()
        # all sends are posted from the pre-phase state, then all receives land (sendrecv semantics)
        send_lo = {c: take(per_rank[cart_rank(c, dims)], ol - 1) for c in all_coords(dims)}
        send_hi = {c: take(per_rank[cart_rank(c, dims)], n - ol) for c in all_coords(dims)}
        for c in all_coords(dims):
            A = per_rank[cart_rank(c, dims)]
            idx = [slice(None)] * nd
            if c[d] > 0 or periods[d]:
                lo = list(c); lo[d] = (lo[d] - 1) % dims[d]
                idx[d] = 0
                A[tuple(idx)] = send_hi[tuple(lo)]
            if c[d] < dims[d] - 1 or periods[d]:
                hi = list(c); hi[d] = (hi[d] + 1) % dims[d]
                idx[d] = n - 1
                A[tuple(idx)] = send_lo[tuple(hi)]


def n_g(ni, dims, periods=(0, 0, 0)):
    return tuple(dims[d] * (ni[d] - 2) + (0 if periods[d] else 2) for d in range(len(ni)))


# ---------------------------------------------------------------------------------------------------------
# multi-rank 3D-VA oracle: N independent oracle blocks + literal update_halo! between the kernels, exactly where the
# reference calls it (Stokes3D.jl:57 ητ, :120 V)
def va_pre(po, ranks, dims, ni, periods=(0, 0, 0)):
    import ctypes as C
    for d in ranks:
        fs = po.make_fields(d, ni)
        po.lib().orc_pre3d_VA(C.byref(fs))
    update_halo([d["etatau"] for d in ranks], dims, ni, periods)


def va_iterate(po, ranks, opts, dims, ni, niter, periods=(0, 0, 0)):
    import ctypes as C
    for _ in range(niter):
        for d in ranks:
            fs = po.make_fields(d, ni)
            po.lib().orc_iterate3d_VA_once(C.byref(fs), C.byref(opts))
        for nm in ("Vx", "Vy", "Vz"):
            update_halo([d[nm] for d in ranks], dims, ni, periods)


def va_solve(po, ranks, opts, dims, ni):
    """_solve! 3D-VA across emulated ranks (Stokes3D.jl:76-167): norm_mpi sums the local interior sums of squares."""
    import math
    nx, ny, nz = ni
    g = opts.n_g
    va_pre(po, ranks, dims, ni)
    err_it1 = err = 1.0
    it, hist = 0, []
    while it < 2 or ((err / err_it1 > opts.eps_rel and err > opts.eps_abs) and it <= opts.iterMax):
        va_iterate(po, ranks, opts, dims, ni, 1)
        it += 1
        if it % opts.nout == 0 and it > 1:
            S = [0.0] * 4
            for d in ranks:  # rank order, like the device all-reduce
                for q, (nm, inter) in enumerate((("Rx", 1), ("Ry", 1), ("Rz", 1), ("RP", 0))):
                    S[q] = S[q] + po.sumsq(d[nm], inter)
            nrm = [math.sqrt(S[0]) / ((g[0] - 2) * (g[1] - 1) * (g[2] - 1)), math.sqrt(S[1]) / ((g[0] - 1) * (g[1] - 2) * (g[2] - 1)),
                   math.sqrt(S[2]) / ((g[0] - 1) * (g[1] - 1) * (g[2] - 2)), math.sqrt(S[3]) / (g[0] * g[1] * g[2])]
            err = max(nrm)
            hist.append((it, err, nrm))
            err_it1 = max(hist[0][2])
    return it, hist


# ---------------------------------------------------------------------------------------------------------
# multi-rank 3D-VC oracle: the loop body in the three pieces the reference separates with update_halo! (Stokes3D.jl:515 ητ,
# :578-580 τyz/τxz/τxy, :596 V)
def vc_iterate(po, ranks, opts, vcs, dims, ni, niter, finish=False, periods=(0, 0, 0)):
    import ctypes as C
    L = po.lib()
    fss = [po.make_fields(d, ni) for d in ranks]
    hs = [C.c_void_p(L.orc_vc3_begin(C.byref(fs), C.byref(opts), C.byref(vc))) for fs, vc in zip(fss, vcs)]
    for _ in range(niter):
        for piece, names in ((0, ("etatau",)), (1, ("tyz", "txz", "txy")), (2, ("Vx", "Vy", "Vz"))):
            for fs, vc, h in zip(fss, vcs, hs):
                L.orc_vc3_step(C.byref(fs), C.byref(opts), C.byref(vc), h, piece)
            for nm in names:
                update_halo([d[nm] for d in ranks], dims, ni, periods)
    for fs, h in zip(fss, hs):
        L.orc_vc3_end(C.byref(fs), C.byref(opts), h, int(finish))


# multi-rank heatdiffusion_PT! iterations: update_halo!(thermal.T) after thermal_bcs! (DiffusionPT_solver.jl:110, 261)
def thermal_iterate(po, ranks, opts, dims, ni, niter, periods=(0, 0, 0)):
    import ctypes as C
    fss = [po.thermal_fields(d, ni) for d in ranks]
    for _ in range(niter):
        for fs in fss:
            po.lib().orc_thermal_iterate_once(C.byref(fs), C.byref(opts))
        update_halo([d["T"] for d in ranks], dims, ni, periods)
    for fs in fss:
        po.lib().orc_thermal_check_res(C.byref(fs), C.byref(opts))
