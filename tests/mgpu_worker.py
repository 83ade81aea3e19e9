"""Multi-GPU parity worker (run by tests/test_gpu_multi.py under torch.distributed.run, one rank per GPU).

Every rank emulates ALL ranks on the CPU with the oracle + a literal ImplicitGlobalGrid exchange (tests/mrank.py)
and compares its own block of the B200 result with its block of the emulation:
  1. update_halo_ on dense arrays of every staggering, 2. the all-reduce, 3. 3D-VA fixed iterations (fused + unfused),
  4. 3D-VA solve to convergence: iteration count and norm history."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mrank  # noqa: E402
from util import bc_flags, device_stokes, max_rel_diff  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from justrelax_jl_b200 import B200Backend, PTArray, _abi, comm, setups, stokes as jst, to_host
    from oracle import pyoracle as po

    if dist.get_rank() == 0:
        po.build()
    dist.barrier()
    ni = (20, 17, 15)
    igg = comm.init_global_grid(*ni)
    rank, dims, world = igg.me, igg.dims, igg.nprocs
    coords_all = [comm.cart_coords(r, dims) for r in range(world)]

    # ---- 1. dense update_halo_ ---------------------------------------------------------------------------
    for grow in [(0, 0, 0), (1, 2, 2), (2, 1, 2), (2, 2, 1), (2, 2, 2), (1, 1, 0), (-1, 0, 0)]:
        ext = tuple(ni[d] + grow[d] for d in range(3))
        hosts = [np.asfortranarray(np.random.default_rng(11 + r).uniform(size=ext)) for r in range(world)]
        A = PTArray(B200Backend)(hosts[rank])
        B = PTArray(B200Backend)(hosts[rank] * 2.0)
        for _ in range(2):  # twice: the second call reuses the staging buffers (epoch parity)
            comm.update_halo_(A, B, ni=ni)
            mrank.update_halo(hosts, dims, ni)
        assert np.array_equal(to_host(A), hosts[rank]), ("update_halo_", grow, rank)
        assert np.array_equal(to_host(B), hosts[rank] * 2.0), ("update_halo_ 2nd array", grow, rank)

    # ---- 2. all-reduce -----------------------------------------------------------------------------------
    vals = [0.1 * (r + 1) for r in range(world)]
    acc = vals[0]
    for v in vals[1:]:
        acc = acc + v
    assert comm.sum_mpi(vals[rank]) == acc
    assert comm.maximum_mpi(float(rank)) == float(world - 1)
    assert comm.minimum_mpi(float(rank) - 3.0) == -3.0

    # ---- 3. 3D-VA, fixed number of iterations -----------------------------------------------------------------
    names = ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "Rx", "Ry", "Rz", "RP", "etatau"]
    # (push: the in-iteration push exchange of the fused kernel, JRB200_VA_PUSH=1, besides the default pack + pull)
    # (ovl: the default direct exchange overlapped with the next iteration's z-march — head planes, planes per published chunk, CTAs of the
    #  second-stream launch; None = pack + pull after every iteration, JRB200_VA_OVL=0)
    for dt, finite_K, unfused, push, ovl in [(np.inf, False, False, "0", (8, 16, 6)), (0.7, True, False, "0", (8, 16, 6)), (0.7, True, True, "0", None),
                                             (np.inf, False, False, "0", (1, 1, 1)), (0.7, True, False, "0", (3, 4, 2)), (np.inf, False, False, "0", None),
                                             (np.inf, False, False, "1", None), (0.7, True, False, "1", None)]:
        os.environ["JRB200_VA_PUSH"] = push
        os.environ["JRB200_VA_OVL"] = "0" if ovl is None else "1"
        if ovl is not None:
            os.environ["JRB200_VA_OVL_HEAD"], os.environ["JRB200_VA_OVL_CHUNK"], os.environ["JRB200_VA_OVL_CTAS"] = (str(v) for v in ovl)
        blocks = []
        for r in range(world):
            s = setups.random_stokes3d(ni, seed=500 + r, dt=dt, finite_K=finite_K)
            blocks.append(po.alloc_stokes(ni, s.fields))
        flags = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)
        n_g = mrank.n_g(ni, dims)
        opts = po.make_opts(s.pt_stokes, s.grid._di.center, dt, flags, n_g, iterMax=100, nout=100)
        st, extra = device_stokes(ni, blocks[rank])
        niter = 7
        mrank.va_pre(po, blocks, dims, ni)
        mrank.va_iterate(po, blocks, opts, dims, ni, niter)
        from justrelax_jl_b200.types import VelocityBoundaryConditions
        bcs = VelocityBoundaryConditions(free_slip=dict(left=True, right=True, front=True, back=True, top=True, bot=True),
                                         no_slip=dict(left=False, right=False, front=False, back=False, top=False, bot=False))
        jst.set_flags(_abi.JR_FLAG_UNFUSED if unfused else 0)
        jst.iterate_(st, s.pt_stokes, s.grid, bcs, (extra["rhogx"], extra["rhogy"], extra["rhogz"]), extra["K"], extra["G"], dt, niter, igg)
        jst.set_flags(0)
        worst = max(max_rel_diff(to_host(st.slots()[nm]), blocks[rank][nm]) for nm in names)
        assert worst <= 1e-12, ("3D-VA iterate", dt, unfused, push, ovl, rank, worst)
    os.environ["JRB200_VA_PUSH"] = "0"
    for k in ("JRB200_VA_OVL", "JRB200_VA_OVL_HEAD", "JRB200_VA_OVL_CHUNK", "JRB200_VA_OVL_CTAS"):
        os.environ.pop(k, None)

    # ---- 4. SolVi3D solve across ranks: iteration count + norms ----------------------------------------------
    # (divergence-free pure shear + inclusion placed by global coordinates: the loop ends on its tolerance, see setups.solvi3d)
    nis = (10, 10, 10)
    blocks, sets = [], []
    for r in range(world):
        ig = type(igg)(me=r, dims=dims, nprocs=world, coords=coords_all[r])
        s = setups.solvi3d(*nis, igg=ig, smooth_passes=0, divfree=True, global_coords=True)
        sets.append(s)
        blocks.append(po.alloc_stokes(nis, s.fields))
    for _ in range(10):
        for d in blocks:
            d["eta"][...] = setups._smooth3(d["eta"], 1.0)
        mrank.update_halo([d["eta"] for d in blocks], dims, nis)
    s = sets[rank]
    import ctypes as C
    opts = po.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), mrank.n_g(nis, dims), iterMax=3000, nout=10)
    for d in blocks:
        fs = po.make_fields(d, nis)
        po.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)
    for nm in ("Vx", "Vy", "Vz"):
        mrank.update_halo([d[nm] for d in blocks], dims, nis)
    st, extra = device_stokes(nis, blocks[rank])
    it_ref, hist = mrank.va_solve(po, blocks, opts, dims, nis)
    out = jst.solve_(st, s.pt_stokes, s.grid, s.flow_bcs, (extra["rhogx"], extra["rhogy"], extra["rhogz"]), extra["K"], extra["G"], s.dt, igg,
                     kwargs=dict(iterMax=3000, nout=10, verbose=False))
    assert out.iter == it_ref and it_ref < 3000, ("iteration count", out.iter, it_ref)
    assert np.allclose(out.err_evo1, [h[1] for h in hist], rtol=1e-9, atol=0), (out.err_evo1[-3:], hist[-3:])
    worst = max(max_rel_diff(to_host(st.slots()[nm]), blocks[rank][nm]) for nm in names[:10])
    assert worst <= 1e-8, ("converged fields", worst)
    # ---- 5. 3D-VC (multiphase visco-elasto-plastic), fixed number of iterations: ητ, τ-shear and V halos every iteration -----
    from justrelax_jl_b200 import PhaseRatios, rheology as R
    from justrelax_jl_b200.stokes3d_vc import iterate3d_VC_
    niv = (14, 12, 11)
    iggv_ng = mrank.n_g(niv, dims)
    blocks, vcs, sets = [], [], []
    for r in range(world):
        sv = setups.random_vc3d(niv, seed=900 + r)
        sets.append(sv)
        d = po.alloc_stokes(niv, sv.fields)
        d["Pargs"] = d["P"]
        blocks.append(d)
        vcs.append(po.vc_inputs(R.lower_stokes(sv.rheology), R.gravity_of(sv.rheology), sv.ratios))
    sv = sets[rank]
    flags = dict(free_slip=[1] * 6, no_slip=[0] * 6, periodic=[0] * 6)
    optsv = po.make_opts(sv.pt_stokes, sv.grid._di.center, sv.dt, flags, iggv_ng, iterMax=100, nout=100, viscosity_relaxation=0.3,
                         viscosity_cutoff=sv.kwargs["viscosity_cutoff"])
    stv, extrav = device_stokes(niv, blocks[rank])
    mrank.vc_iterate(po, blocks, optsv, vcs, dims, niv, 5, finish=True)
    prv = PhaseRatios.from_arrays(B200Backend, **sv.ratios)
    iggv = type(igg)(me=rank, dims=dims, nprocs=world, coords=coords_all[rank])
    ρgv = (extrav["rhogx"], extrav["rhogy"], extrav["rhogz"])
    iterate3d_VC_(stv, sv.pt_stokes, sv.grid, bcs, ρgv, prv, sv.rheology, dict(T=extrav["T"], P=stv.P), sv.dt, 5, iggv, finish=True,
                  kwargs=dict(viscosity_relaxation=0.3, viscosity_cutoff=sv.kwargs["viscosity_cutoff"]))
    vnames = ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy", "tyz_c", "txz_c", "txy_c", "eta", "etatau", "lam", "Rx", "Ry", "Rz", "RP",
              "pyz", "pxz", "pxy", "tII", "EII_pl", "txy_o"]
    worst_vc = max(max_rel_diff(to_host(stv.slots()[nm]), blocks[rank][nm]) for nm in vnames)
    assert worst_vc <= 1e-12, ("3D-VC iterate", rank, worst_vc, {nm: max_rel_diff(to_host(stv.slots()[nm]), blocks[rank][nm]) for nm in vnames})

    # ---- 6. heatdiffusion_PT! 3D with phase ratios, fixed number of iterations: T halo every iteration ------------------------------
    from justrelax_jl_b200 import thermal as jth
    from justrelax_jl_b200.types import Geometry
    import test_gpu_thermal as tg
    nit = (13, 12, 10)
    li = (1.0e5, 1.1e5, 1.2e5)
    gridt = Geometry(nit, li)
    bct = tg.bc_variants(3)[0]
    tblocks = [po.alloc_thermal(nit, tg.random_thermal(nit, 40 + r, 3)) for r in range(world)]
    ptt = type("PT", (), {})()
    ptt.ϵ, ptt.max_lxyz, ptt.Vpdτ = 1e-8, max(li), min(gridt.di.center) * 0.5
    ot = po.thermal_opts(_di=gridt._di.center, dt=1.0e11, eps=1e-8, iterMax=10, nout=4, max_lxyz=ptt.max_lxyz, Vpdtau=ptt.Vpdτ, form=1,
                         phases=tg.PHASES, bc=bct)
    tht, extrat = tg.to_device(nit, tblocks[rank])
    mrank.thermal_iterate(po, tblocks, ot, dims, nit, 4)
    ptt.θr_dτ, ptt.dτ_ρ = extrat["theta_r_dtau"], extrat["dtau_rho"]
    pht = tg._Phase()
    pht.center, pht.Vx, pht.Vy, pht.Vz = extrat["phase_c"], extrat["phase_x"], extrat["phase_y"], extrat["phase_z"]
    jth.thermal_iterate_(tht, ptt, bct, tg.rheology_of(tg.PHASES), dict(P=extrat["P"], T=tht.T), 1.0e11, gridt, 4,
                         kwargs=dict(verbose=False, phase=pht, igg=iggv))
    tg.compare(tht, tblocks[rank], ["T", "qTx", "qTy", "qTz", "qTx2", "qTy2", "qTz2", "ResT"], f"thermal 3D multi-rank rank {rank}")

    # ---- 7. compute_lithostatic_pressure!(P, ρg, dz, igg) with the vertical direction split across ALL ranks ---------------------------
    comm.finalize_global_grid()
    nl = (9, 8, 12)
    iggz = comm.init_global_grid(*nl, dims=(1, 1, world))
    nzg = world * (nl[2] - 2) + 2
    rg_glob = np.asfortranarray(np.random.default_rng(77).uniform(1.0, 3.0, size=(nl[0], nl[1], nzg)))
    dzl = 0.37
    import ctypes as C2
    P_glob = np.zeros_like(rg_glob, order="F")
    dp = lambda a: a.ctypes.data_as(C2.POINTER(C2.c_double))
    po.lib().orc_lithostatic_pressure(3, (C2.c_int32 * 3)(nl[0], nl[1], nzg), dp(P_glob), dp(rg_glob), C2.c_double(dzl), None, None)
    k0 = iggz.coords[2] * (nl[2] - 2)
    rg_loc = np.asfortranarray(rg_glob[:, :, k0:k0 + nl[2]])
    from justrelax_jl_b200 import zeros
    P_loc = zeros(B200Backend, *nl)
    jst.compute_lithostatic_pressure_(P_loc, PTArray(B200Backend)(rg_loc), dzl, iggz)
    assert np.allclose(to_host(P_loc), P_glob[:, :, k0:k0 + nl[2]], rtol=1e-13, atol=0), ("lithostatic pressure across ranks", rank)
    try:   # the three-argument method must refuse a split vertical direction (test/test_lithostatic_pressure3D_MPI.jl:95)
        jst.compute_lithostatic_pressure_(P_loc, PTArray(B200Backend)(rg_loc), dzl)
        raise AssertionError("expected the split-column error")
    except ValueError as e:
        assert "split across MPI ranks" in str(e)

    # ---- 8. init_global_grid(...; periodx = true, periodz = true): the grid of ranks wraps around ----------------------------------------
    comm.finalize_global_grid()
    periods = (1, 0, 1)
    iggp = comm.init_global_grid(*ni, periodx=1, periodz=1)
    assert tuple(iggp.dims) == tuple(dims)
    for grow in [(0, 0, 0), (1, 2, 2), (2, 1, 2), (2, 2, 1), (2, 2, 2)]:
        ext = tuple(ni[d] + grow[d] for d in range(3))
        hosts = [np.asfortranarray(np.random.default_rng(31 + r).uniform(size=ext)) for r in range(world)]
        A = PTArray(B200Backend)(hosts[rank])
        for _ in range(2):
            comm.update_halo_(A)
            mrank.update_halo(hosts, dims, ni, periods)
        assert np.array_equal(to_host(A), hosts[rank]), ("periodic update_halo_", grow, rank)
    # the reference's known-answer test (test/test_periodic_boundary_conditions_MPI.jl:22-48, here on 3D arrays and any number of ranks):
    # T .= coords[1] + 1; thermal_bcs!(periodic left/right) ; update_halo!(T) → the x ghost planes hold the x-neighbours' values
    from justrelax_jl_b200.types import TemperatureBoundaryConditions
    cx, dimx = iggp.coords[0], dims[0]
    Tp = PTArray(B200Backend)(np.full(tuple(n + 2 for n in ni), float(cx + 1), order="F"))
    bcp = TemperatureBoundaryConditions(no_flux=dict(left=False, right=False, front=True, back=True, top=True, bot=True),
                                        periodic=dict(left=True, right=True, front=False, back=False, top=False, bot=False))
    jth.thermal_bcs_(Tp, bcp)
    comm.update_halo_(Tp)
    Th = to_host(Tp)
    assert np.all(Th[0, 1:-1, 1:-1] == float((cx - 1) % dimx + 1)) and np.all(Th[-1, 1:-1, 1:-1] == float((cx + 1) % dimx + 1)), ("periodic KAT T", rank)
    Vyp = PTArray(B200Backend)(np.full((ni[0] + 2, ni[1] + 1, ni[2] + 2), float(cx + 1), order="F"))
    Vxp = PTArray(B200Backend)(np.full((ni[0] + 1, ni[1] + 2, ni[2] + 2), float(cx + 1), order="F"))
    Vzp = PTArray(B200Backend)(np.full((ni[0] + 2, ni[1] + 2, ni[2] + 1), float(cx + 1), order="F"))
    comm.update_halo_(Vxp, Vyp, Vzp)
    Vh = to_host(Vyp)
    assert np.all(Vh[0, :, 1:-1] == float((cx - 1) % dimx + 1)) and np.all(Vh[-1, :, 1:-1] == float((cx + 1) % dimx + 1)), ("periodic KAT Vy", rank)
    # 3D-VA iterations on the periodic grid of ranks (fused and unfused) against the emulation
    for dt, finite_K, unfused in [(np.inf, False, False), (0.7, True, False), (0.7, True, True)]:
        blocks = []
        for r in range(world):
            s = setups.random_stokes3d(ni, seed=700 + r, dt=dt, finite_K=finite_K)
            blocks.append(po.alloc_stokes(ni, s.fields))
        flags = dict(free_slip=[0, 0, 1, 1, 0, 0], no_slip=[0] * 6, periodic=[0] * 6)
        opts = po.make_opts(s.pt_stokes, s.grid._di.center, dt, flags, mrank.n_g(ni, dims, periods), iterMax=100, nout=100)
        st, extra = device_stokes(ni, blocks[rank])
        mrank.va_pre(po, blocks, dims, ni, periods)
        mrank.va_iterate(po, blocks, opts, dims, ni, 5, periods)
        bcsp = VelocityBoundaryConditions(free_slip=dict(left=False, right=False, front=True, back=True, top=False, bot=False),
                                          no_slip=dict(left=False, right=False, front=False, back=False, top=False, bot=False))
        jst.set_flags(_abi.JR_FLAG_UNFUSED if unfused else 0)
        jst.iterate_(st, s.pt_stokes, s.grid, bcsp, (extra["rhogx"], extra["rhogy"], extra["rhogz"]), extra["K"], extra["G"], dt, 5, iggp)
        jst.set_flags(0)
        worst_p = max(max_rel_diff(to_host(st.slots()[nm]), blocks[rank][nm]) for nm in names)
        assert worst_p <= 1e-12, ("3D-VA iterate on a periodic grid of ranks", dt, unfused, rank, worst_p)

    dist.barrier()
    print(f"MGPU_OK rank {rank}/{world} dims {dims}: halo, all-reduce, 3D-VA iterate (fused+unfused), solve iter={out.iter} worst={worst:.2e}, 3D-VC worst={worst_vc:.2e}, thermal OK, periodic ranks OK", flush=True)
    comm.finalize_global_grid()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
