#!/usr/bin/env python
"""Config 5 of BASELINE.json: 3D thermal convection with grid-based phases, coupled Stokes (variant 3D-VC) + heatdiffusion_PT!
(rheology form with phase ratios), weak-scaled: every rank owns an n³ block of ONE global problem decomposed like ImplicitGlobalGrid
(overlap 2; 8 GPUs → 2×2×2, local 257³ = global 512³).  One JSON line (rank 0): Stokes and thermal PT iterations/s, T_eff per GPU with the
A_eff of SURVEY.md §8d (512 and 200 B/cell at N = 3 phases), max over ranks of the device-timed regions.

Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_convection.py [--size 257]
       (or plain `python scripts/bench_convection.py` on one GPU)
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=257)
    ap.add_argument("--stokes-iters", type=int, default=100)
    ap.add_argument("--thermal-iters", type=int, default=100)
    ap.add_argument("--steps", type=int, default=2, help="coupled time steps timed (after one warm-up step)")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from justrelax_jl_b200 import B200Backend, PhaseRatios, PTArray, StokesArrays, comm, setups, stokes as jst, thermal as jth
    from justrelax_jl_b200.stokes3d_vc import iterate3d_VC_
    from justrelax_jl_b200.types import IGG, ThermalArrays

    n = args.size
    igg = comm.init_global_grid(n, n, n) if world > 1 else IGG()
    s = setups.convection3d(n, n, n, igg=igg)
    dev = lambda a: PTArray(B200Backend)(a)
    st = StokesArrays(B200Backend, n, n, n, vertex_normals=False)
    th = ThermalArrays(B200Backend, n, n, n)
    th.T.copy_(dev(s.T)); th.Told.copy_(th.T)
    pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
    # args as the miniapp passes them (Blob3D.jl:280): ΔT = thermal.ΔT (ni .+ 2) switches compute_P! to the thermal-stress form
    a = dict(T=th.T, P=st.P, dt=s.dt, ΔT=th.ΔT)
    z = lambda: dev(np.zeros(s.ni, order="F"))
    ρg = (z(), z(), z())
    jst.flow_bcs_(st, s.flow_bcs)
    # lithostatic initial pressure from the buoyancy force (Blob3D.jl:296-299: compute_ρg! + init_P!, five fixed-point passes), on the GPU
    for _ in range(5):
        jst.compute_ρg_(ρg, pr, s.rheology, dict(T=th.T, P=st.P), st)
        jst.compute_lithostatic_pressure_(st.P, ρg[2], float(s.di[2]), igg if world > 1 else None)
    from justrelax_jl_b200 import zeros
    Vv = [zeros(B200Backend, n + 1, n + 1, n + 1) for _ in range(3)]
    pt_th = jth.PTThermalCoeffs(B200Backend, s.rheology, pr, a, s.dt, s.ni, s.di, s.li, ϵ=1e-5, CFL=s.thermal_CFL)
    kw_s = dict(viscosity_cutoff=s.kwargs["viscosity_cutoff"])
    kw_t = dict(phase=pr, verbose=False, igg=igg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        # Blob3D.jl:355-399: buoyancy + viscosity, Stokes solve, thermal solve (fixed iteration counts here)
        jst.compute_ρg_(ρg, pr, s.rheology, a, st)
        jst.compute_viscosity_(st, pr, a, s.rheology, s.kwargs["viscosity_cutoff"])
        rs = iterate3d_VC_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, a, s.dt, args.stokes_iters, igg, finish=True, kwargs=kw_s)
        # between the solves, still on the GPU: τII/εII, time step, shear heating (Χ = 0 for this setup: the call is what is timed),
        # vertex velocities for the advection step
        jst.tensor_invariant_(st.ε, s.ni)
        jst.compute_dt_(st, s.di, math.inf, igg if world > 1 else None)
        jst.compute_shear_heating_(th, st, pr, s.rheology, s.dt)
        jst.velocity2vertex_(Vv, (st.V.Vx, st.V.Vy, st.V.Vz), s.ni)
        rt = jth.thermal_iterate_(th, pt_th, s.thermal_bc, s.rheology, dict(T=th.T, P=st.P), s.dt, s.grid, args.thermal_iters, kwargs=kw_t)
        return rs, rt

    step()
    barrier()
    ts = tt = 0.0
    launches = 0
    for _ in range(args.steps):
        rs, rt = step()
        ts += rs.time
        tt += rt.time
        launches += rs.kernel_launches + rt.kernel_launches
    barrier()
    t = torch.tensor([ts, tt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts, tt = (float(v) for v in t.tolist())
    vmax = float(st.V.Vz.abs().max().item())
    Tfin = bool(torch.isfinite(th.T).all().item())
    if world > 1:
        comm.finalize_global_grid()
    if rank == 0:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        cells = n ** 3
        its, itt = args.steps * args.stokes_iters / ts, args.steps * args.thermal_iters / tt
        halo = 0
        if world > 1:   # received bytes per Stokes iteration on the busiest rank: ητ + τyz,τxz,τxy + Vx,Vy,Vz planes per neighbour
            exts = [(n, n, n), (n, n + 1, n + 1), (n + 1, n, n + 1), (n + 1, n + 1, n), (n + 1, n + 2, n + 2), (n + 2, n + 1, n + 2), (n + 2, n + 2, n + 1)]
            for e in exts:
                for d in range(3):
                    if igg.dims[d] > 1:
                        halo += (2 if igg.dims[d] > 2 else 1) * 8 * (e[0] * e[1] * e[2] // e[d])
        print(json.dumps(dict(
            workload=f"3D thermal convection, grid-based 3-phase ratios, coupled Stokes (3D-VC) + heatdiffusion_PT!, {n}^3 per GPU", n_gpus=world,
            decomposition="x".join(str(v) for v in igg.dims), global_grid=list(igg.n_g(s.ni)), coupled_steps=args.steps,
            stokes=dict(iters_per_s=its, ms_per_iter=1e3 / its, A_eff_bytes_per_cell=512, T_eff_GBs_per_gpu=512 * cells * its / 1e9,
                        T_eff_frac_of_measured_peak=512 * cells * its / 1e9 / peak, halo_bytes_per_iter_per_gpu=halo),
            thermal=dict(iters_per_s=itt, ms_per_iter=1e3 / itt, A_eff_bytes_per_cell=200, T_eff_GBs_per_gpu=200 * cells * itt / 1e9,
                         T_eff_frac_of_measured_peak=200 * cells * itt / 1e9 / peak),
            kernel_launches=launches, max_abs_Vz=vmax, T_finite=Tfin, scaling="weak", dtype="f64", data="synthetic")), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
