#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 PT backend (BASELINE.json metric).

Metric : 3D Stokes PT iterations/s (and T_eff = A_eff / t_iter) on the SolVi inclusion setup,
         255^3 cells per GPU, Float64, variant 3D-VA (miniapps/benchmarks/stokes3D/solvi/SolVi3D.jl) —
         BASELINE.json configs[3], the configuration the metric is quoted on.
Step   : one PT iteration (one pass of the fused hot path over the whole local grid).
value  : whole-job iterations/s with all fields resident in HBM (CUDA events, max over ranks).
e2e    : the same metric through the public API with HOST buffers: per measurement the inputs
         (η, ρg, K, G, V, P, τ) are uploaded from pinned host memory, `steps` PT iterations run, and the
         solution (V, P, τ) is read back — what one `solve!` call costs a user whose data lives on the host.
--impl reference : the CPU restatement of the reference (oracle/, OpenMP on all host cores, same kernel
         split and array traffic as the reference's ParallelStencil-Threads backend), same config/metric.

Usage: python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n 255]
       (N > 1: launched by torch.distributed.run, one rank per GPU.)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

A_EFF_BYTES_PER_CELL = 192  # (2*D_u + D_k)*8 with D_u=10 (V×3,P,τ×6), D_k=4 (η,ρg×3) — SURVEY.md §8d config 4
A_EFF_CONST_RHOG = 168      # the same with D_k=1: the library does not stream spatially constant ρg (SolVi: ρg ≡ 0)


# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the dominant kernel from the committed `ncu --set full`
# capture of this same workload (profiles/r01_va_tma_ncu_full.txt: k_va_tma<10,0,0,3,0> at 255^3, constant body force: 1.674083 GB
# read + 1.293379 GB written).  bench.py cannot run ncu itself (a number taken under a profiler is never a bench value), so the
# figure is quoted for the configuration it was captured on and null for anything else.
NCU_TRAFFIC_BYTES = {(255, True): 2_967_462_000}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_run(n, steps, warmup, budget_s=120.0):
    """Time the CPU restatement (reference kernel split, OpenMP) for `steps` PT iterations at n^3."""
    import numpy as np
    from justrelax_jl_b200 import setups
    from oracle import pyoracle as po
    from util import bc_flags

    po.build()
    s = setups.solvi3d(n, n, n)
    d = po.alloc_stokes(s.ni, s.fields)
    opts = po.make_opts(s.pt_stokes, s.grid._di.center, s.dt, bc_flags(s.flow_bcs), s.ni, iterMax=10 ** 9, nout=10 ** 9)
    fs = po.make_fields(d, s.ni)
    po.lib().orc_flow_bcs3(C.byref(fs), C.byref(opts), 0)
    if warmup:
        po.iterate3d_VA(d, s.ni, opts, warmup)
    # bounded: stop after `budget_s` seconds even if fewer than `steps` iterations were timed
    done, t0 = 0, time.perf_counter()
    while done < steps:
        po.iterate3d_VA(d, s.ni, opts, 1)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    t = time.perf_counter() - t0
    return done / t, t, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    n = args.n
    ips, t, done = oracle_run(n, args.steps, min(args.warmup, 3))
    teff = A_EFF_BYTES_PER_CELL * n ** 3 * ips / 1e9
    line = {
        "impl": "reference", "metric": "3D Stokes PT iterations/s (SolVi3D, Float64)", "value": ips, "unit": "iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / ips, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D SolVi inclusion Stokes {n}^3, variant 3D-VA (K,G arrays), dt=Inf, free slip",
                   "grid": [n, n, n]},
        "T_eff_GBs": teff,
        "cpu_baseline": {"value": ips, "unit": "iters/s", "cores": int(os.environ["OMP_NUM_THREADS"]), "kind": "port",
                         "sample": f"{done} PT iterations at {n}^3 after {min(args.warmup, 3)} warm-up, 120 s cap (oracle/: C + OpenMP restatement "
                                   "of the reference's unfused kernel sequence; the Julia reference cannot run here)"},
        "e2e": {"value": ips, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 backend has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from justrelax_jl_b200 import B200Backend, PTArray, StokesArrays, setups, stokes as jst, to_host
    from justrelax_jl_b200 import _abi

    n = args.n
    # weak scaling: every rank owns an n^3 block of ONE global SolVi problem, decomposed like ImplicitGlobalGrid
    # (overlap 2, dims from MPI_Dims_create: 2x1x1, 2x2x1, 2x2x2), halos exchanged every iteration over NVLink
    igg = None
    halo_fn = None
    if world > 1:
        from justrelax_jl_b200 import comm
        igg = comm.init_global_grid(n, n, n)

        def halo_fn(h):  # update_halo!(η) between the smoothing passes of the setup (SolVi3D.jl:38-42)
            d = PTArray(B200Backend)(h)
            comm.update_halo_(d, ni=(n, n, n))
            h[...] = to_host(d)
    s = setups.solvi3d(n, n, n, igg=igg, update_halo=halo_fn)
    st = StokesArrays(B200Backend, n, n, n, vertex_normals=False)
    dev = {k: PTArray(B200Backend)(v) for k, v in s.fields.items()}
    for k in ("Vx", "Vy", "Vz", "eta"):
        st.slots()[k].copy_(dev[k])
    jst.flow_bcs_(st, s.flow_bcs)
    if world > 1:
        comm.update_halo_(st.V.Vx, st.V.Vy, st.V.Vz, ni=(n, n, n))  # SolVi3D.jl:101
    ρg = (dev["rhogx"], dev["rhogy"], dev["rhogz"])
    run = lambda k: jst.iterate_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, dev["K"], dev["G"], s.dt, k, igg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then time exactly K steps on the device (events inside the library, on its stream) ----
    run(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    r = run(args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_dev = torch.tensor([r.time], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    t = float(t_dev.item())
    ips = args.steps / t
    cells = n ** 3
    info = jst.plan_info()
    a_eff = A_EFF_CONST_RHOG if info["rhog_const"] else A_EFF_BYTES_PER_CELL

    # the same K steps with the body-force arrays streamed (what a setup with spatially varying ρg costs;
    # A_eff = 192 B/cell, the PTsolvers convention of SURVEY.md §8d)
    os.environ["JRB200_VA_STREAM_RHOG"] = "1"
    run(3)
    barrier()
    r_s = run(args.steps)
    del os.environ["JRB200_VA_STREAM_RHOG"]
    t_s = torch.tensor([r_s.time], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
    ips_streamed = args.steps / float(t_s.item())

    # optional: the reference's kernel split on the same GPU (unfused CUDA path of this library)
    unfused_ips = None
    if args.unfused:
        jst.set_flags(_abi.JR_FLAG_UNFUSED)
        run(3)
        ru = run(max(args.steps // 4, 5))
        jst.set_flags(0)
        unfused_ips = ru.iter / ru.time

    # ---- e2e: host buffers -> upload -> K iterations -> download ----
    host_in = {k: torch.from_numpy(np.ascontiguousarray(v.T)).pin_memory() for k, v in s.fields.items()}
    state_names = ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy"]
    host_out = {k: torch.empty(tuple(reversed(st.slots()[k].shape)), dtype=torch.float64).pin_memory() for k in state_names}
    cview = lambda a: a.permute(*reversed(range(a.dim())))  # contiguous reversed-dims view of a column-major device array

    def e2e_once():
        for k, h in host_in.items():
            dst = st.slots()[k] if k in ("Vx", "Vy", "Vz", "eta") else dev[k]
            cview(dst).copy_(h, non_blocking=True)
        for k in ("P", "txx", "tyy", "tzz", "tyz", "txz", "txy"):
            st.slots()[k].zero_()
        jst.flow_bcs_(st, s.flow_bcs)
        if world > 1:
            comm.update_halo_(st.V.Vx, st.V.Vy, st.V.Vz, ni=(n, n, n))
        run(args.steps)
        for k in state_names:
            host_out[k].copy_(cview(st.slots()[k]), non_blocking=True)
        torch.cuda.synchronize()

    e2e_once()
    barrier()
    t0 = time.perf_counter()
    e2e_once()
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ips = args.steps / float(t_e2e.item())
    h2d = sum(h.numel() * 8 for h in host_in.values())
    d2h = sum(h.numel() * 8 for h in host_out.values())

    if world > 1:
        comm.finalize_global_grid()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # received halo bytes per iteration on the busiest rank: per exchanged dimension one plane per neighbour of Vx, Vy, Vz
    halo_bytes = 0
    if world > 1:
        ext = [(n + 1, n + 2, n + 2), (n + 2, n + 1, n + 2), (n + 2, n + 2, n + 1)]
        for e in ext:
            for d in range(3):
                if igg.dims[d] > 1:
                    halo_bytes += (2 if igg.dims[d] > 2 else 1) * 8 * (e[0] * e[1] * e[2] // e[d])
    peak, peak_kind = peaks()
    achieved = a_eff * cells / (t / args.steps) / 1e9
    achieved_streamed = A_EFF_BYTES_PER_CELL * cells * ips_streamed / 1e9
    line = {
        "metric": "3D Stokes PT iterations/s (SolVi3D, Float64)", "value": ips * 1.0, "unit": "iters/s",
        "cell_updates_per_s": ips * cells * world,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"3D SolVi inclusion Stokes {n}^3 per GPU, variant 3D-VA (K,G arrays), dt=Inf, free slip",
                   "grid_per_gpu": [n, n, n], "l2": f"working set {25 * 8 * cells / 1e9:.2f} GB/iteration >> 126 MB L2 (no flush needed)",
                   "decomposition": (f"IGG-compatible {igg.dims[0]}x{igg.dims[1]}x{igg.dims[2]} block decomposition, overlap 2, global grid "
                                     f"{'x'.join(str(v) for v in igg.n_g((n, n, n)))}, V halos exchanged every iteration (CUDA-IPC pull over NVLink)"
                                     if world > 1 else "single block"),
                   "halo_bytes_per_iter_per_gpu": halo_bytes,
                   "plan": info},
        "T_eff_GBs_per_gpu": achieved, "T_eff_GBs_total": achieved * world, "T_eff_frac_of_8TBs": achieved / 8000.0, "A_eff_bytes_per_cell": a_eff,
        "streamed_rhog": {"value": ips_streamed, "unit": "iters/s", "A_eff_bytes_per_cell": A_EFF_BYTES_PER_CELL,
                          "T_eff_GBs_per_gpu": achieved_streamed, "T_eff_frac_of_measured_peak": achieved_streamed / peak,
                          "T_eff_frac_of_8TBs": achieved_streamed / 8000.0,
                          "note": "same run with JRB200_VA_STREAM_RHOG=1: the three (zero) body-force arrays are read every iteration"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES.get((n, bool(info["rhog_const"]))),
                     "traffic_source": "profiles/r01_va_tma_ncu_full.txt (ncu --set full, one launch)", "peak_kind": peak_kind, "kernel": "k_va_tma (TMA-staged fused 3D-VA iteration)",
                     "algorithmic_bytes_per_launch": a_eff * cells},
        "e2e": {"value": e2e_ips, "unit": "iters/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                "note": f"one solve of {args.steps} PT iterations incl. upload of 12 input arrays and download of V,P,τ"},
        "gpu_launches": int(r.kernel_launches),
        "clocks": clocks,
    }
    if unfused_ips is not None:
        line["reference_kernel_split_on_gpu_iters_s"] = unfused_ips
    # CPU baseline (N=1 only): bounded sample of the same workload on the host cores
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        cpu_ips, cpu_t, cpu_done = oracle_run(n, args.cpu_steps, 1, budget_s=30.0)
        line["cpu_baseline"] = {"value": cpu_ips, "unit": "iters/s", "cores": int(os.environ["OMP_NUM_THREADS"]), "kind": "port",
                                "sample": f"{cpu_done} PT iterations at {n}^3 (+1 warm-up), oracle/ C+OpenMP restatement"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=255)
    ap.add_argument("--cpu-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="also time the reference-structured (unfused) CUDA path")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
