#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 PT backend (BASELINE.json metric).

Metric : 3D Stokes PT iterations/s (and T_eff = A_eff / t_iter) on the SolVi inclusion setup,
         255^3 cells per GPU, Float64, variant 3D-VA (miniapps/benchmarks/stokes3D/solvi/SolVi3D.jl) —
         BASELINE.json configs[3], the configuration the metric is quoted on.
Step   : one PT iteration (one pass of the fused hot path over the whole local grid).
value  : whole-job block-iterations/s (PT iterations/s of one 255^3 block × number of blocks = GPUs; = iterations/s at N = 1)
         with all fields resident in HBM, measured INSIDE a running PT loop (jr_stokes3d_VA_begin → W warm-up iterations →
         K timed iterations → end): the steady-state iteration rate of `solve!`, CUDA events, max over ranks.  The cost of
         entering/leaving one solve (layout pack, observable last iteration) is reported beside it as `per_solve_overhead_ms`.
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
# capture of this same workload (profiles/r02_va_tma_ncu_full.txt: k_va_tma<10,0,0,3,0,0> at 255^3, constant body force: 1.675674 GB
# read + 1.293895 GB written).  bench.py cannot run ncu itself (a number taken under a profiler is never a bench value), so the
# figure is quoted for the configuration it was captured on and null for anything else.
NCU_TRAFFIC_BYTES = {(255, True): 2_969_569_000}
NCU_TRAFFIC_SOURCE = "profiles/r02_va_tma_ncu_full.txt (ncu --set full, one launch)"


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


def workload(n):
    return f"3D SolVi inclusion Stokes {n}^3 per GPU, variant 3D-VA (K,G arrays), dt=Inf, free slip"


METRIC = "3D Stokes PT block-iterations/s (SolVi3D, Float64; one block = the n^3 cells of one GPU, summed over GPUs)"


def all_cores():
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core it can (stated in `cores`)."""
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    return cores


def oracle_run(n, steps, warmup, budget_s=120.0, keep=None):
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
    if keep is not None:
        keep["fields"], keep["iters"] = d, warmup + done
    return done / t, t, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = all_cores()
    n = args.n
    ips, t, done = oracle_run(n, args.steps, min(args.warmup, 3))
    teff = A_EFF_BYTES_PER_CELL * n ** 3 * ips / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / ips, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload(n), "grid_per_gpu": [n, n, n],
                   "note": "the CPU processes the blocks of the weak-scaled problem one after the other: its block-iterations/s "
                           "does not depend on N"},
        "T_eff_GBs": teff,
        "cpu_baseline": {"value": ips, "unit": "iters/s", "cores": cores, "kind": "port",
                         "sample": f"{done} PT iterations at {n}^3 after {min(args.warmup, 3)} warm-up, 120 s cap (oracle/: C + OpenMP restatement "
                                   "of the reference's unfused kernel sequence; the Julia reference cannot run here)"},
        "e2e": {"value": ips, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def bind_to_gpu_numa(local):
    """pin this rank (and the pinned host buffers it allocates afterwards) to the CPU cores next to its GPU"""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


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
    numa_cpus = bind_to_gpu_numa(local) if world > 1 else None
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
    state_names = ["Vx", "Vy", "Vz", "P", "txx", "tyy", "tzz", "tyz", "txz", "txy"]

    def reset_state():
        for k in ("Vx", "Vy", "Vz", "eta"):
            st.slots()[k].copy_(dev[k])
        for k in ("P", "txx", "tyy", "tzz", "tyz", "txz", "txy"):
            st.slots()[k].zero_()
        jst.flow_bcs_(st, s.flow_bcs)
        if world > 1:
            comm.update_halo_(st.V.Vx, st.V.Vy, st.V.Vz, ni=(n, n, n))  # SolVi3D.jl:101

    reset_state()
    ρg = (dev["rhogx"], dev["rhogy"], dev["rhogz"])
    session = lambda: jst.IterationSession(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, dev["K"], dev["G"], s.dt, igg)
    run = lambda k: jst.iterate_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, dev["K"], dev["G"], s.dt, k, igg)
    nout = int(s.kwargs["nout"])  # the reference's residual sampling interval of this miniapp (SolVi3D.jl: nout = 100)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_iterations(it, K):
        """K PT iterations of the open session, sampled like the reference's loop: every `nout`-th iteration is an observable one
        followed by the four residual norms (Stokes3D.jl:125-142).  Device time: CUDA events on the library's stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, done = 0, 0
        e0.record()
        while done < K:
            k = min(K - done, nout - (done % nout))
            sample = (done + k) % nout == 0
            r = it.step(k, observe_last=sample)
            launches += r.kernel_launches
            done += k
            if sample:
                jst.residual_norms3d_(st, igg)
                launches += 4
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) * 1e-3, launches

    # ---- warm-up, then time exactly K steps on the device, inside one running PT loop ----
    W = max(args.warmup, 3)
    sampler = ClockSampler(local)
    with session() as it:
        it.step(W)
        barrier()
        if rank == 0:
            sampler.start()
        barrier()
        t_loc, launches = timed_iterations(it, args.steps)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
    t = max_over_ranks(t_loc)
    ips = args.steps / t
    cells = n ** 3
    info = jst.plan_info()
    a_eff = A_EFF_CONST_RHOG if info["rhog_const"] else A_EFF_BYTES_PER_CELL

    # the same K steps with the body-force arrays streamed (what a setup with spatially varying ρg costs;
    # A_eff = 192 B/cell, the PTsolvers convention of SURVEY.md §8d)
    os.environ["JRB200_VA_STREAM_RHOG"] = "1"
    with session() as it:
        it.step(W)
        barrier()
        t_s, _ = timed_iterations(it, args.steps)
    del os.environ["JRB200_VA_STREAM_RHOG"]
    ips_streamed = args.steps / max_over_ranks(t_s)

    # what ONE solve-like call of K iterations costs on top of its iterations: entering the TMA box layout (pack), the observable
    # last iteration (diagnostics + dense state), leaving.  A real solve! pays this once per ~10^3 iterations.
    run(3)
    barrier()
    r_call = run(args.steps)
    t_call = max_over_ranks(r_call.time)
    per_solve_overhead_ms = max(0.0, (t_call - t) * 1e3)

    # optional: the reference's kernel split on the same GPU (unfused CUDA path of this library)
    unfused_ips = None
    if args.unfused:
        jst.set_flags(_abi.JR_FLAG_UNFUSED)
        run(3)
        ru = run(max(args.steps // 4, 5))
        jst.set_flags(0)
        unfused_ips = ru.iter / ru.time

    # ---- e2e: host buffers -> upload -> K iterations -> download ----
    # arrays the setup declares spatially constant (SolVi: ρg = 0, K = Inf, G = 1, P = τ = 0) are filled on the device, as the
    # reference's @zeros / @fill do; only the arrays that carry data cross PCIe
    is_const = {k: bool(np.all(v == v.flat[0])) for k, v in s.fields.items()}
    host_in = {k: torch.from_numpy(np.ascontiguousarray(v.T)).pin_memory() for k, v in s.fields.items() if not is_const[k]}
    const_in = {k: float(v.flat[0]) for k, v in s.fields.items() if is_const[k]}
    host_out = {k: torch.empty(tuple(reversed(st.slots()[k].shape)), dtype=torch.float64).pin_memory() for k in state_names}
    cview = lambda a: a.permute(*reversed(range(a.dim())))  # contiguous reversed-dims view of a column-major device array

    def e2e_once():
        for k, h in host_in.items():
            dst = st.slots()[k] if k in ("Vx", "Vy", "Vz", "eta") else dev[k]
            cview(dst).copy_(h, non_blocking=True)
        for k, v in const_in.items():
            dev[k].fill_(v)
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
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_ips = args.steps / t_e2e
    h2d = sum(h.numel() * 8 for h in host_in.values())
    d2h = sum(h.numel() * 8 for h in host_out.values())

    # ---- parity of this very workload against the CPU oracle (N = 1): the same M iterations from the same initial state ----
    parity = None
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = all_cores()
        keep = {}
        cpu_ips, cpu_t, cpu_done = oracle_run(n, args.cpu_steps, 1, budget_s=30.0, keep=keep)
        cpu = {"value": cpu_ips, "unit": "iters/s", "cores": cores, "kind": "port",
               "sample": f"{cpu_done} PT iterations at {n}^3 (+1 warm-up), oracle/ C+OpenMP restatement"}
        reset_state()
        with session() as it:          # same plan as the timed region; the last iteration is observable
            it.step(keep["iters"], observe_last=True)
        from util import max_rel_diff
        worst = {k: max_rel_diff(to_host(st.slots()[k]), keep["fields"][k]) for k in state_names + ["Rx", "Ry", "Rz", "RP"]}
        parity = {"max_rel": max(worst.values()), "iterations": keep["iters"], "fields": len(worst), "tolerance": 1e-12,
                  "worst_field": max(worst, key=worst.get)}

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
        "metric": METRIC, "value": ips * world, "unit": "iters/s",
        "iters_per_s_per_gpu": ips, "cell_updates_per_s": ips * cells * world,
        "n_gpus": world, "steps": args.steps, "warmup": W, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload(n),
                   "grid_per_gpu": [n, n, n], "l2": f"working set {25 * 8 * cells / 1e9:.2f} GB/iteration >> 126 MB L2 (no flush needed)",
                   "decomposition": (f"IGG-compatible {igg.dims[0]}x{igg.dims[1]}x{igg.dims[2]} block decomposition, overlap 2, global grid "
                                     f"{'x'.join(str(v) for v in igg.n_g((n, n, n)))}, V halos exchanged every iteration (CUDA-IPC peer memory over NVLink)"
                                     if world > 1 else "single block"),
                   "halo_bytes_per_iter_per_gpu": halo_bytes,
                   "timed_region": f"iterations {W + 1}..{W + args.steps} of one running PT loop; residual norms every {nout} iterations as in the miniapp",
                   "plan": info},
        "T_eff_GBs_per_gpu": achieved, "T_eff_GBs_total": achieved * world, "T_eff_frac_of_8TBs": achieved / 8000.0, "A_eff_bytes_per_cell": a_eff,
        "streamed_rhog": {"value": ips_streamed * world, "unit": "iters/s", "A_eff_bytes_per_cell": A_EFF_BYTES_PER_CELL,
                          "T_eff_GBs_per_gpu": achieved_streamed, "T_eff_frac_of_measured_peak": achieved_streamed / peak,
                          "T_eff_frac_of_8TBs": achieved_streamed / 8000.0,
                          "note": "same run with JRB200_VA_STREAM_RHOG=1: the three (zero) body-force arrays are read every iteration"},
        "per_solve_overhead_ms": per_solve_overhead_ms,
        "solve_call": {"value": args.steps / t_call * world, "unit": "iters/s",
                       "note": f"one jr_stokes3d_iterate_VA call of {args.steps} iterations incl. layout entry/exit and the observable last iteration"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_TRAFFIC_BYTES.get((n, bool(info["rhog_const"]))),
                     "traffic_source": NCU_TRAFFIC_SOURCE, "peak_kind": peak_kind, "kernel": "k_va_tma (TMA-staged fused 3D-VA iteration)",
                     "algorithmic_bytes_per_launch": a_eff * cells},
        "e2e": {"value": e2e_ips * world, "unit": "iters/s", "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                "note": f"one solve of {args.steps} PT iterations incl. upload of the {len(host_in)} non-constant input arrays "
                        f"(constants {sorted(const_in)} are filled on the device) and download of V,P,τ",
                "numa_bound_cpus": numa_cpus},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if parity is not None:
        line["parity_max_rel"] = parity["max_rel"]
        line["parity"] = parity
    if unfused_ips is not None:
        line["reference_kernel_split_on_gpu_iters_s"] = unfused_ips
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not (parity["max_rel"] <= parity["tolerance"]):
        raise SystemExit(f"bench.py: parity against the oracle FAILED: {parity}")


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
