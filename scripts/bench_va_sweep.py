#!/usr/bin/env python
"""Tuning sweep of the fused 3D-VA kernel at 255^3 (SolVi3D): ms per PT iteration for a list of env settings
(JRB200_VA_* knobs are read at every solve entry, so one process can time them all).  Not a bench line."""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from justrelax_jl_b200 import B200Backend, PTArray, StokesArrays, setups, stokes as jst

    n = int(os.environ.get("SWEEP_N", "255"))
    steps = int(os.environ.get("SWEEP_STEPS", "200"))
    s = setups.solvi3d(n, n, n)
    st = StokesArrays(B200Backend, n, n, n, vertex_normals=False)
    dev = {k: PTArray(B200Backend)(v) for k, v in s.fields.items()}
    for k in ("Vx", "Vy", "Vz", "eta"):
        st.slots()[k].copy_(dev[k])
    jst.flow_bcs_(st, s.flow_bcs)
    ρg = (dev["rhogx"], dev["rhogy"], dev["rhogz"])
    class _R:
        pass

    def run(k):
        # steady state: k iterations inside a running loop (after 5 warm-up iterations of the same session)
        with jst.IterationSession(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, dev["K"], dev["G"], s.dt) as it:
            it.step(5)
            return it.step(k)

    sleep_s = float(os.environ.get("SWEEP_SLEEP", "0"))
    settings = [dict(kv.split("=") for kv in item.split(",") if kv) for item in sys.argv[1:]] or [{}]
    # SWEEP_REPEAT > 1 cycles through the settings several times (A/B/A/B …): sustained load moves the SM clock under the
    # power cap, so only interleaved repeats compare fairly
    settings = settings * int(os.environ.get("SWEEP_REPEAT", "1"))
    try:
        import pynvml
        pynvml.nvmlInit()
        nv = pynvml.nvmlDeviceGetHandleByIndex(0)
    except Exception:
        nv = None
    for env in settings:
        for k, v in env.items():
            os.environ[k] = v
        run(5)
        best = None
        clks, pws, stop = [], [], [False]

        def sample():
            while not stop[0] and nv is not None:
                clks.append(pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM))
                pws.append(pynvml.nvmlDeviceGetPowerUsage(nv) / 1e3)
                time.sleep(0.004)

        th = threading.Thread(target=sample, daemon=True)
        th.start()
        for _ in range(3):
            if sleep_s:
                time.sleep(sleep_s)   # let the board's power average (and the SM clock) recover: the short-run regime of the driver's bench
            r = run(steps)
            best = r.time if best is None else min(best, r.time)
        torch.cuda.synchronize()
        stop[0] = True
        th.join()
        clk = sorted(clks)[len(clks) // 2] if clks else None
        pw = sorted(pws)[len(pws) // 2] if pws else None
        print(json.dumps({"env": env, "sm_mhz_median": clk, "power_w_median": pw, "ms_per_iter": 1e3 * best / steps, "iters_per_s": steps / best, "launches": r.kernel_launches,
                          "plan": jst.plan_info()}), flush=True)
        for k in env:
            del os.environ[k]


if __name__ == "__main__":
    main()
