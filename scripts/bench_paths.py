#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configurations on ONE B200, one JSON line each (profiles/rNN_paths.jsonl).
bench.py stays the headline (config 4).  T_eff uses the PTsolvers convention with the A_eff of SURVEY.md §8d.

  solcx2d     config 2: 2D SolCx 511², variant 2D-V2                       A_eff = 120 B/cell
  shearband2d config 3: 2D shear band 1023², variant 2D-VC (2 phases)      A_eff = 280 B/cell
  vc3d        config 5 (Stokes half): 3D-VC, 3 phases, n³ per GPU          A_eff = 512 B/cell
  thermal3d   config 5 (thermal half): heatdiffusion_PT!, 3 phases, n³     A_eff = 200 B/cell
Usage: python scripts/bench_paths.py [--only name,...] [--steps K] [--n 257]
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
    ap.add_argument("--only", default="solcx2d,shearband2d,vc3d,thermal3d")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--n", type=int, default=257)
    args = ap.parse_args()
    import numpy as np
    import torch

    from justrelax_jl_b200 import B200Backend, PhaseRatios, PTArray, StokesArrays, setups, stokes as jst, thermal as jth
    from justrelax_jl_b200.stokes3d_vc import iterate3d_VC_
    from justrelax_jl_b200.types import ThermalArrays

    torch.cuda.set_device(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dev = lambda a: PTArray(B200Backend)(a)

    def emit(name, workload, cells, a_eff, r, extra=None):
        t = r.time / args.steps
        line = dict(workload=name, config=workload, steps=args.steps, ms_per_step=1e3 * t, iters_per_s=1 / t, cells=cells, A_eff_bytes_per_cell=a_eff,
                    T_eff_GBs=a_eff * cells / t / 1e9, T_eff_frac_of_measured_peak=a_eff * cells / t / 1e9 / peak, peak_GBs=peak,
                    kernel_launches_per_step=r.kernel_launches / args.steps)
        line.update(extra or {})
        print(json.dumps(line), flush=True)

    for name in args.only.split(","):
        if name == "solcx2d":
            s = setups.solcx2d(511, 511)
            st = StokesArrays(B200Backend, *s.ni)
            d = {k: dev(v) for k, v in s.fields.items()}
            st.viscosity.η.copy_(d["eta"])
            run = lambda k: jst.iterate2d_V2_(st, s.pt_stokes, s.grid, s.flow_bcs, (d["rhogx"], d["rhogy"]), d["G"], d["K"], s.dt, k)
            run(args.warmup)
            emit(name, "2D SolCx 511x511, variant 2D-V2, G=K=Inf, free slip (31 MB working set: L2-resident)", 511 * 511, 120, run(args.steps))
        elif name == "shearband2d":
            s = setups.shearband2d(1023)
            st = StokesArrays(B200Backend, *s.ni)
            st.V.Vx.copy_(dev(s.fields["Vx"])); st.V.Vy.copy_(dev(s.fields["Vy"]))
            pr = PhaseRatios.from_arrays(B200Backend, **s.ratios)
            T = dev(s.fields["T"])
            a = dict(T=T, P=st.P)
            jst.compute_viscosity_(st, pr, a, s.rheology, (-math.inf, math.inf))
            jst.flow_bcs_(st, s.flow_bcs)
            z = lambda: dev(np.zeros(s.ni, order="F"))
            ρg = (z(), z())
            run = lambda k: jst.iterate2d_VC_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, a, s.dt, k)
            run(args.warmup)
            emit(name, "2D shear band 1023x1023, variant 2D-VC, 2 phases, Drucker-Prager", 1023 * 1023, 280, run(args.steps))
        elif name in ("vc3d", "thermal3d"):
            n = args.n
            s = setups.convection3d(n, n, n)
            T = dev(s.T)
            if name == "vc3d":
                st = StokesArrays(B200Backend, n, n, n, vertex_normals=False)
                pr = PhaseRatios.from_arrays(B200Backend, **{k: v for k, v in s.ratios.items() if k in ("center", "xy", "yz", "xz")})
                a = dict(T=T, P=st.P)
                z = lambda: dev(np.zeros(s.ni, order="F"))
                ρg = (z(), z(), z())
                jst.flow_bcs_(st, s.flow_bcs)
                kw = dict(viscosity_cutoff=s.kwargs["viscosity_cutoff"])
                run = lambda k: iterate3d_VC_(st, s.pt_stokes, s.grid, s.flow_bcs, ρg, pr, s.rheology, a, s.dt, k, kwargs=kw)
                run(args.warmup)
                emit(name, f"3D convection Stokes {n}^3, variant 3D-VC, 3 phases (DP crust, blob, weak layer), PT_Density", n ** 3, 512, run(args.steps))
            else:
                th = ThermalArrays(B200Backend, n, n, n)
                th.T.copy_(T); th.Told.copy_(T)
                pr = PhaseRatios.from_arrays(B200Backend, **{k: v for k, v in s.ratios.items() if k in ("center", "Vx", "Vy", "Vz")})
                P = dev(np.zeros(s.ni, order="F"))
                a = dict(T=th.T, P=P)
                pt = jth.PTThermalCoeffs(B200Backend, s.rheology, pr, a, s.dt, s.ni, s.di, s.li, ϵ=1e-5, CFL=s.thermal_CFL)
                kw = dict(phase=pr, verbose=False)
                run = lambda k: jth.thermal_iterate_(th, pt, s.thermal_bc, s.rheology, a, s.dt, s.grid, k, kwargs=kw)
                run(args.warmup)
                emit(name, f"3D convection heatdiffusion_PT! {n}^3, rheology form with 3-phase ratios", n ** 3, 200, run(args.steps))
        else:
            raise SystemExit(f"unknown workload {name}")


if __name__ == "__main__":
    main()
