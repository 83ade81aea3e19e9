#!/usr/bin/env python
"""Summarise an .ncu-rep (one or more kernels, `ncu --set full`) into the few counters DESIGN.md / bench.py quote.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "sm__cycles_elapsed.avg.per_second",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in out.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print(f"== {d['Kernel Name'][1]}  grid {d.get('Grid Size', ('', ''))[1]} block {d.get('Block Size', ('', ''))[1]}")
        for k in KEYS:
            if k in d:
                print(f"   {k:75s} {d[k][1]:>16s} {d[k][0]}")
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
            rd = float(d["dram__bytes_read.sum"][1]) * scale[d["dram__bytes_read.sum"][0]]
            wr = float(d["dram__bytes_write.sum"][1]) * scale[d["dram__bytes_write.sum"][0]]
            print(f"   {'dram traffic (read+write)':75s} {(rd + wr) / 1e9:16.6f} Gbyte")
            # stall reasons from the warp-state sampler (share of all samples)
            st = {k[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v[1]) for k, v in d.items()
                  if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v[1] not in ("", "n/a")}
            tot = sum(st.values())
            if tot > 0:
                top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
                print("   warp-state samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in top))
        except Exception:
            pass


if __name__ == "__main__":
    main(sys.argv[1])
