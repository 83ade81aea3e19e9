#!/usr/bin/env python
"""Warp-state samples and instructions executed per CUDA source line of the kernel(s) in an .ncu-rep captured with
`ncu --set full --import-source on` (objects built with -lineinfo).
usage: python scripts/ncu_hotspots.py gpurun_out/x.ncu-rep [top] > profiles/x_source_hotspots.txt
       (or pass the CSV written by `ncu -i x.ncu-rep --page source --print-source cuda,sass --csv`)"""
import csv
import io
import subprocess
import sys


def main(path, top=40):
    if path.endswith(".csv"):
        text = open(path).read()
    else:
        text = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, cur, agg = None, None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 4 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < 8:
            continue
        line, src, addr = r[0], r[1], r[2]
        if not line or addr != "-":   # per-line rows carry a line number and no SASS address
            continue
        try:
            inst, samp = int(float(r[7])), int(float(r[4]))
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[4:]))
        a = agg.setdefault((cur, int(line)), [0, 0, src.strip()[:105], {}])
        a[0] += inst
        a[1] += samp
        for k, v in d.items():
            if k.startswith("stall_") and "(" not in k and v not in ("-", ""):
                a[3][k] = a[3].get(k, 0) + int(float(v))
    tot = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"# {path}: warp instructions executed {tot}, warp-state samples {ts}; top {top} source lines by samples")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        tp = sorted(v[3].items(), key=lambda kv: -kv[1])[:2]
        print(f"{100 * v[1] / ts:6.2f}% samp {100 * v[0] / tot:6.2f}% inst  {k[0]}:{k[1]:4d}  {v[2]}   [{', '.join(f'{a[6:]} {b}' for a, b in tp)}]")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
