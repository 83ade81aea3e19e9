#!/bin/bash
# session 7, call w: early / non-coherent loads in 3D-VC prep + vel, lam hoisted in the stress centre, L1 prefetch of stage-2 operands in 2D-VC
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py tests/test_gpu_stokes2d.py tests/test_golden_fixtures.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s7w_tests.log
for r in 1 2; do timeout 300 python scripts/bench_paths.py --only vc3d,shearband2d --steps 100 2>&1 | grep '^{' | tee -a gpurun_out/s7w_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_vc3' -c 30 --csv --log-file gpurun_out/s7w_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7w_l2.log 2>&1
python - <<'P'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/s7w_launches_vc3d.csv')) if len(r)>5]
hdr=rows[0]; d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[hdr.index('Kernel Name')][:36]].append(float(r[-1]))
    except: pass
for k,v in d.items(): print(k, len(v), round(sum(v)/len(v)/1e3,1))
P
