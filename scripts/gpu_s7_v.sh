#!/bin/bash
# session 7, call v: staged 3D-VC stress kernel, centre first with its global operands loaded behind the tile copies
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py tests/test_golden_fixtures.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s7v_tests.log
for r in 1 2; do timeout 300 python scripts/bench_paths.py --only vc3d --steps 100 2>&1 | grep '^{' | tee -a gpurun_out/s7v_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_vc3' -c 30 --csv --log-file gpurun_out/s7v_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7v_l2.log 2>&1
grep -E "k_vc3_stress_sm<0" gpurun_out/s7v_launches_vc3d.csv | head -3 | awk -F, '{print $NF}'
