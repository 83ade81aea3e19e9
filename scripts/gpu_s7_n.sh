#!/bin/bash
# session 7, call n: staged 3D-VC stress kernel with 4-row CTAs (4 CTAs of 128 threads per SM instead of 2 of 256)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s7n_tests.log
for r in 1 2; do timeout 300 python scripts/bench_paths.py --only vc3d --steps 60 2>&1 | grep '^{' | tee -a gpurun_out/s7n_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_vc3' -c 30 --csv --log-file gpurun_out/s7n_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7n_l2.log 2>&1
grep -E "k_vc3_stress_sm<0" gpurun_out/s7n_launches_vc3d.csv | head -3 | awk -F, '{print $NF}'
