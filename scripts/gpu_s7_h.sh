#!/bin/bash
# session 7, call h: L2 prefetch of the next plane(s) — thermal fused kernel and the three 3D-VC kernels; parity + A/B
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_stokes3d_vc.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s7h_tests.log
for f in 1 0 1 0; do echo "TH_PREFETCH=$f"; JRB200_TH_PREFETCH=$f timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | tee -a gpurun_out/s7h_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
for f in 1 0 1 0; do echo "VC3_PREFETCH=$f"; JRB200_VC3_PREFETCH=$f timeout 300 python scripts/bench_paths.py --only vc3d --steps 60 2>&1 | grep '^{' | tee -a gpurun_out/s7h_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_th_|k_vc3|k_free_slip3' -c 60 --csv --log-file gpurun_out/s7h_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7h_l2.log 2>&1
