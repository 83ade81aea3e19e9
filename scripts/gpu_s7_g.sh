#!/bin/bash
# session 7, call g: fused thermal kernel v2 (compile-time phase bound, loads issued before first use)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermal.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s7g_tests.log
for f in 1 0 1 0; do echo "FUSED=$f"; JRB200_TH_FUSED=$f timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | tee -a gpurun_out/s7g_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
for kc in 8 16 22; do echo "KCHUNK=$kc"; JRB200_TH_KCHUNK=$kc timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | grep -oE '"ms_per_step": [0-9.]+'; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_th_fused3' -s 4 -c 1 -o gpurun_out/s7g_thermal_fused -f python scripts/bench_paths.py --only thermal3d --steps 12 --warmup 2 > gpurun_out/s7g_n1.log 2>&1
