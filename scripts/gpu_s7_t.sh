#!/bin/bash
# session 7, call t: remainder-column packing (nx mod 32 small) in the fused thermal kernel and the three 3D-VC kernels
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_stokes3d_vc.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s7t_tests.log
for f in 1 0 1 0; do echo "TH_PACK=$f"; JRB200_TH_PACK=$f timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | tee -a gpurun_out/s7t_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
for f in 1 0 1 0; do echo "VC3_PACK=$f"; JRB200_VC3_PACK=$f timeout 300 python scripts/bench_paths.py --only vc3d --steps 60 2>&1 | grep '^{' | tee -a gpurun_out/s7t_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
