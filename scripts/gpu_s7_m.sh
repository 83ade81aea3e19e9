#!/bin/bash
# session 7, call m: fused thermal at 3/4/5 CTAs per SM; 2D-VC with branch-free reciprocals
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_stokes2d.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s7m_tests.log
for m in 3 4 5 3 4 5; do echo "TH_MINB=$m"; JRB200_TH_MINB=$m timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | tee -a gpurun_out/s7m_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
for kc in 8 12 24; do echo "TH_MINB=4 KCHUNK=$kc"; JRB200_TH_MINB=4 JRB200_TH_KCHUNK=$kc timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | grep -oE '"ms_per_step": [0-9.]+'; done
echo "2D"; timeout 300 python scripts/bench_paths.py --only solcx2d,shearband2d --steps 200 2>&1 | grep '^{' | tee -a gpurun_out/s7m_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'
