#!/bin/bash
# session 7, call s: 2D kernel tile heights (wave quantisation): parity for every height, timing of configs 2 and 3
cd /root/repo
mkdir -p gpurun_out
for t in 0 12 18 20; do echo "TY=$t tests"; JRB200_2D_TY=$t timeout 600 python -m pytest tests/test_gpu_stokes2d.py -x -q -m gpu 2>&1 | tail -2; done | tee gpurun_out/s7s_tests.log
for t in 0 12 14 16 18 20; do echo "TY=$t"; JRB200_VERBOSE=1 JRB200_2D_TY=$t timeout 300 python scripts/bench_paths.py --only solcx2d,shearband2d --steps 300 2>&1 | grep -E '^\{|tile height' | tee -a gpurun_out/s7s_paths.log | grep -oE '"ms_per_step": [0-9.]+|tile height [0-9]+' | sort -u; done
