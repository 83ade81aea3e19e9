#!/bin/bash
# session 7, call x: 2D-V2 stage-2 prefetch; ncu capture of the 3D-VC kernels at HEAD
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes2d.py -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/s7x_tests.log
for r in 1 2; do timeout 300 python scripts/bench_paths.py --only solcx2d --steps 400 2>&1 | grep '^{' | tee -a gpurun_out/s7x_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_' -s 9 -c 3 -o gpurun_out/s7x_vc3d -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7x_n2.log 2>&1
