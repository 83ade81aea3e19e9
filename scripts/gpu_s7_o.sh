#!/bin/bash
# session 7, call o: 2D-V2 kernel at 3 CTAs/SM; final ncu captures of the fused thermal kernel and the 3D-VC kernels at HEAD
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes2d.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s7o_tests.log
for r in 1 2; do timeout 300 python scripts/bench_paths.py --only solcx2d --steps 400 2>&1 | grep '^{' | tee -a gpurun_out/s7o_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_th_fused3' -s 4 -c 1 -o gpurun_out/s7o_thermal_fused -f python scripts/bench_paths.py --only thermal3d --steps 12 --warmup 2 > gpurun_out/s7o_n1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_th_' -c 40 --csv --log-file gpurun_out/s7o_launches_thermal3d.csv python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s7o_l1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_' -s 9 -c 3 -o gpurun_out/s7o_vc3d -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7o_n2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_vc3|k_free_slip3' -c 40 --csv --log-file gpurun_out/s7o_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7o_l2.log 2>&1
ls gpurun_out | grep s7o
