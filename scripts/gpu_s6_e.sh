#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py tests/test_golden_fixtures.py -x -q -m gpu > gpurun_out/s6e_tests.log 2>&1; tail -5 gpurun_out/s6e_tests.log
timeout 600 python scripts/bench_paths.py --only vc3d --steps 60 > gpurun_out/s6e_paths.jsonl 2> gpurun_out/s6e_paths.err
cat gpurun_out/s6e_paths.jsonl; tail -5 gpurun_out/s6e_paths.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_stress' -s 6 -c 1 -o gpurun_out/s6e_vc3d_stress -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6e_n1.log 2>&1
