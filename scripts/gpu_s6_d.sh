#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s6d_tests.log 2>&1; tail -6 gpurun_out/s6d_tests.log
timeout 600 python scripts/bench_paths.py --steps 100 > gpurun_out/s6d_paths.jsonl 2> gpurun_out/s6d_paths.err
cat gpurun_out/s6d_paths.jsonl; tail -5 gpurun_out/s6d_paths.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/s6d_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6d_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/s6d_launches_thermal3d.csv python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s6d_l2.log 2>&1
