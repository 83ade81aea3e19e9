#!/bin/bash
# session 7, call a: multi-iteration persistent 3D-VA kernel — parity, sweep, bench, launch list, full ncu set
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_stokes3d.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/s7a_tests3d.log
timeout 600 python scripts/bench_va_sweep.py JRB200_VA_MULTI=0 JRB200_VA_MULTI=1 JRB200_VA_MULTI=1,JRB200_VA_SLACK=0 JRB200_VA_MULTI=1,JRB200_VA_SLACK=2 JRB200_VA_MULTI=1,JRB200_VA_SLACK=4 JRB200_VA_MULTI=1,JRB200_VA_STREAM_RHOG=1 JRB200_VA_MULTI=0,JRB200_VA_STREAM_RHOG=1 JRB200_VA_MULTI=1,JRB200_VA_MULTI_MAX=99 2>&1 | tee gpurun_out/s7a_sweep.log | cut -c1-200
timeout 600 python bench.py --steps 200 --warmup 10 2>&1 | tail -1 | tee gpurun_out/s7a_bench_n1.json | cut -c1-400
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_stokes3d.py 2>&1 | tail -5 | tee gpurun_out/s7a_tests_rest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/s7a_launches_multi.csv python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/s7a_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_va_tma' -s 2 -c 1 -o gpurun_out/s7a_va_multi -f env SWEEP_STEPS=10 python scripts/bench_va_sweep.py > gpurun_out/s7a_n.log 2>&1
ls -la gpurun_out | tail -8
