#!/bin/bash
# session 7, call b: multi-iteration kernel with the boundary block out of the common path — parity, interleaved A/B, ncu
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s7b_tests3d.log
SWEEP_REPEAT=3 timeout 600 python scripts/bench_va_sweep.py JRB200_VA_MULTI=0 JRB200_VA_MULTI=1 JRB200_VA_MULTI=1,JRB200_VA_MULTI_MAX=99 2>&1 | tee gpurun_out/s7b_sweep.log | cut -c1-220
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/s7b_launches_multi.csv env SWEEP_STEPS=10 python scripts/bench_va_sweep.py JRB200_VA_MULTI=1 JRB200_VA_MULTI=0 > gpurun_out/s7b_l.log 2>&1
ls -la gpurun_out | tail -5
