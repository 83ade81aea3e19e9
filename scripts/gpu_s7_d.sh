#!/bin/bash
# session 7, call d: where does the multi-iteration launch lose time?  (timing experiments, interleaved)
cd /root/repo
mkdir -p gpurun_out
SWEEP_REPEAT=2 timeout 900 python scripts/bench_va_sweep.py JRB200_VA_MULTI=0 JRB200_VA_MULTI=1 JRB200_VA_MULTI=1,JRB200_VA_MULTI_MAX=1 JRB200_VA_MULTI=1,JRB200_VA_DBG_NOBC=1 JRB200_VA_MULTI=1,JRB200_VA_DBG_NOBAR=1 JRB200_VA_MULTI=1,JRB200_VA_DBG_NOBC=1,JRB200_VA_DBG_NOBAR=1 2>&1 | tee gpurun_out/s7d_sweep.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_va_tma|k_bc_box3' -c 24 --csv --log-file gpurun_out/s7d_launches_multi.csv env SWEEP_STEPS=10 python scripts/bench_va_sweep.py JRB200_VA_MULTI=1 JRB200_VA_MULTI=0 > gpurun_out/s7d_l.log 2>&1
