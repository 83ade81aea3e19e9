#!/bin/bash
cd /root/repo
JRB200_VA_BY=8 JRB200_VA_NCHUNK=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_va_tma -s 10 -c 2 -o gpurun_out/r1_tma_by8 -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_by8.log 2>&1
JRB200_VA_BY=10 JRB200_VA_NCHUNK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_va_tma -s 10 -c 1 -o gpurun_out/r1_tma_by10 -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_by10.log 2>&1
JRB200_VA_BY=8 JRB200_VA_NCHUNK=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches_v1.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_by8.log gpurun_out/ncu_by10.log
