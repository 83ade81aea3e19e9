#!/bin/bash
cd /root/repo
export JRB200_VERBOSE=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1_tests_v1.log 2>&1
echo "tests rc=$?" >> gpurun_out/r1_tests_v1.log
tail -5 gpurun_out/r1_tests_v1.log
for cfg in "10 1" "8 3" "8 4" "16 2" "10 2" "8 2"; do
  set -- $cfg
  echo "=== BY=$1 NCHUNK=$2" >> gpurun_out/r1_bench_v1.log
  JRB200_VA_BY=$1 JRB200_VA_NCHUNK=$2 timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu >> gpurun_out/r1_bench_v1.log 2>&1
done
grep -E "===|value" gpurun_out/r1_bench_v1.log | sed -E 's/.*"value": ([0-9.]+).*"T_eff_GBs_per_gpu": ([0-9.]+).*/ips=\1 Teff=\2/'
