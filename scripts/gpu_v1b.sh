#!/bin/bash
cd /root/repo
export JRB200_VERBOSE=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r1_tests_v1b.log 2>&1
echo "tests rc=$?" >> gpurun_out/r1_tests_v1b.log
tail -4 gpurun_out/r1_tests_v1b.log
rm -f gpurun_out/r1_bench_v1b.log
run() {
  echo "=== $*" >> gpurun_out/r1_bench_v1b.log
  env "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu >> gpurun_out/r1_bench_v1b.log 2>&1
}
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=3
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=3 JRB200_VA_POL_ST=0
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=3 JRB200_VA_POL_LD=1
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=3 JRB200_VA_POL_LD=2
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=3 JRB200_VA_L2PROMO=0
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=3 JRB200_VA_L2PROMO=2
run JRB200_VA_BY=10 JRB200_VA_NCHUNK=1
run JRB200_VA_BY=10 JRB200_VA_NCHUNK=1 JRB200_VA_POL_LD=2
run JRB200_VA_BY=16 JRB200_VA_NCHUNK=6
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=6
run JRB200_VA_BY=8 JRB200_VA_NCHUNK=5
grep -E "===|value" gpurun_out/r1_bench_v1b.log | sed -E 's/.*"value": ([0-9.]+).*"T_eff_GBs_per_gpu": ([0-9.]+).*/ips=\1 Teff=\2/'
