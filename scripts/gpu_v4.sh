#!/bin/bash
cd /root/repo
export JRB200_VERBOSE=1
rm -f gpurun_out/r1_bench_v4.log
run() {
  echo "=== $*" >> gpurun_out/r1_bench_v4.log
  env "$@" timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu >> gpurun_out/r1_bench_v4.log 2>&1
}
run JRB200_X=0
run JRB200_VA_SLACK=0
run JRB200_VA_SLACK=2
run JRB200_VA_SLACK=3
run JRB200_VA_L2PROMO=3
run JRB200_VA_L2PROMO=0
run JRB200_VA_POL_LD=0
grep -E "===|value" gpurun_out/r1_bench_v4.log | sed -E 's/.*"value": ([0-9.]+).*"T_eff_GBs_per_gpu": ([0-9.]+).*/ips=\1 Teff=\2/'
grep jrb200 gpurun_out/r1_bench_v4.log | head -2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_va_tma -s 10 -c 1 -o gpurun_out/r1_v4_by10 -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/r1_v4_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1_launches_v4.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r1_v4_launches.log 2>&1
