#!/bin/bash
cd /root/repo
prof() {
  name=$1; shift
  env "$@" timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_va_tma -s 10 -c 1 -o gpurun_out/$name -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/$name.log 2>&1
}
prof r1_v2_by10 JRB200_VA_BY=10 JRB200_VA_NCHUNK=1 JRB200_VA_SLACK=1
prof r1_v2_by10f JRB200_VA_BY=10 JRB200_VA_NCHUNK=1 JRB200_VA_SLACK=1 JRB200_VA_FAST_RCP=1
prof r1_v2_by8 JRB200_VA_BY=8 JRB200_VA_NCHUNK=6
for f in "JRB200_VA_BY=10 JRB200_VA_NCHUNK=1 JRB200_VA_SLACK=1 JRB200_VA_FAST_RCP=1" "JRB200_VA_BY=8 JRB200_VA_NCHUNK=6 JRB200_VA_FAST_RCP=1" "JRB200_VA_BY=8 JRB200_VA_NCHUNK=6 JRB200_VA_SLACK=100000"; do
 echo "=== $f"; env $f timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | sed -E 's/.*"value": ([0-9.]+).*"T_eff_GBs_per_gpu": ([0-9.]+).*/ips=\1 Teff=\2/'
done
