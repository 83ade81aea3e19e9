#!/bin/bash
# session-3 first check: gpu tests, smoke, bench line, launch list, one full ncu capture of the fused kernel
cd /root/repo
export JRB200_VERBOSE=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s3a_pytest.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/s3a_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/s3a_smoke.log 2>&1; tail -2 gpurun_out/s3a_smoke.log
timeout 400 python bench.py > gpurun_out/s3a_bench.json 2> gpurun_out/s3a_bench.err; cat gpurun_out/s3a_bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/s3a_bench_ref.json 2> gpurun_out/s3a_bench_ref.err; cat gpurun_out/s3a_bench_ref.json
nproc; lscpu | grep "Model name"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/s3a_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/s3a_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_va_tma -s 10 -c 1 -o gpurun_out/s3a_va_tma -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/s3a_ncu.log 2>&1
ls -la gpurun_out
