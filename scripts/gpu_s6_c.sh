#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py -x -q > gpurun_out/s6c_tests_vc.log 2>&1; tail -5 gpurun_out/s6c_tests_vc.log
timeout 600 python scripts/bench_paths.py --only vc3d,thermal3d --steps 100 > gpurun_out/s6c_paths.jsonl 2> gpurun_out/s6c_paths.err
cat gpurun_out/s6c_paths.jsonl; tail -5 gpurun_out/s6c_paths.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/s6c_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6c_l1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_' -s 12 -c 3 -o gpurun_out/s6c_vc3d_full -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6c_n1.log 2>&1
python scripts/debug_conv.py 193 2>&1 | grep -E "T:|Vz" | tail -4
