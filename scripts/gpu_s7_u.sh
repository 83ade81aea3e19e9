#!/bin/bash
# session 7, call u: validation at HEAD — full GPU suite, all path lines, headline bench, smoke, launch lists
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s7u_tests_gpu.log
timeout 600 python scripts/bench_paths.py --steps 200 2>&1 | grep '^{' | tee gpurun_out/s7u_paths.jsonl | grep -oE '"workload": "[a-z0-9]+"|"ms_per_step": [0-9.]+'
timeout 600 python bench.py --steps 200 --warmup 10 2>&1 | tail -1 | tee gpurun_out/s7u_bench_n1.json | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_vc3|k_free_slip3' -c 40 --csv --log-file gpurun_out/s7u_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7u_l2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_stokes2d|k_res2d' -c 30 --csv --log-file gpurun_out/s7u_launches_solcx2d.csv python scripts/bench_paths.py --only solcx2d --steps 8 --warmup 2 > gpurun_out/s7u_l3.log 2>&1
