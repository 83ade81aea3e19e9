#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_stokes3d_vc.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/bench_paths.py --only thermal3d,vc3d --steps 60 2>&1 | grep -oE '"workload": "[a-z0-9]+"|"ms_per_step": [0-9.]+'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/s6j_launches_thermal3d.csv python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s6j_l2.log 2>&1
