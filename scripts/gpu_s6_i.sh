#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py tests/test_golden_fixtures.py -x -q -m gpu 2>&1 | tail -2
echo "staged:"; timeout 300 python scripts/bench_paths.py --only vc3d --steps 40 2>&1 | grep -oE '"ms_per_step": [0-9.]+'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_stress' -s 6 -c 1 -o gpurun_out/s6i_stress_sm2 -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6i_n2.log 2>&1
