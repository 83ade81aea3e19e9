#!/bin/bash
# session 7, call k: role-split 3D-VC stress kernel (four warps per node row) — parity and A/B
cd /root/repo
mkdir -p gpurun_out
for r in 8 4; do echo "ROLES=$r tests"; JRB200_VC3_ROLES=$r timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py -x -q -m gpu 2>&1 | tail -3; done | tee gpurun_out/s7k_tests.log
for r in 0 8 4 0 8 4; do echo "ROLES=$r"; JRB200_VC3_ROLES=$r timeout 300 python scripts/bench_paths.py --only vc3d --steps 60 2>&1 | grep '^{' | tee -a gpurun_out/s7k_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
JRB200_VC3_ROLES=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_stress_rs' -s 3 -c 1 -o gpurun_out/s7k_vc3_rs -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7k_n.log 2>&1
