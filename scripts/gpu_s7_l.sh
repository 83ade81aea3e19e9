#!/bin/bash
# session 7, call l: occupancy variants — fused thermal at 3 CTAs/SM (78 registers), 3D-VC prep at 4 CTAs/SM (64 registers)
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_stokes3d_vc.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s7l_tests.log
for m in 2 3 2 3; do echo "TH_MINB=$m"; JRB200_TH_MINB=$m timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | tee -a gpurun_out/s7l_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'; done
for kc in 8 32; do echo "TH_MINB=3 KCHUNK=$kc"; JRB200_TH_MINB=3 JRB200_TH_KCHUNK=$kc timeout 300 python scripts/bench_paths.py --only thermal3d --steps 100 2>&1 | grep '^{' | grep -oE '"ms_per_step": [0-9.]+'; done
echo "vc3d (prep at 4 CTAs/SM)"; timeout 300 python scripts/bench_paths.py --only vc3d --steps 60 2>&1 | grep '^{' | tee -a gpurun_out/s7l_paths.jsonl | grep -oE '"ms_per_step": [0-9.]+'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_vc3' -c 30 --csv --log-file gpurun_out/s7l_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7l_l2.log 2>&1
grep -E "k_vc3_prep<0" gpurun_out/s7l_launches_vc3d.csv | head -3 | awk -F, '{print $NF}'
