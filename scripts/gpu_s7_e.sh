#!/bin/bash
# session 7, call e: refresh the secondary-path numbers and profiles at HEAD (thermal with precomputed face K, staged 3D-VC)
cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/bench_paths.py --steps 100 2>&1 | grep '^{' | tee gpurun_out/s7e_paths.jsonl | cut -c1-260
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_th_|k_vc3|k_free_slip3|k_maxloc' -c 80 --csv --log-file gpurun_out/s7e_launches_thermal3d.csv python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s7e_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_th_|k_vc3|k_free_slip3|k_maxloc|k_rhog|k_visc' -c 80 --csv --log-file gpurun_out/s7e_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7e_l2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_th_flux|k_th_update' -s 4 -c 2 -o gpurun_out/s7e_thermal -f python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s7e_n1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_' -s 9 -c 3 -o gpurun_out/s7e_vc3d -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s7e_n2.log 2>&1
ls -la gpurun_out | tail -8
