#!/bin/bash
# session 6, call A: 3D-VC parity, secondary bench lines, launch lists + full ncu of the VC / thermal / 2D kernels
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py tests/test_gpu_stokes2d.py -x -q > gpurun_out/s6a_tests.log 2>&1
tail -15 gpurun_out/s6a_tests.log
timeout 600 python scripts/bench_paths.py --steps 100 > gpurun_out/s6a_paths.jsonl 2> gpurun_out/s6a_paths.err
cat gpurun_out/s6a_paths.jsonl; tail -5 gpurun_out/s6a_paths.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s6a_launches_vc3d.csv python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6a_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s6a_launches_thermal3d.csv python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s6a_l2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s6a_launches_2d.csv python scripts/bench_paths.py --only solcx2d,shearband2d --steps 8 --warmup 2 > gpurun_out/s6a_l3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_|k_free_slip3' -s 12 -c 4 -o gpurun_out/s6a_vc3d_full -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6a_n1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_th_' -s 12 -c 4 -o gpurun_out/s6a_thermal3d_full -f python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s6a_n2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stokes2d' -s 12 -c 1 -o gpurun_out/s6a_shearband2d_full -f python scripts/bench_paths.py --only shearband2d --steps 8 --warmup 2 > gpurun_out/s6a_n3.log 2>&1
ls -la gpurun_out
