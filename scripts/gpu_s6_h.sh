#!/bin/bash
# session 6, call H (1 GPU): full GPU suite, headline bench, secondary benches, launch lists and full-set ncu captures for profiles/
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/s6h_tests.log 2>&1; tail -4 gpurun_out/s6h_tests.log
timeout 600 python bench.py > gpurun_out/s6h_bench_n1.json 2> gpurun_out/s6h_bench_n1.err; cat gpurun_out/s6h_bench_n1.json | cut -c1-400
timeout 600 python scripts/bench_paths.py --steps 100 > gpurun_out/s6h_paths.jsonl 2> gpurun_out/s6h_paths.err; cut -c1-330 gpurun_out/s6h_paths.jsonl
timeout 600 python scripts/bench_convection.py --size 257 --stokes-iters 100 --thermal-iters 100 > gpurun_out/s6h_conv_n1.json 2> gpurun_out/s6h_conv_n1.err; cat gpurun_out/s6h_conv_n1.json
for w in vc3d thermal3d shearband2d solcx2d; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/s6h_launches_$w.csv python scripts/bench_paths.py --only $w --steps 8 --warmup 2 > gpurun_out/s6h_l_$w.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_vc3_' -s 12 -c 3 -o gpurun_out/s6h_vc3d_full -f python scripts/bench_paths.py --only vc3d --steps 8 --warmup 2 > gpurun_out/s6h_n1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_th_flux|k_th_update' -s 8 -c 2 -o gpurun_out/s6h_thermal3d_full -f python scripts/bench_paths.py --only thermal3d --steps 8 --warmup 2 > gpurun_out/s6h_n2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_stokes2d' -s 6 -c 1 -o gpurun_out/s6h_shearband2d_full -f python scripts/bench_paths.py --only shearband2d --steps 8 --warmup 2 > gpurun_out/s6h_n3.log 2>&1
ls gpurun_out | head -50
