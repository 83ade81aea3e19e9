#!/bin/bash
# session 6, call G (8 GPUs): multi-GPU parity at 2/4/8 ranks (halo, all-reduce, 3D-VA, 3D-VC, thermal), config 5 coupled bench at 2x2x2, headline bench at N=8
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s6g_tests_multi.log 2>&1; tail -8 gpurun_out/s6g_tests_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 scripts/bench_convection.py --size 257 --stokes-iters 100 --thermal-iters 100 > gpurun_out/s6g_conv_n8.json 2> gpurun_out/s6g_conv_n8.err
cat gpurun_out/s6g_conv_n8.json; tail -3 gpurun_out/s6g_conv_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/s6g_bench_n8.json 2> gpurun_out/s6g_bench_n8.err
cat gpurun_out/s6g_bench_n8.json; tail -3 gpurun_out/s6g_bench_n8.err
