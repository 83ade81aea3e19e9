#!/bin/bash
# session 6, call B (2 GPUs): remaining 3D-VC tests, multi-GPU parity incl. 3D-VC + thermal, coupled convection bench at N=2
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stokes3d_vc.py -x -q > gpurun_out/s6b_tests_vc.log 2>&1; tail -5 gpurun_out/s6b_tests_vc.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s6b_tests_multi.log 2>&1; tail -30 gpurun_out/s6b_tests_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/bench_convection.py --n 193 --stokes-iters 50 --thermal-iters 50 > gpurun_out/s6b_conv_n2.json 2> gpurun_out/s6b_conv_n2.err
cat gpurun_out/s6b_conv_n2.json; tail -5 gpurun_out/s6b_conv_n2.err
timeout 600 python scripts/bench_convection.py --n 193 --stokes-iters 50 --thermal-iters 50 > gpurun_out/s6b_conv_n1.json 2> gpurun_out/s6b_conv_n1.err
cat gpurun_out/s6b_conv_n1.json; tail -5 gpurun_out/s6b_conv_n1.err
