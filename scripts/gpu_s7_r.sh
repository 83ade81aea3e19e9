#!/bin/bash
# session 7, call r: 3D multiphase thermal solve parity (fused pairs inside the solve loop); memcheck of the fused thermal kernel
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_thermal.py -x -q -m gpu -k "diffusion3d_multiphase or config1" 2>&1 | tail -8 | tee gpurun_out/s7r_tests.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/s7r_memcheck_thermal.log python -m pytest tests/test_gpu_thermal.py -x -q -m gpu -k "test_fused_flux_update_3d and (ni0 or ni1) and not 32-" 2>&1 | tail -2
echo "rc=$?"; tail -3 gpurun_out/s7r_memcheck_thermal.log
