#!/bin/bash
# session 7, last call: 2D tests at HEAD (etav prefetch), memcheck of the packed-column 3D-VC kernels and the 2D kernel
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stokes2d.py -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/s7zz_tests2d.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/s7zz_memcheck_vc3.log python -m pytest tests/test_gpu_stokes3d_vc.py -x -q -m gpu -k "fixed or exit_kernels" 2>&1 | tail -2
echo "rc=$?"; tail -2 gpurun_out/s7zz_memcheck_vc3.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/s7zz_memcheck_2d.log python -m pytest tests/test_gpu_stokes2d.py -x -q -m gpu -k "fixed" 2>&1 | tail -2
echo "rc=$?"; tail -2 gpurun_out/s7zz_memcheck_2d.log
