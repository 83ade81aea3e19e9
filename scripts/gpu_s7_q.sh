#!/bin/bash
# session 7, call q: compute-sanitizer memcheck over the kernels added this session (small cases)
cd /root/repo
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/s7q_memcheck_thermal.log python -m pytest "tests/test_gpu_thermal.py::test_fused_flux_update_3d" -x -q -m gpu -k "9-8-7 or 34-17-21" 2>&1 | tail -2
echo "rc=$?"; tail -3 gpurun_out/s7q_memcheck_thermal.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/s7q_memcheck_vc3.log python -m pytest tests/test_gpu_stokes3d_vc.py -x -q -m gpu -k "fixed" 2>&1 | tail -2
echo "rc=$?"; tail -3 gpurun_out/s7q_memcheck_vc3.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/s7q_memcheck_va.log python -m pytest "tests/test_gpu_stokes3d.py::test_multi_iteration_capped_batches" "tests/test_gpu_stokes3d.py::test_mixed_boundary_flags" -x -q -m gpu 2>&1 | tail -2
echo "rc=$?"; tail -3 gpurun_out/s7q_memcheck_va.log
