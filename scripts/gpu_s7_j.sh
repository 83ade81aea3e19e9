#!/bin/bash
# session 7, call j: branch-free reciprocals in the staged 3D-VC stress kernel; full GPU suite; all paths; headline bench
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s7j_tests_gpu.log
timeout 600 python scripts/bench_paths.py --steps 100 2>&1 | grep '^{' | tee gpurun_out/s7j_paths.jsonl | grep -oE '"workload": "[a-z0-9]+"|"ms_per_step": [0-9.]+'
timeout 600 python bench.py --steps 200 --warmup 10 2>&1 | tail -1 | tee gpurun_out/s7j_bench_n1.json | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
