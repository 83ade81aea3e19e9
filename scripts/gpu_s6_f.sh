#!/bin/bash
cd /root/repo
for v in 1 3 4; do
  echo "VAR=$v"; JRB200_VC_VAR=$v timeout 300 python scripts/bench_paths.py --only vc3d --steps 40 2>&1 | grep -oE '"ms_per_step": [0-9.]+'
done
