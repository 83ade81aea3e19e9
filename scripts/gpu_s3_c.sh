#!/bin/bash
# 8-GPU check: parity worker on 2x2x2 (edges + corners), bench N=4 and N=8
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "8 or 4" > gpurun_out/s3c_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/s3c_pytest.log
for N in 4 8; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/s3c_bench$N.json 2> gpurun_out/s3c_bench$N.err; echo "bench$N rc=$?"; cut -c1-300 gpurun_out/s3c_bench$N.json; tail -3 gpurun_out/s3c_bench$N.err
done
