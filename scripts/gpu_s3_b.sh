#!/bin/bash
# multi-GPU check on 2 GPUs: parity worker, then bench N=1 and N=2
cd /root/repo
export JRB200_VERBOSE=0
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/s3b_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/s3b_pytest.log
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu > gpurun_out/s3b_bench1.json 2> gpurun_out/s3b_bench1.err; cut -c1-400 gpurun_out/s3b_bench1.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/s3b_bench2.json 2> gpurun_out/s3b_bench2.err; echo "bench2 rc=$?"; cut -c1-600 gpurun_out/s3b_bench2.json; tail -5 gpurun_out/s3b_bench2.err
