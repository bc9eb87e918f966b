#!/bin/bash
# 8-GPU run: per-launch probe + the driver's bench command at N=8.   usage: tools/gpu_n8.sh TAG
TAG=${1:-n8}; N=8
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/shard_probe.py 40 > gpurun_out/${TAG}_probe.txt 2> gpurun_out/${TAG}_probe.err; echo "probe rc=$?"; head -3 gpurun_out/${TAG}_probe.txt; tail -3 gpurun_out/${TAG}_probe.err | grep -v "^\*\|OMP_NUM"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench20.json 2> gpurun_out/${TAG}_bench20.err; echo "bench20 rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench20.json; tail -3 gpurun_out/${TAG}_bench20.err | grep -v "^\*\|OMP_NUM"
