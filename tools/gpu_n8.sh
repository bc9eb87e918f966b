#!/bin/bash
TAG=$1; N=$2
mkdir -p gpurun_out
ORX_SHARDED=mailbox timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "rc=$?"; cat gpurun_out/${TAG}_bench_n${N}.json | cut -c1-3000; grep -i "error\|Traceback" -A3 gpurun_out/${TAG}_bench_n${N}.err | head -20
