#!/bin/bash
# 8-GPU validation: home-routed parity test at world 4 (pytest caps it) + probe + bench at N=8.   usage: tools/gpu_n8.sh TAG
TAG=${1:-n8}; N=8
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu -x -k home > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/shard_probe.py 40 > gpurun_out/${TAG}_probe.txt 2> gpurun_out/${TAG}_probe.err; echo "probe rc=$?"; grep whole_step gpurun_out/${TAG}_probe.txt | head -8; tail -3 gpurun_out/${TAG}_probe.err | grep -v "^\*\|OMP_NUM"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench20.json 2> gpurun_out/${TAG}_bench20.err; echo "bench20 rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench20.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench20_n4.json 2> gpurun_out/${TAG}_bench20_n4.err; echo "bench20 n4 rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench20_n4.json
