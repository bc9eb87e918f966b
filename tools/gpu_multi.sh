#!/bin/bash
# Multi-GPU validation (gpurun --gpus N): sharded parity tests, per-launch probe, bench.   usage: tools/gpu_multi.sh TAG N
TAG=${1:-multi}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_shard_loopback.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
for PDL in 1; do
ORX_PDL=$PDL timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$PDL tools/shard_probe.py 40 > gpurun_out/${TAG}_probe_pdl$PDL.txt 2> gpurun_out/${TAG}_probe.err; echo "probe PDL=$PDL rc=$?"; head -2 gpurun_out/${TAG}_probe_pdl$PDL.txt; tail -3 gpurun_out/${TAG}_probe.err | grep -v "^\*\|OMP_NUM"
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err | grep -v "^\*\|OMP_NUM"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench20.json 2> gpurun_out/${TAG}_bench20.err; echo "bench20 rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench20.json
