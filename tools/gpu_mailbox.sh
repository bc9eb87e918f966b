#!/bin/bash
# usage: tools/gpu_mailbox.sh TAG NGPUS
TAG=$1; N=$2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded.py -q -m gpu -k "mailbox" -x > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 tools/mailbox_probe.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -8 | tee gpurun_out/${TAG}_probe.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  echo "== $name rc=$?"; cut -c1-230 gpurun_out/${TAG}_bench_${name}.json; grep -o '"e2e": {"value": [0-9.]*' gpurun_out/${TAG}_bench_${name}.json; grep -i "error\|Traceback" gpurun_out/${TAG}_bench_${name}.err | head -3
}
run mbox ORX_SHARDED=mailbox
