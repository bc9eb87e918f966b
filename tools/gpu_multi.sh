#!/bin/bash
# multi-GPU cycle: tools/gpu_multi.sh <tag> <ngpus>
TAG=$1; N=$2
mkdir -p gpurun_out
python tools/tc_probe.py 2>&1 | tail -12
python -m pytest tests -q -m gpu 2>&1 | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_n$N.json"))
    print("N=$N value %.1fM e2e %.1fM ms/step %.4f roofline %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], {k:d["roofline"][k] for k in ("bound","achieved","frac")}))
except Exception as e:
    print("N=$N failed", e); print(open("gpurun_out/${TAG}_bench_n$N.err").read()[-2500:])
PY
