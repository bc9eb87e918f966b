#!/bin/bash
TAG=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  ORX_FUSED=1 ORX_FUSED_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "test_pairwise_step and 128 and (adagrad or sgd)" 2>&1 | tail -1
  ORX_FUSED=1 ORX_FUSED_CFG=$cfg timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench_$cfg.json 2> gpurun_out/${TAG}_bench_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$cfg.json"))
    r=d["roofline"]
    print("cfg $cfg: value %.1fM e2e %.1fM ms/step %.4f  kernel %.4f ms frac %.3f" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["frac"]))
except Exception as e:
    print("cfg $cfg failed", e); print(open("gpurun_out/${TAG}_bench_$cfg.err").read()[-1500:])
PY
done
