#!/bin/bash
TAG=$1
mkdir -p gpurun_out
for cfg in 24x2 24x3 16x4 16x3 32x2; do
  ORX_FUSED=1 ORX_FUSED_CFG=$cfg ORX_FUSED_DBG=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>&1 | grep -E "fused dbg|rror" | tail -2
  ORX_FUSED=1 ORX_FUSED_CFG=$cfg timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench_$cfg.json 2> gpurun_out/${TAG}_bench_$cfg.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$cfg.json"))
    r=d["roofline"]
    print("cfg $cfg: value %.1fM e2e %.1fM ms/step %.4f  kernel %.4f ms frac %.3f  ucml %.1fM sgd %.1fM" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["frac"], d["extra"]["ucml_adagrad_triplets_per_sec"]/1e6, d["extra"]["bpr_sgd_triplets_per_sec"]/1e6))
except Exception as e:
    print("cfg $cfg failed", e); print(open("gpurun_out/${TAG}_bench_$cfg.err").read()[-1500:])
PY
done
echo "--- all gpu tests, fused on"; ORX_FUSED=1 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -12
