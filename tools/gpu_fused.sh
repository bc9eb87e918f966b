#!/bin/bash
TAG=$1
mkdir -p gpurun_out
ORX_FUSED=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -q -m gpu -x 2>&1 | tail -15
for cfg in "0 8" "1 4" "1 6" "1 8"; do
  set -- $cfg
  ORX_FUSED=$1 ORX_FUSED_STAGES=$2 timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench_f$1_s$2.json 2> gpurun_out/${TAG}_bench_f$1_s$2.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_f$1_s$2.json"))
    r=d["roofline"]
    print("fused=$1 stages=$2: value %.1fM e2e %.1fM ms/step %.4f  kernel %.4f ms frac %.3f phases %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["frac"], {k: round(x,4) for k,x in r["phase_ms_per_step"].items()}))
except Exception as e:
    print("fused=$1 stages=$2 failed", e); print(open("gpurun_out/${TAG}_bench_f$1_s$2.err").read()[-1500:])
PY
done
ORX_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pair_fused -s 5 -c 1 -f -o gpurun_out/${TAG}_fused python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_ncu.log
