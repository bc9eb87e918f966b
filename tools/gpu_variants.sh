#!/bin/bash
# A/B the k_pair_step tuning variants: tools/gpu_variants.sh <tag> v...
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "pairwise_step or full_size or host_buffers" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for v in "$@"; do
  ORX_PAIR_VARIANT=$v python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "test_pairwise_step and 128" 2>&1 | tail -1
  ORX_PAIR_VARIANT=$v python bench.py --steps 1000 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench_v$v.json 2> gpurun_out/${TAG}_bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_v$v.json"))
    r=d["roofline"]
    print("variant $v: value %.1fM e2e %.1fM ms/step %.4f  kernel %.4f ms frac %.3f phases %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["frac"], {k: round(x,4) for k,x in r["phase_ms_per_step"].items()}))
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/${TAG}_bench_v$v.err").read()[-1500:])
PY
done
