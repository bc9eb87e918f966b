#!/bin/bash
# First GPU call of the next round: validate the default build, then every opt-in experimental path written blind at
# the end of round 1 (tools/README.md), each with its parity tests and a short bench.   usage: tools/gpu_round2.sh TAG
TAG=${1:-r2a}
mkdir -p gpurun_out
log() { echo "=== $*"; }
log "default: full gpu suite"
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest_default.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_default.log
python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench_default.json 2>/dev/null; cut -c1-260 gpurun_out/${TAG}_bench_default.json
run() {   # name, env assignments...
  name=$1; shift
  log "$name: $*"
  env "$@" timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py -q -m gpu -x > gpurun_out/${TAG}_pytest_${name}.log 2>&1
  tail -2 gpurun_out/${TAG}_pytest_${name}.log; grep -n "^FAILED\|^ERROR" gpurun_out/${TAG}_pytest_${name}.log | head -5
  env "$@" timeout 300 python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_${name}.json").read())
    print("   value %.1f M/s  e2e %.1f M/s  ms/step %.4f  phases %s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["roofline"].get("phase_ms_per_step")))
except Exception as e:
    print("   bench failed:", e)
PY
}
run la7 ORX_PAIR_VARIANT=7
run la8 ORX_PAIR_VARIANT=8
run copystream ORX_HOST_COPY_STREAM=1
run la7_overlap ORX_PAIR_VARIANT=7 ORX_HOST_COPY_STREAM=1 ORX_OVERLAP_INDEX=1
log "DLRM: default, v3, v3+pad"
python -m pytest tests/test_gpu_dlrm.py -q -m gpu 2>&1 | tail -1
PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 10 2>&1 | tail -1
ORX_MLP_TC_V=3 python -m pytest tests/test_gpu_dlrm.py -q -m gpu 2>&1 | tail -3
ORX_MLP_TC_V=3 PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 10 2>&1 | tail -1
ORX_MLP_TC_V=3 ORX_DLRM_PAD=1 python -m pytest tests/test_gpu_dlrm.py -q -m gpu 2>&1 | tail -3
ORX_MLP_TC_V=3 ORX_DLRM_PAD=1 PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 10 2>&1 | tail -1
ORX_DLRM_PAD=1 PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 10 2>&1 | tail -1
