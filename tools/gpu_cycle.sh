#!/bin/bash
# One GPU cycle under gpurun: parity tests, smoke, bench (+variants), ncu launch list + full capture.
# usage: tools/gpu_cycle.sh <tag> [variants...]
TAG=${1:-r1}; shift
VARIANTS=${@:-0}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -30 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
for v in $VARIANTS; do
  ORX_PAIR_VARIANT=$v python bench.py --steps 500 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench_v$v.json 2> gpurun_out/${TAG}_bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_v$v.json"))
    r=d["roofline"]
    print("variant $v: value %.1fM e2e %.1fM ms/step %.4f  kernel %.4f ms frac %.3f phases %s clocks %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["ms_per_step"], r["kernel_ms"], r["frac"], r["phase_ms_per_step"], d["clocks"]))
except Exception as e:
    print("variant $v failed", e); print(open("gpurun_out/${TAG}_bench_v$v.err").read()[-2000:])
PY
done
BEST=${BEST_VARIANT:-0}
ORX_PAIR_VARIANT=$BEST python bench.py --steps 500 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
ORX_PAIR_VARIANT=$BEST timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
ORX_PAIR_VARIANT=$BEST timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pair_step|k_sparse_tail|k_index_build" -s 9 -c 3 -f -o gpurun_out/${TAG}_step \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/${TAG}_ncu_bench.log 2>&1
ls -la gpurun_out | tail -8
