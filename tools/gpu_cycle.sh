#!/bin/bash
# One GPU cycle under gpurun: parity tests, smoke, bench, ncu launch list + full capture of the top kernel.
# usage: tools/gpu_cycle.sh <tag> [pytest-args...]
TAG=${1:-r1}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -m pytest tests -q -m gpu "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
python bench.py --steps 500 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pair_step -s 3 -c 2 -f -o gpurun_out/${TAG}_pairstep \
    python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/${TAG}_ncu_bench.log 2>&1
ls -la gpurun_out | tail -12
