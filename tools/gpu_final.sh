#!/bin/bash
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log; grep -n "^FAILED\|^ERROR" gpurun_out/${TAG}_pytest.log | head -20
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 1000 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json | cut -c1-400
PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 10 2>&1 | tail -1 | tee gpurun_out/${TAG}_dlrm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 420 --csv --log-file gpurun_out/${TAG}_dlrm_launches.csv env PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 3 --vocab 100000 > gpurun_out/${TAG}_ncu.log 2>&1
ls gpurun_out | grep ${TAG}
