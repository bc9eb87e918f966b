#!/bin/bash
TAG=$1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dlrm.py -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log; grep -n "^FAILED\|^ERROR" gpurun_out/${TAG}_pytest.log | head
PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 10 2>&1 | tail -1 | tee gpurun_out/${TAG}_dlrm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 600 --csv --log-file gpurun_out/${TAG}_dlrm_launches.csv env PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 3 --vocab 100000 > gpurun_out/${TAG}_ncu.log 2>&1
wc -l gpurun_out/${TAG}_dlrm_launches.csv
