#!/bin/bash
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -6 gpurun_out/${TAG}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py --steps 1000 --warmup 20 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json | cut -c1-2600
python bench.py --impl reference --steps 10 --warmup 2 2>/dev/null | cut -c1-300
PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 10 2>&1 | tail -1
PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 5 --simt 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pair_step" -s 6 -c 2 -f -o gpurun_out/${TAG}_pairstep python bench.py --steps 4 --warmup 3 --no-cpu >> gpurun_out/${TAG}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"k_gemm_tc" -c 3 -f -o gpurun_out/${TAG}_gemm_tc env PYTHONPATH=compat:. python tools/bench_dlrm.py --steps 1 --vocab 100000 >> gpurun_out/${TAG}_ncu.log 2>&1
ls gpurun_out | grep ${TAG}
