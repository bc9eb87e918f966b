#!/bin/bash
# One GPU: DLRM parity tests, Dense-layer GEMM accuracy + speed probes, DLRM bench line.   usage: tools/gpu_dlrm.sh TAG
TAG=${1:-dlrm}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dlrm.py -q -m gpu -x 2>&1 | tail -12
timeout 300 python tools/tc_probe.py 2>&1 | tail -14
timeout 300 python tools/gemm_probe.py 10 2>&1 | tail -7
timeout 600 python bench.py --workload dlrm --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_dlrm.json 2> gpurun_out/${TAG}_bench_dlrm.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_dlrm.json; tail -3 gpurun_out/${TAG}_bench_dlrm.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tma" -s 6 -c 3 -f -o gpurun_out/${TAG}_gemm python tools/gemm_probe.py 1 > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ncu -i gpurun_out/${TAG}_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_raw.csv 2>/dev/null; python tools/ncu_pick.py gpurun_out/${TAG}_gemm_raw.csv
