#!/bin/bash
# One GPU: loopback parity tests of the sharded step, then the HBM-only cost of its launches (2 virtual ranks), A/B.
TAG=${1:-sh1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard_loopback.py -q -m gpu -x 2>&1 | tail -3
ORX_SH_SERVE=ldst timeout 600 python -m pytest tests/test_gpu_shard_loopback.py -q -m gpu -x -k "matches_oracle or dims" 2>&1 | tail -1
for v in tma ldst; do echo "serve=$v"; ORX_SH_SERVE=$v python tools/shard_loopback_probe.py 2 20 2>&1 | tail -1; done
echo "pdl=0"; ORX_PDL=0 python tools/shard_loopback_probe.py 2 20 2>&1 | tail -1
echo "R=8"; python tools/shard_loopback_probe.py 8 10 2>&1 | tail -1
echo "--- gemm probe"; python tools/gemm_probe.py 10 2>&1 | tail -7
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tc2" -s 6 -c 3 -f -o gpurun_out/${TAG}_gemm python tools/gemm_probe.py 1 > gpurun_out/${TAG}_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
ncu -i gpurun_out/${TAG}_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_raw.csv 2>/dev/null; wc -l gpurun_out/${TAG}_gemm_raw.csv
