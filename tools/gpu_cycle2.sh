#!/bin/bash
TAG=$1
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/randrow_bw tools/randrow_bw.cu && /tmp/randrow_bw | tee gpurun_out/${TAG}_randrow_bw.txt
python -m pytest tests/test_gpu_dlrm.py tests/test_gpu_kernels.py -q -m gpu -k "dlrm or mlp or interaction or sparse_apply or owner or slots" 2>&1 | tail -25
