#!/bin/bash
TAG=$1
mkdir -p gpurun_out
ORX_FUSED=1 ORX_FUSED_CFG=24x2 ORX_FUSED_DBG=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>&1 | grep -E "fused dbg|Error|error" | tail -8
ORX_FUSED=1 ORX_FUSED_CFG=8x8 ORX_FUSED_DBG=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu 2>&1 | grep -E "fused dbg|Error|error" | tail -4
echo "--- dlrm tests (tcgen05)"; timeout 600 python -m pytest tests/test_gpu_dlrm.py -q -m gpu 2>&1 | tail -30
echo "--- dlrm tests (SIMT)"; ORX_MLP_SIMT=1 timeout 600 python -m pytest tests/test_gpu_dlrm.py -q -m gpu 2>&1 | tail -5
