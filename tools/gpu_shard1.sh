#!/bin/bash
# One GPU: loopback parity tests of the sharded step (R virtual ranks), plus the fused-prologue path at world 1 under memcheck.
TAG=${1:-sh1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard_loopback.py -q -m gpu -x > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_shard_loopback.py -q -m gpu -x -k "announced or wrap" > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -3
