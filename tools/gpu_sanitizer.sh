#!/bin/bash
# compute-sanitizer on the scatter / atomic kernels: memcheck over the pairwise, sparse-apply and sharded
# (loopback) parity tests, racecheck (shared-memory hazards) over a smaller selection.   usage: tools/gpu_sanitizer.sh TAG
TAG=${1:-san}
mkdir -p gpurun_out
SEL='test_pairwise_step[128-5000-9000-4096-adagrad-bpr] or test_pairwise_step[12-37-53-96-sgd-ucml] or test_sparse_apply[1-300-500-adagrad] or test_pairwise_step_host_runs_ahead or test_epoch_wrap'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${TAG}_memcheck_kernels.txt python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "$SEL" > gpurun_out/${TAG}_memcheck_kernels_pytest.log 2>&1; echo "memcheck kernels rc=$?"; tail -2 gpurun_out/${TAG}_memcheck_kernels_pytest.log; tail -3 gpurun_out/${TAG}_memcheck_kernels.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/${TAG}_memcheck_shard.txt python -m pytest tests/test_gpu_shard_loopback.py -q -m gpu -x -k "matches_oracle[0-1-2] or matches_oracle[1-1-3] or bad_ids or heavy" > gpurun_out/${TAG}_memcheck_shard_pytest.log 2>&1; echo "memcheck shard rc=$?"; tail -2 gpurun_out/${TAG}_memcheck_shard_pytest.log; tail -3 gpurun_out/${TAG}_memcheck_shard.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/${TAG}_racecheck.txt python -m pytest tests/test_gpu_kernels.py tests/test_gpu_shard_loopback.py -q -m gpu -x -k "test_pairwise_step[128-5000-9000-4096-adagrad-bpr] or test_sparse_apply[1-300-500-adagrad] or matches_oracle[0-1-2]" > gpurun_out/${TAG}_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${TAG}_racecheck_pytest.log; tail -3 gpurun_out/${TAG}_racecheck.txt
