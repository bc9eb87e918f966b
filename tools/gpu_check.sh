#!/bin/bash
# One-GPU validation: full gpu suite, smoke, bench (default line with secondary workloads).   usage: tools/gpu_check.sh TAG [pytest args]
TAG=${1:-chk}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/${TAG}_pytest.log
grep -n "^FAILED\|^ERROR" gpurun_out/${TAG}_pytest.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
