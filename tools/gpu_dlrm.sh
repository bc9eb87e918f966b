#!/bin/bash
# One GPU: DLRM parity tests, Dense-layer GEMM accuracy + speed probes, DLRM bench line, launch list.   usage: tools/gpu_dlrm.sh TAG
TAG=${1:-dlrm}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dlrm.py -q -m gpu -x 2>&1 | tail -12
timeout 300 python tools/gemm_probe.py 10 2>&1 | tail -7
timeout 600 python bench.py --workload dlrm --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_dlrm.json 2> gpurun_out/${TAG}_bench_dlrm.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_dlrm.json; tail -3 gpurun_out/${TAG}_bench_dlrm.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_dlrm_launches.csv python bench.py --workload dlrm --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu launches rc=$?"
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/${TAG}_dlrm_launches.csv")))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ki=h.index('Kernel Name'); mi=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[start+1:]:
    if len(r)<=mi: continue
    try: v=float(r[mi].replace(',',''))
    except: continue
    n=r[ki].split('(')[0][:56]
    agg[n][0]+=1; agg[n][1]+=v
for n,(c,v) in sorted(agg.items(), key=lambda x:-x[1][1])[:16]:
    print(f"{n:58s} {c:5d} {v/1e3:10.1f} us")
PY
