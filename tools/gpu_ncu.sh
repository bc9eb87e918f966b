#!/bin/bash
# One GPU: ncu --set full captures of the shipped kernels (one or two launches each, taken well inside the run), the
# launch lists of the default bench and of the DLRM step, and their summaries (tools/ncu_pick.py).  usage: gpu_ncu.sh TAG
TAG=${1:-ncu}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {  # name regex skip count command...
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 900 $NCU -k regex:"$re" -s $skip -c $cnt -o gpurun_out/${TAG}_$name "$@" > gpurun_out/${TAG}_ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i gpurun_out/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null
  python tools/ncu_pick.py gpurun_out/${TAG}_${name}_raw.csv > gpurun_out/${TAG}_${name}_summary.txt 2>&1; wc -l gpurun_out/${TAG}_${name}_summary.txt
  rm -f gpurun_out/${TAG}_$name.ncu-rep gpurun_out/${TAG}_${name}_raw.csv      # gpurun_out/ travels back only below 64 MiB
}
cap pair "k_pair_step|k_sparse_tail|k_index_build" 60 6 python bench.py --workload bpr --steps 30 --warmup 5 --no-cpu --no-secondary
cap censor "k_censor" 30 1 python bench.py --workload ucml --steps 20 --warmup 5 --no-cpu --no-secondary
cap point "k_point_step|k_point_tail|k_sparse_tail" 12 3 python tools/point_probe.py 10
cap gemm "k_gemm_tma" 30 6 python bench.py --workload dlrm --steps 3 --warmup 2 --no-cpu
cap dlrm_rest "k_interact|k_gather_strided|k_sparse_apply|k_act_bwd|k_col_sum|k_pred_loss|k_splitk" 200 14 python bench.py --workload dlrm --steps 3 --warmup 2 --no-cpu
cap shard "k_sh_" 24 12 python tools/shard_loopback_probe.py 2 6
# launch lists (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_bpr.csv python bench.py --workload bpr --steps 20 --warmup 5 --no-cpu --no-secondary > /dev/null 2>&1; echo "launch list bpr rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/${TAG}_launches_dlrm.csv python bench.py --workload dlrm --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1; echo "launch list dlrm rc=$?"
ls -la gpurun_out | grep ${TAG}_ | awk '{print $5, $9}'
