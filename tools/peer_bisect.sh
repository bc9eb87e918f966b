#!/bin/bash
for m in 0 7; do
  echo "mask $m"; ORX_PEER_MASK=$m ORX_PEER_DBG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$m tools/peer_probe.py 2>&1 | grep -E "peer dbg|barrier|apply|free" | tail -5
done
