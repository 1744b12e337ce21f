#!/bin/bash
# r2-al: wavefront pool size once more, now that items are one sample long (the 4 Mi default was tuned at 8 per item)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for p in default 1048576 2097152 8388608 16777216; do
  if [ $p = default ]; then unset RTB200_WF_POOL; else export RTB200_WF_POOL=$p; fi
  echo "--- pool $p"; timeout 300 python tools/wf_probe2.py random:800 final:512 | grep -v "^$"
done 2>&1 | tee $O/al_pool.txt
