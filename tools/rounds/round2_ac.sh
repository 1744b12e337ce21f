#!/bin/bash
# r2-ac: samples per work item of the megakernel (RTB200_CHUNKS forces the number of sample chunks; default: ~8 samples per item)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  for c in default 250 63 42 32 16; do
    if [ $c = default ]; then unset RTB200_CHUNKS; else export RTB200_CHUNKS=$c; fi
    echo "--- chunks $c"; timeout 300 python tools/wf_probe2.py cornell:1000 cornell_smoke:1000 random:800 | grep -v "^$"
  done
done 2>&1 | tee $O/ac_chunks.txt
