#!/bin/bash
# r2-ag: the 2^23-item floor at the sample counts one of eight GPUs sees (Cornell 125 spp, RTiOW 100 spp, final 250 spp)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  for c in default 16 12 8 4; do
    if [ $c = default ]; then unset RTB200_CHUNKS; else export RTB200_CHUNKS=$c; fi
    echo "--- cornell x125 / smoke x125, chunks $c (default: the floor, 24)"; timeout 300 python tools/wf_probe2.py cornell:125 cornell_smoke:125 | grep -v "^$"
  done
  for c in default 100 25; do
    if [ $c = default ]; then unset RTB200_CHUNKS; else export RTB200_CHUNKS=$c; fi
    echo "--- random x100, chunks $c (default 50)"; timeout 300 python tools/wf_probe2.py random:100 | grep -v "^$"
  done
done 2>&1 | tee $O/ag_floor.txt
