#!/bin/bash
# r2-ad: samples per work item, the BVH scenes: mesh (megakernel) and final (wavefront); RTiOW once more at finer steps
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  for c in default 4 16 32 64; do
    if [ $c = default ]; then unset RTB200_CHUNKS; else export RTB200_CHUNKS=$c; fi
    echo "--- mesh x64, chunks $c (default 8: 8 samples per item)"; timeout 300 python tools/wf_probe2.py mesh:64 | grep -v "^$"
  done
  for c in default 16 64 128 256; do
    if [ $c = default ]; then unset RTB200_CHUNKS; else export RTB200_CHUNKS=$c; fi
    echo "--- final x512, chunks $c (default 64)"; timeout 300 python tools/wf_probe2.py final:512 | grep -v "^$"
  done
  for c in 150 200 400 800; do
    export RTB200_CHUNKS=$c
    echo "--- random x800, chunks $c (default 100)"; timeout 300 python tools/wf_probe2.py random:800 | grep -v "^$"
  done
done 2>&1 | tee $O/ad_chunks.txt
