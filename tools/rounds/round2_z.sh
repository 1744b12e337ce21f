#!/bin/bash
# r2-z: a 48-register build (10 resident blocks per SM) between the 40- and the 64-register ones.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  echo "--- default budgets"; timeout 300 python tools/wf_probe2.py random:128 cornell_smoke:250 mesh:16
  echo "--- 48 registers x 10 blocks (RTB200_RENDER_VARIANT=3; media scenes 7)"
  RTB200_RENDER_VARIANT=3 timeout 300 python tools/wf_probe2.py random:128 mesh:16
  RTB200_RENDER_VARIANT=7 timeout 300 python tools/wf_probe2.py cornell_smoke:250
done 2>&1 | tee $O/z_ab.txt
