#!/bin/bash
# r2-g: register budgets again, now with a 72-register build (7 blocks/SM: the smallest budget whose BVH node loop
# holds its ray constants in registers) - every config on budgets 0 (80), 1 (64), 2 (40), 3 (72).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
echo "== defaults" | tee $O/g_ab.txt
timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/g_ab.txt
for B in 0 1 2 3; do
  echo "== megakernel, budget $B" | tee -a $O/g_ab.txt
  RTB200_RENDER_VARIANT=$B RTB200_PIPELINE=megakernel timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 mesh:16 final:32 2>&1 | tee -a $O/g_ab.txt
done
ls -la $O | tail -3
