#!/bin/bash
# Rebuild the device library with different launch bounds and time the five configs (run on the GPU box).
#   tools/sweep_build.sh "128:2 128:3 128:4 256:1 256:2" "cornell random mesh"
set -e
cd "$(dirname "$0")/.."
for v in $1; do
  blk=${v%%:*}; mb=${v##*:}
  rm -f raytracinginrust_b200/lib/librtb200.so
  make -C raytracinginrust_b200/csrc -s EXTRA_NVCCFLAGS="-DRT_RENDER_BLOCK=$blk -DRT_MIN_BLOCKS=$mb" > /dev/null 2>&1
  regs=$(grep -A3 "render_kernel" raytracinginrust_b200/lib/ptxas.log | grep -o "Used [0-9]* registers" | head -1)
  spill=$(grep -A2 "Function properties for _ZN9rtb200dev13render_kernel" raytracinginrust_b200/lib/ptxas.log | grep -o "[0-9]* bytes spill stores" | head -1)
  echo "== block $blk minBlocks $mb: $regs, $spill"
  python tools/gpu_probe.py $2 2>&1 | awk '{print "   ", $1, $2, $9, $10, $11, $12, $13, $14, $15}'
done
rm -f raytracinginrust_b200/lib/librtb200.so
make -C raytracinginrust_b200/csrc -s > /dev/null 2>&1
