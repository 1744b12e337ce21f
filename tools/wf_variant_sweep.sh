#!/bin/bash
# Time the wavefront pipeline of every prebuilt variant (tools/build_variants.sh) on the given scenes.
cd "$(dirname "$0")/.."
for d in variants_build/*/; do
  name=$(basename $d)
  echo "== $name"
  RTB200_LIB_DIR=$d RTB200_PIPELINE=${PIPE:-wavefront} timeout 300 python tools/gpu_probe.py $@ 2>&1 | awk '{print "   ", $1, $2, $9, $10, $11, $12, $13, $14, $15}'
done
