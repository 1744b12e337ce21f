#!/bin/bash
# Build tuning variants of the device library side by side (CPU only; nvcc cross-compiles):
#   tools/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."
# Each lands in variants_build/<name>/ (librtb200.so + host lib); select one at run time with
# RTB200_LIB_DIR=variants_build/<name>.
cd "$(dirname "$0")/.."
root=$(pwd)
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (
    out=$root/variants_build/$name
    mkdir -p $out
    cd raytracinginrust_b200/csrc
    make -s -j4 OUT=$out EXTRA_NVCCFLAGS="$flags" > $out/build.log 2>&1 || echo "build failed: $name"
    grep -E "wf_extend|wf_shade|render_kernel" -A2 $out/ptxas.log | grep -oE "Used [0-9]+ registers|[0-9]+ bytes spill stores" | paste -sd' ' | sed "s|^|$name: |"
  ) &
done
wait
