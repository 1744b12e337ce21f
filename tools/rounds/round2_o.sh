#!/bin/bash
# r2-o: shade-side merges (one unit-sphere draw, one normalisation and one sphere test for the materials / primitive
# kinds of a warp that need them) against the build before them; all configs, three rounds; then the GPU suite.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
rm -f $O/o_ab.txt
for R in 1 2 3; do
  for V in base lib; do
    D=variants_build/$V; [ $V = lib ] && D=raytracinginrust_b200/lib
    echo "== $V round $R" | tee -a $O/o_ab.txt
    RTB200_LIB_DIR=$D timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/o_ab.txt
  done
done
timeout 1200 python -m pytest tests -x -q -m gpu > $O/o_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/o_pytest.log
