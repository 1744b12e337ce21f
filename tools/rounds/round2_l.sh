#!/bin/bash
# r2-l: instruction trims, level 2 (Vec3 / f64 with one reciprocal - checked bit for bit against the compiler's division
# first - and one cosine / PI where the ONB's w is the normal) against level 1 (r2-k) and none; then the whole GPU suite.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
rm -f $O/l_ab.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "vec3_division" > $O/l_selftest.log 2>&1; echo "division selftest rc=$?"; tail -2 $O/l_selftest.log
for R in 1 2 3; do
  for V in nomicro micro1 lib; do
    D=variants_build/$V; [ $V = lib ] && D=raytracinginrust_b200/lib
    echo "== $V round $R" | tee -a $O/l_ab.txt
    RTB200_LIB_DIR=$D timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/l_ab.txt
  done
done
timeout 1200 python -m pytest tests -x -q -m gpu > $O/l_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/l_pytest.log
