#!/bin/bash
# r2-k: three instruction trims that cannot change a value (branch-free accept, no mean over a single light, the
# item -> (chunk, pixel) arithmetic by reciprocal multiplication): A/B on all configs, alternating builds, 3 rounds.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
rm -f $O/k_ab.txt
for R in 1 2 3; do
  for V in nomicro lib; do
    D=variants_build/$V; [ $V = lib ] && D=raytracinginrust_b200/lib
    echo "== $V round $R" | tee -a $O/k_ab.txt
    RTB200_LIB_DIR=$D timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/k_ab.txt
  done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "render_matches or deterministic or wavefront_equals or first_hit" > $O/k_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/k_pytest.log
