#!/bin/bash
# r2-n: a medium whose boundary is one box or sphere answers both boundary queries (medium.rs:29-30) from one slab /
# root computation on one transformed ray: A/B on the two scenes with media, then the GPU suite and the bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
rm -f $O/n_ab.txt
for R in 1 2 3; do
  echo "== generic round $R" | tee -a $O/n_ab.txt
  RTB200_NO_CONVEX_MEDIA=1 timeout 300 python tools/wf_probe2.py cornell_smoke:500 final:128 cornell:250 2>&1 | tee -a $O/n_ab.txt
  echo "== fused round $R" | tee -a $O/n_ab.txt
  timeout 300 python tools/wf_probe2.py cornell_smoke:500 final:128 cornell:250 2>&1 | tee -a $O/n_ab.txt
done
timeout 1200 python -m pytest tests -x -q -m gpu > $O/n_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/n_pytest.log
timeout 900 python bench.py > $O/n_bench.json 2> $O/n_bench.err; echo "bench rc=$?"
