#!/bin/bash
# r2-m: trims level 2 as it is now (one cosine / PI where the ONB's w is the normal, no square root for vectors of
# squared length exactly 1) and the dropped cull test of an enclosing flat group, each against its absence; then a
# line-level ncu capture of the Cornell kernel as it stands, the GPU suite and the bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
rm -f $O/m_ab.txt
for R in 1 2 3; do
  echo "== micro1+keepcull round $R" | tee -a $O/m_ab.txt
  RTB200_KEEP_ENCLOSING_CULL=1 RTB200_LIB_DIR=variants_build/micro1 timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/m_ab.txt
  echo "== micro1 round $R" | tee -a $O/m_ab.txt
  RTB200_LIB_DIR=variants_build/micro1 timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/m_ab.txt
  echo "== lib round $R" | tee -a $O/m_ab.txt
  timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/m_ab.txt
done
timeout 1200 python -m pytest tests -x -q -m gpu > $O/m_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/m_pytest.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/m_render_kernel_cornell -f python tools/profile_scene.py cornell 64 > $O/m_ncu.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py > $O/m_bench.json 2> $O/m_bench.err; echo "bench rc=$?"
ls -la $O | tail -4
