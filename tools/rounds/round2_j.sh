#!/bin/bash
# r2-j: the tree walked by the warp (trace_group_coop: idle lanes take pending subtrees of the long walks).
# Parity first, under a timeout (the loop has warp-level synchronisation: a hang must not take the box), then the
# A/B against the build without it and over the burst length, on every config that has a tree.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "first_hit or path_radiance" > $O/j_pytest_quick.log 2>&1; echo "quick parity rc=$?"; tail -3 $O/j_pytest_quick.log
if ! grep -q " passed" $O/j_pytest_quick.log || grep -q "failed" $O/j_pytest_quick.log; then echo "parity failed: stopping"; exit 1; fi
for V in nocoop lib burst4 burst16 burst32; do
  D=variants_build/$V; [ $V = lib ] && D=raytracinginrust_b200/lib
  echo "== $V (lib: cooperative walk, burst 8)" | tee -a $O/j_ab.txt
  RTB200_LIB_DIR=$D timeout 300 python tools/wf_probe2.py mesh:16 random:128 final:64 cornell:250 2>&1 | tee -a $O/j_ab.txt
done
echo "== lib, mesh at other budgets" | tee -a $O/j_ab.txt
for B in 0 2; do RTB200_RENDER_VARIANT=$B timeout 120 python tools/wf_probe2.py mesh:16 random:128 2>&1 | sed "s/^/budget $B: /" | tee -a $O/j_ab.txt; done
timeout 1200 python -m pytest tests -x -q -m gpu > $O/j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/j_pytest.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/j_render_kernel_mesh_coop -f python tools/profile_scene.py mesh 4 > $O/j_ncu_mesh.log 2>&1; echo "ncu rc=$?"
ls -la $O | tail -4
