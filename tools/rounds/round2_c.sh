#!/bin/bash
# r2-c: the deferred-traversal megakernel (render_deferred_kernel) and the FFMA slab test on the B200:
# bit-identity first, then the A/B over register budgets and thresholds, the shared-memory top-levels builds,
# all five configs on the new slab test, and one full ncu capture of the new kernel on the mesh scene.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "deferred or mesh or random" > $O/c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c_pytest.log
echo "== all configs, defaults" | tee $O/c_ab.txt
timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee -a $O/c_ab.txt
echo "== mesh, inline traversal (render_kernel), budgets 0 (80 regs) 1 (64 regs)" | tee -a $O/c_ab.txt
for B in 0 1; do RTB200_DEFER=0 RTB200_RENDER_VARIANT=$B timeout 120 python tools/wf_probe2.py mesh:16 2>&1 | sed "s/^/budget $B: /" | tee -a $O/c_ab.txt; done
echo "== mesh, deferred traversal: budget (0: 80 regs x6, 1: 64 x8, 2: 96 x5, 3: 128 x4) x threshold" | tee -a $O/c_ab.txt
for B in 0 1 2 3; do for T in 8 12 16 20 24 28 32; do
  RTB200_DEFER=1 RTB200_RENDER_VARIANT=$B RTB200_DEFER_THRESHOLD=$T timeout 120 python tools/wf_probe2.py mesh:16 2>&1 | sed "s/^/budget $B thr $T: /" | tee -a $O/c_ab.txt
done; done
echo "== mesh, top BVH levels in shared memory (budget 0 and 1, threshold 16)" | tee -a $O/c_ab.txt
for V in top128 top256; do for B in 0 1; do
  RTB200_LIB_DIR=variants_build/$V RTB200_RENDER_VARIANT=$B timeout 120 python tools/wf_probe2.py mesh:16 2>&1 | sed "s/^/$V budget $B: /" | tee -a $O/c_ab.txt
done; done
echo "== random spheres with deferral forced (every ray needs the BVH: the other extreme)" | tee -a $O/c_ab.txt
RTB200_DEFER=1 timeout 120 python tools/wf_probe2.py random:128 2>&1 | tee -a $O/c_ab.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_deferred_kernel --launch-skip 1 --launch-count 1 \
  -o $O/c_render_deferred_kernel_mesh -f python tools/profile_scene.py mesh 4 > $O/c_ncu_mesh.log 2>&1; echo "ncu mesh rc=$?"
ls -la $O | tail -8
