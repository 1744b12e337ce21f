#!/bin/bash
# r2-e: the 4-wide BVH on the B200: GPU suite (with the full-size property tests), all configs, mesh budgets,
# ncu digest of the mesh kernel, and the default bench line.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu --durations=6 > $O/e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/e_pytest.log
echo "== all configs, defaults (4-wide BVH)" | tee $O/e_ab.txt
timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 mesh:16 final:64 final:256 2>&1 | tee -a $O/e_ab.txt
echo "== mesh, megakernel budgets (0: 80 regs, 1: 64, 2: 40)" | tee -a $O/e_ab.txt
for B in 0 1 2; do RTB200_RENDER_VARIANT=$B timeout 120 python tools/wf_probe2.py mesh:16 random:128 2>&1 | sed "s/^/budget $B: /" | tee -a $O/e_ab.txt; done
echo "== final on the megakernel, mesh + random on the wavefront pipeline" | tee -a $O/e_ab.txt
RTB200_PIPELINE=megakernel timeout 120 python tools/wf_probe2.py final:64 2>&1 | tee -a $O/e_ab.txt
RTB200_PIPELINE=wavefront RTB200_WF_LEAVE=33 timeout 120 python tools/wf_probe2.py mesh:16 random:128 2>&1 | tee -a $O/e_ab.txt
timeout 900 python bench.py > $O/e_bench.json 2> $O/e_bench.err; echo "bench rc=$?"; tail -2 $O/e_bench.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/e_render_kernel_mesh -f python tools/profile_scene.py mesh 4 > $O/e_ncu_mesh.log 2>&1; echo "ncu mesh rc=$?"
ls -la $O | tail -6
