#!/bin/bash
# r2-d: after the group split became the default and the deferred kernel was removed: the whole GPU suite, all five
# configs, the wavefront pipeline on the mesh scene (its extend stage replaces finished rays), and the ncu capture
# of the mesh scene's render_kernel on the new tables.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu --durations=5 > $O/d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/d_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/d_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/d_smoke.log
echo "== all configs, defaults" | tee $O/d_ab.txt
timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 mesh:16 final:64 final:256 2>&1 | tee -a $O/d_ab.txt
echo "== mesh on the wavefront pipeline: persistent extend with ray replacement (leave thresholds), sorted simple extend (33)" | tee -a $O/d_ab.txt
for L in default 4 8 16 33; do
  if [ $L = default ]; then RTB200_PIPELINE=wavefront timeout 120 python tools/wf_probe2.py mesh:16 2>&1 | sed "s/^/leave $L: /" | tee -a $O/d_ab.txt
  else RTB200_PIPELINE=wavefront RTB200_WF_LEAVE=$L timeout 120 python tools/wf_probe2.py mesh:16 2>&1 | sed "s/^/leave $L: /" | tee -a $O/d_ab.txt; fi
done
echo "== mesh, megakernel budgets (0: 80 regs, 1: 64, 2: 40)" | tee -a $O/d_ab.txt
for B in 0 1 2; do RTB200_RENDER_VARIANT=$B timeout 120 python tools/wf_probe2.py mesh:16 2>&1 | sed "s/^/budget $B: /" | tee -a $O/d_ab.txt; done
timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/d_render_kernel_mesh -f python tools/profile_scene.py mesh 4 > $O/d_ncu_mesh.log 2>&1; echo "ncu mesh rc=$?"
RTB200_PIPELINE=wavefront RTB200_WF_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -c 200 --csv \
  --log-file $O/d_launches_wavefront_mesh.csv python tools/profile_scene.py mesh 4 > $O/d_ncu_list.log 2>&1; echo "ncu list rc=$?"
ls -la $O | tail -8
