#!/bin/bash
# r2-s: (1) the four GPU tests against the reference's published pictures; (2) A/B of "the next sample starts in the
# shade stage" on the Next Week final scene (wavefront pipeline): RTB200_WF_SHADE_REGEN=0 is the old flow.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/s_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_reference_images.py -q -s -m gpu > $O/s_pictures.log 2>&1; echo "pictures rc=$?"; grep -E "cornell:|smoke:|silhouette|checker:|passed|failed" $O/s_pictures.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wavefront or deterministic or full_size" > $O/s_wf_tests.log 2>&1; echo "wavefront tests rc=$?"; tail -2 $O/s_wf_tests.log
for rep in 1 2; do
  echo "--- shade regen off (old flow)"; RTB200_WF_SHADE_REGEN=0 timeout 300 python tools/wf_probe2.py final:64 final:256 final:1024
  echo "--- shade regen on"; timeout 300 python tools/wf_probe2.py final:64 final:256 final:1024
done 2>&1 | tee $O/s_ab_final.txt
echo "--- other scenes on the wavefront pipeline, off / on" | tee $O/s_ab_others.txt
RTB200_PIPELINE=wavefront RTB200_WF_SHADE_REGEN=0 timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 2>&1 | tee -a $O/s_ab_others.txt
RTB200_PIPELINE=wavefront timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 2>&1 | tee -a $O/s_ab_others.txt
