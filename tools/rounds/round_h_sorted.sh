#!/bin/bash
# First GPU call of the two experiments written in r1-h without a GPU: the sorted megakernel (csrc/device/sorted.inl,
# RTB200_PIPELINE=sorted / sorted256) and one wavefront shade kernel per hit class (RTB200_WF_SHADE=perclass).
# 1. bit-identity with render_kernel on the five configs (opt-in test, under a timeout: the kernel has barriers in a
#    data-dependent loop and has never run on a GPU);  2. A/B device times, same images expected (crc column).
#   gpurun --timeout 300 -- 'bash tools/round_h_sorted.sh'
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
RTB200_TEST_EXPERIMENTS=1 timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -k 'sorted_megakernel or per_class_shade' > $O/h_sorted_pytest.log 2>&1
echo "sorted parity rc=$?"; tail -3 $O/h_sorted_pytest.log
for P in megakernel sorted sorted256; do
  echo "== RTB200_PIPELINE=$P"
  RTB200_PIPELINE=$P timeout 120 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 mesh:16 final:64 2>&1 | tee $O/h_sorted_ab_$P.txt
done
echo "== wavefront shade: one sorted pass vs one kernel per class (RTB200_WF_SHADE=perclass)"
for S in sorted perclass; do
  RTB200_PIPELINE=wavefront RTB200_WF_SHADE=$S timeout 120 python tools/wf_probe2.py final:256 cornell_smoke:250 cornell:250 2>&1 | tee $O/h_shade_ab_$S.txt
done
# ncu evidence only when the identity test passed: one --set full capture of each render kernel on the Cornell box
# (second launch: the first is the warm-up render of profile_scene.py), digests next to the reports
if grep -q " passed" $O/h_sorted_pytest.log && ! grep -q "failed" $O/h_sorted_pytest.log; then
  for P in megakernel sorted sorted256; do
    K=render_kernel; [ "$P" != megakernel ] && K=render_sorted_kernel
    RTB200_PIPELINE=$P timeout 300 ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 1 --launch-count 1 \
      -o $O/h_$P -f python tools/profile_scene.py cornell 64 > $O/h_ncu_$P.log 2>&1; echo "ncu $P rc=$?"
    python tools/ncu_summary.py $O/h_$P.ncu-rep > $O/h_${P}_summary.txt 2>&1 || true
  done
fi
