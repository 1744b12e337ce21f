#!/bin/bash
# r2-ab: ncu --set full of the wavefront stages on the Next Week final scene (host-driven round loop: kernel nodes of a
# graph with conditional nodes cannot be profiled), end-of-round tree.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
export RTB200_WF_GRAPH=0
timeout 120 python tools/profile_scene.py final 32 > $O/ab_final.txt 2>&1; cat $O/ab_final.txt
for k in wf_extend_simple_kernel wf_shade_kernel wf_generate_kernel; do
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 40 --launch-count 1 \
    -o $O/ab_$k -f python tools/profile_scene.py final 32 > $O/ab_ncu_$k.log 2>&1; echo "ncu $k rc=$?"; tail -2 $O/ab_ncu_$k.log
done
ls -la $O | grep "ab_.*rep"
