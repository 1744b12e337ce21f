#!/bin/bash
# r2-aa: ncu --set full of the mesh render kernel (the last capture, r2-b, predates the group split and the FFMA slab)
# and of the wavefront stages on the Next Week final scene, end-of-round tree.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 120 python tools/profile_scene.py mesh 4 > $O/aa_mesh.txt 2>&1; cat $O/aa_mesh.txt
timeout 500 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/aa_render_kernel_mesh -f python tools/profile_scene.py mesh 4 > $O/aa_ncu_mesh.log 2>&1; echo "ncu mesh rc=$?"
timeout 120 python tools/profile_scene.py final 32 > $O/aa_final.txt 2>&1; cat $O/aa_final.txt
for k in wf_extend_simple_kernel wf_shade_kernel wf_generate_kernel; do
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 40 --launch-count 1 \
    -o $O/aa_$k -f python tools/profile_scene.py final 32 > $O/aa_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la $O | grep aa_
