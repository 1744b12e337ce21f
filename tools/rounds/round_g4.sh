#!/bin/bash
# 1-GPU call: traversal inner-loop break A/B, pipeline A/B on the mesh scene, kernel times of the output stage,
# steady-state captures of the wavefront stages and of the mesh megakernel.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
for v in base ib8 ib16 ib24 ndiv; do
  d=variants_build/$v; [ $v = base ] && d=raytracinginrust_b200/lib
  echo "== $v" >> $O/g4_ab.txt
  RTB200_LIB_DIR=$d timeout 300 python tools/wf_probe2.py mesh:16 random:128 final:64 cornell:250 >> $O/g4_ab.txt 2>&1
done
cat $O/g4_ab.txt
timeout 200 python tools/pipeline_ab.py 4 mesh > $O/g4_pipeline_ab_mesh.txt 2>&1; cat $O/g4_pipeline_ab_mesh.txt
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"ppm_|format_rgb8" -c 40 --csv \
  --log-file $O/g4_launches_encode_4k.csv python tools/encode_probe.py > $O/g4_ncu_encode.log 2>&1; echo "ncu encode rc=$?"
for k in wf_extend_simple_kernel wf_shade_kernel; do
  RTB200_WF_GRAPH=0 timeout 400 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 3 --launch-count 1 \
    -o $O/g4_$k -f python tools/profile_scene.py final 64 > $O/g4_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 500 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/g4_render_kernel_mesh -f python tools/profile_scene.py mesh 4 > $O/g4_ncu_mesh.log 2>&1; echo "ncu mesh rc=$?"
ls -la $O | grep g4_
