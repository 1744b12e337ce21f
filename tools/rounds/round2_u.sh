#!/bin/bash
# r2-u: (1) the two GPU tests added after r2-t (CUDA path against the second restatement's PBR material and sphere
# lights); (2) ncu --set full of the RTiOW render kernel (configs[0]; the last capture of it is from r1-b) and of the
# Cornell smoke kernel; (3) the bench line with the mesh workload at 128 spp.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_second_hand.py -q -s -m gpu -k "pbr or showcase" > $O/u_second_hand_pbr.log 2>&1; echo "pbr rc=$?"; grep -E "paths|passed|failed" $O/u_second_hand_pbr.log
timeout 120 python tools/profile_scene.py random 128 > $O/u_random.txt 2>&1; cat $O/u_random.txt
timeout 500 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/u_render_kernel_random -f python tools/profile_scene.py random 128 > $O/u_ncu_random.log 2>&1; echo "ncu random rc=$?"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/u_render_kernel_smoke -f python tools/profile_scene.py cornell_smoke 250 > $O/u_ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"
timeout 900 python bench.py --no-cpu-baseline > $O/u_bench.json 2> $O/u_bench.err; echo "bench rc=$?"; tail -2 $O/u_bench.err
ls -la $O | tail -6
