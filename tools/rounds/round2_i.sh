#!/bin/bash
# r2-i: the sliced tree walk (render_sliced_kernel): bit-identity, then slice length x register budget on the mesh
# and RTiOW scenes against render_kernel.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sliced" > $O/i_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/i_pytest.log
echo "== render_kernel (defaults)" | tee $O/i_ab.txt
timeout 120 python tools/wf_probe2.py mesh:16 random:128 2>&1 | tee -a $O/i_ab.txt
for V in lib slice16 slice64 slice128; do
  D=variants_build/$V; [ $V = lib ] && D=raytracinginrust_b200/lib
  for B in 0 1 2 3; do
    echo "== sliced, steps $V (lib = 32), budget $B (0: 80 regs x6, 1: 64 x8, 2: 96 x5, 3: 128 x4)" | tee -a $O/i_ab.txt
    RTB200_LIB_DIR=$D RTB200_SLICED=1 RTB200_RENDER_VARIANT=$B timeout 120 python tools/wf_probe2.py mesh:16 random:128 2>&1 | tee -a $O/i_ab.txt
  done
done
RTB200_SLICED=1 timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_sliced_kernel --launch-skip 1 --launch-count 1 \
  -o $O/i_render_sliced_kernel_mesh -f python tools/profile_scene.py mesh 4 > $O/i_ncu_mesh.log 2>&1; echo "ncu rc=$?"
ls -la $O | tail -4
