#!/bin/bash
# 1-GPU call: class-sorted shade tiles (wavefront) against the previous shade pass; images must keep their CRC.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
rm -f $O/g6_ab.txt
run() { # name, env, cases...
  v=$1; e=$2; shift; shift
  d=variants_build/$v; [ $v = new ] && d=raytracinginrust_b200/lib
  echo "== $v $e" >> $O/g6_ab.txt
  env $e RTB200_LIB_DIR=$d timeout 240 python tools/wf_probe2.py "$@" >> $O/g6_ab.txt 2>&1 || echo "   (failed or timed out: rc=$?)" >> $O/g6_ab.txt
}
run old X=1 final:64 final:256
run new X=1 final:64 final:256
run s2 X=1 final:64 final:256
run s8 X=1 final:64 final:256
run old RTB200_PIPELINE=wavefront cornell:250 cornell_smoke:250 random:128 mesh:16
run new RTB200_PIPELINE=wavefront cornell:250 cornell_smoke:250 random:128 mesh:16
cat $O/g6_ab.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wavefront or extra or variants or deterministic" 2>&1 | tail -3
