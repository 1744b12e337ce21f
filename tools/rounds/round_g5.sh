#!/bin/bash
# 1-GPU call: instruction-fetch experiment on the megakernel: all warps of a block start a segment together
# (RT_MK_SYNC) at several block sizes.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
rm -f $O/g5_ab.txt
run() { # name, cases...
  v=$1; shift
  d=variants_build/$v; [ $v = base ] && d=raytracinginrust_b200/lib
  echo "== $v" >> $O/g5_ab.txt
  RTB200_LIB_DIR=$d timeout 240 python tools/wf_probe2.py "$@" >> $O/g5_ab.txt 2>&1 || echo "   (failed or timed out: rc=$?)" >> $O/g5_ab.txt
}
for v in base sync128 sync256 sync384 sync768; do run $v cornell:250 cornell_smoke:250 random:128 mesh:16; done
cat $O/g5_ab.txt
