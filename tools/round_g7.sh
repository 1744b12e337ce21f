#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
rm -f $O/g7_ab.txt
for v in new nid nip nil niall; do
  d=variants_build/$v; [ $v = new ] && d=raytracinginrust_b200/lib
  echo "== $v" >> $O/g7_ab.txt
  RTB200_LIB_DIR=$d timeout 240 python tools/wf_probe2.py final:256 cornell:250 random:128 >> $O/g7_ab.txt 2>&1 || echo "   (failed: rc=$?)" >> $O/g7_ab.txt
done
cat $O/g7_ab.txt
