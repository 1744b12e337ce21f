#!/bin/bash
# 2-GPU diagnosis of rt_render_multi on the mesh scene (did not scale in round_g2): host timeline and per-device times
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
RTB200_MULTI_TIMING=1 timeout 300 python tools/multi_probe.py mesh:32 cornell:1000 > $O/g3_probe_peer.jsonl 2> $O/g3_probe_peer.err; echo "rc=$?"
cat $O/g3_probe_peer.jsonl; grep multi $O/g3_probe_peer.err | tail -24
RTB200_NO_PEER=1 RTB200_MULTI_TIMING=1 timeout 300 python tools/multi_probe.py mesh:32 > $O/g3_probe_nopeer.jsonl 2> $O/g3_probe_nopeer.err; echo "rc=$?"
cat $O/g3_probe_nopeer.jsonl; grep multi $O/g3_probe_nopeer.err | tail -8
timeout 200 python -m pytest tests/test_gpu_output_multi.py -x -q -m gpu 2>&1 | tail -2
