#!/bin/bash
# r2-r: the GPU suite once more with the tests added after r2-p (the CUDA path against tests/second_hand.py on seven
# scenes), smoke, and the probe of all five configs (library unchanged since r2-p: same numbers expected).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/r_gpus.txt 2>&1; nproc >> $O/r_gpus.txt
timeout 1500 python -m pytest tests -x -q -m gpu --durations=6 > $O/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r_pytest.log
timeout 600 python -m pytest tests/test_gpu_second_hand.py -q -s -m gpu > $O/r_second_hand.log 2>&1; echo "second-hand rc=$?"; grep "paths" $O/r_second_hand.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r_smoke.log
timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 > $O/r_probe_all.txt 2>&1; cat $O/r_probe_all.txt
