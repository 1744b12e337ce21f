#!/bin/bash
# r2-w (N GPUs): the bench line under torchrun with the tree of r2-v (device blocks parked between scenes; compile once +
# broadcast), the multi-GPU tests and smoke() on every GPU of the box.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N" | tee $O/w_gpus.txt; nproc >> $O/w_gpus.txt
timeout 600 python -m pytest tests/test_gpu_output_multi.py -x -q -m gpu > $O/w_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -2 $O/w_multi_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/w_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > $O/w_bench.json 2> $O/w_bench.err; echo "bench N=$N rc=$?"; tail -3 $O/w_bench.err
python - <<PY
import json
d=json.load(open("$O/w_bench.json"))
print("N=%d cornell value %.0f e2e %.0f (%.1f..%.1f ms of %.1f)" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["e2e"]["ms_min"], d["e2e"]["ms_max"], d["ms_per_step"]))
for k,v in d["workloads"].items(): print("  %-22s value %.0f e2e %.0f (%.1f..%.1f ms of %.1f)" % (k, v["value"], v["e2e"]["value"], v["e2e"]["ms_min"], v["e2e"]["ms_max"], v["ms_per_step"]))
PY
