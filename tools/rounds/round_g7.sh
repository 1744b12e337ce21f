#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 400 python tools/wf_probe2.py final:2048 cornell:1000 cornell_smoke:1000 random:800 mesh:64 > $O/g7_newchunks.txt 2>&1
cat $O/g7_newchunks.txt
timeout 900 python -m pytest tests -x -q -m gpu > $O/g7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/g7_pytest.log
timeout 600 python bench.py --workload final --steps 2 --warmup 3 > $O/g7_bench_final.json 2> $O/g7_bench_final.err; echo "bench final rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/g7_bench_final.json').read().strip().split("\n")[-1])
print("final", d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
PY
