#!/bin/bash
# r2-ah: one sample per item for scenes with a tree over spheres / boxes: probe, GPU suite, bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/wf_probe2.py random:800 random:100 final:512 final:250 final:2000 2>&1 | tee $O/ah_probe.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/ah_pytest.log
timeout 900 python bench.py > $O/ah_bench.json 2> $O/ah_bench.err; echo "bench rc=$?"; tail -2 $O/ah_bench.err
python - <<PY
import json
d=json.load(open("$O/ah_bench.json"))
print("cornell value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) ppm %.0f frac %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_min"], d["e2e"]["ms_max"], d["ms_per_step"], d["e2e_ppm"]["value"], d["roofline"]["frac"]))
for k,v in d["workloads"].items(): print("  %-22s value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) frac %.3f" % (k, v["value"], v["e2e"]["value"], v["e2e"]["ms_min"], v["e2e"]["ms_max"], v["ms_per_step"], v["roofline"]["frac"]))
PY
