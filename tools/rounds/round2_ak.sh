#!/bin/bash
# r2-ak: RTiOW-class scenes take the wavefront pipeline above 8e7 paths: probe, GPU suite, bench line
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/wf_probe2.py random:800 random:400 random:320 random:100 2>&1 | tee $O/ak_probe.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/ak_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/ak_pytest.log
timeout 900 python bench.py > $O/ak_bench.json 2> $O/ak_bench.err; echo "bench rc=$?"; tail -2 $O/ak_bench.err
python - <<PY
import json
d=json.load(open("$O/ak_bench.json"))
print("cornell value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) ppm %.0f frac %.4f" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_min"], d["e2e"]["ms_max"], d["ms_per_step"], d["e2e_ppm"]["value"], d["roofline"]["frac"]))
for k,v in d["workloads"].items(): print("  %-22s value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) frac %.3f" % (k, v["value"], v["e2e"]["value"], v["e2e"]["ms_min"], v["e2e"]["ms_max"], v["ms_per_step"], v["roofline"]["frac"]))
PY
