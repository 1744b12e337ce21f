#!/bin/bash
# r2-ae: samples per work item by scene class (flat 32, sphere / box trees 2, meshes 8): GPU suite, smoke, the bench line,
# and the full ncu capture of the bench command's render kernel for profiles/ncu_traffic.json (the chunk count changed).
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/ae_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/ae_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/ae_smoke.log
timeout 300 python tools/wf_probe2.py cornell:1000 cornell_smoke:1000 random:800 final:512 mesh:64 2>&1 | tee $O/ae_probe.txt
timeout 900 python bench.py > $O/ae_bench.json 2> $O/ae_bench.err; echo "bench rc=$?"; tail -2 $O/ae_bench.err
python - <<PY
import json
d=json.load(open("$O/ae_bench.json"))
print("cornell value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) ppm %.0f frac %.4f chunks %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_min"], d["e2e"]["ms_max"], d["ms_per_step"], d["e2e_ppm"]["value"], d["roofline"]["frac"], d["config"]["pipeline_info"].get("chunks")))
for k,v in d["workloads"].items(): print("  %-22s value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) frac %.3f" % (k, v["value"], v["e2e"]["value"], v["e2e"]["ms_min"], v["e2e"]["ms_max"], v["ms_per_step"], v["roofline"]["frac"]))
PY
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-workloads"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ae_launches_bench.csv $CMD > $O/ae_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/ae_render_kernel_bench -f $CMD > $O/ae_ncu_full.log 2>&1; echo "ncu full rc=$?"
