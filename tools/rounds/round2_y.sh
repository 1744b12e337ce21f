#!/bin/bash
# r2-y: the tree at the end of the round, as the driver runs it: GPU suite, smoke, reference arm, bench (with the CPU
# baselines), and the compile phases of the mesh scene with the finer build tasks.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/y_gpus.txt 2>&1; nproc >> $O/y_gpus.txt
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > $O/y_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/y_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/y_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/y_smoke.log
RTB200_COMPILE_TIMING=1 timeout 300 python tools/compile_probe.py mesh 4 > $O/y_compile_probe.txt 2>&1; tail -12 $O/y_compile_probe.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/y_bench_reference.json 2> $O/y_bench_reference.err; echo "reference rc=$? lines $(wc -l < $O/y_bench_reference.json)"
SECONDS=0
timeout 900 python bench.py > $O/y_bench.json 2> $O/y_bench.err; echo "bench rc=$? lines $(wc -l < $O/y_bench.json) in $SECONDS s"; tail -2 $O/y_bench.err
python - <<PY
import json
d=json.load(open("$O/y_bench.json"))
print("cornell value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) ppm %.0f frac %.4f cpu %.1f on %d" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_min"], d["e2e"]["ms_max"], d["ms_per_step"], d["e2e_ppm"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]))
for k,v in d["workloads"].items(): print("  %-22s value %.0f e2e %.0f (%.1f..%.1f ms of %.1f) frac %.3f" % (k, v["value"], v["e2e"]["value"], v["e2e"]["ms_min"], v["e2e"]["ms_max"], v["ms_per_step"], v["roofline"]["frac"]))
PY
