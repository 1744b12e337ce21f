#!/bin/bash
# final validation of round 1-g: GPU suite, smoke, bench lines (Cornell default, final, mesh)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/g9_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/g9_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/g9_smoke.log 2>&1; echo "smoke rc=$?"; cat $O/g9_smoke.log
timeout 600 python bench.py > $O/g9_bench_cornell.json 2> $O/g9_bench_cornell.err; echo "bench cornell rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/g9_bench_reference.json 2> $O/g9_bench_reference.err; echo "bench reference rc=$?"
timeout 600 python bench.py --workload final --steps 2 --warmup 3 > $O/g9_bench_final.json 2> $O/g9_bench_final.err; echo "bench final rc=$?"
timeout 900 python bench.py --workload mesh --steps 2 --warmup 3 --cpu-spp 4 > $O/g9_bench_mesh.json 2> $O/g9_bench_mesh.err; echo "bench mesh rc=$?"
python - <<'PY'
import json
for w in ("cornell","final","mesh"):
    try:
        d=json.loads(open('gpurun_out/g9_bench_%s.json'%w).read().strip().split("\n")[-1])
        print(w, round(d["value"],1), round(d["mrays_per_s"],1), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1), "ppm", d["e2e_ppm"] and round(d["e2e_ppm"]["value"],1), d["clocks"])
    except Exception as e: print(w, "failed", e)
PY
