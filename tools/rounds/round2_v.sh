#!/bin/bash
# r2-v: device blocks parked between scenes (api.cu ScratchCache).  The test of it, the GPU suite, then the bench line
# with the cache off and on (the end-to-end legs create and destroy a scene per iteration), the whole-job probe both ways.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_output_multi.py -x -q -s -m gpu -k "recycled" > $O/v_recycle.log 2>&1; echo "recycle rc=$?"; grep -E "parked|passed|failed|Error" $O/v_recycle.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/v_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/v_smoke.log
for mode in off on; do
  if [ $mode = off ]; then export RTB200_SCRATCH_CACHE_MB=0; else unset RTB200_SCRATCH_CACHE_MB; fi
  echo "--- cache $mode"
  timeout 200 python tools/ppm_phase_probe.py cornell 600 600 1000 100 5 2> $O/v_ppm_phases_$mode.txt | tail -1
  timeout 900 python bench.py --no-cpu-baseline > $O/v_bench_$mode.json 2> $O/v_bench_$mode.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open("$O/v_bench_$mode.json"))
print("cornell value %.0f e2e %.0f (%.1f..%.1f ms) ppm %.0f" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_min"], d["e2e"]["ms_max"], d["e2e_ppm"]["value"]))
for k,v in d["workloads"].items(): print("  %-22s value %.0f e2e %.0f (%.1f..%.1f ms of %.1f)" % (k, v["value"], v["e2e"]["value"], v["e2e"]["ms_min"], v["e2e"]["ms_max"], v["ms_per_step"]))
PY
done 2>&1 | tee $O/v_ab.txt
