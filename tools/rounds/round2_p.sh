#!/bin/bash
# r2-p: the tree as it stands at the end of the round: GPU suite, smoke, both bench arms, the launch list and the full
# ncu capture of the bench command (-> profiles/ncu_traffic.json), the probe of all five configs.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/p_gpus.txt 2>&1; nproc >> $O/p_gpus.txt
timeout 1200 python -m pytest tests -x -q -m gpu --durations=6 > $O/p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/p_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/p_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/p_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/p_bench_reference.json 2> $O/p_bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py > $O/p_bench.json 2> $O/p_bench.err; echo "bench rc=$?"; tail -2 $O/p_bench.err
timeout 300 python tools/wf_probe2.py cornell:500 cornell_smoke:250 random:128 mesh:16 final:64 > $O/p_probe_all.txt 2>&1; cat $O/p_probe_all.txt
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-workloads"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/p_launches_bench.csv $CMD > $O/p_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/p_render_kernel_bench -f $CMD > $O/p_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -8
