#!/bin/bash
# r2-b: first GPU call of round 2 on the reworked bench.py: GPU tests, smoke (now with rt_render_multi), both bench
# arms, the launch list and one full ncu capture of the bench command (the source of profiles/ncu_traffic.json),
# and a baseline probe of all five configs on this box for the kernel work that follows.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/b_gpus.txt 2>&1; nproc >> $O/b_gpus.txt
timeout 900 python -m pytest tests -x -q -m gpu --durations=8 > $O/b_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/b_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/b_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/b_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/b_bench_reference.json 2> $O/b_bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py > $O/b_bench.json 2> $O/b_bench.err; echo "bench rc=$?"; tail -2 $O/b_bench.err
timeout 300 python tools/wf_probe2.py cornell:250 cornell_smoke:250 random:128 mesh:16 final:64 > $O/b_probe_all.txt 2>&1; cat $O/b_probe_all.txt
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-workloads"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/b_launches_bench.csv $CMD > $O/b_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/b_render_kernel_bench -f $CMD > $O/b_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:render_kernel --launch-skip 1 --launch-count 1 \
  -o $O/b_render_kernel_mesh -f python tools/profile_scene.py mesh 4 > $O/b_ncu_mesh.log 2>&1; echo "ncu mesh rc=$?"
ls -la $O | tail -20
