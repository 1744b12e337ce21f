#!/bin/bash
# 1-GPU call: validation of the sorted wavefront stages: full GPU suite, bench lines, steady-state captures.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
RTB200_PIPELINE=wavefront timeout 200 python tools/wf_probe2.py mesh:16 > $O/g8_mesh_wavefront.txt 2>&1
RTB200_PIPELINE=wavefront RTB200_WF_LEAVE=33 timeout 200 python tools/wf_probe2.py mesh:16 >> $O/g8_mesh_wavefront.txt 2>&1
cat $O/g8_mesh_wavefront.txt
timeout 900 python -m pytest tests -x -q -m gpu --durations=5 > $O/g8_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/g8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/g8_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/g8_bench_cornell.json 2> $O/g8_bench_cornell.err; echo "bench cornell rc=$?"
timeout 600 python bench.py --workload final --steps 2 --warmup 3 > $O/g8_bench_final.json 2> $O/g8_bench_final.err; echo "bench final rc=$?"
timeout 300 python bench.py --workload random --steps 3 --warmup 3 > $O/g8_bench_random.json 2> $O/g8_bench_random.err; echo "bench random rc=$?"
timeout 300 python bench.py --workload cornell_smoke --steps 3 --warmup 3 > $O/g8_bench_cornell_smoke.json 2> $O/g8_bench_cornell_smoke.err; echo "bench smoke rc=$?"
RTB200_WF_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/g8_launches_wavefront_final.csv python tools/profile_scene.py final 64 > $O/g8_ncu_list.log 2>&1
for k in wf_extend_simple_kernel wf_shade_kernel; do
  RTB200_WF_GRAPH=0 timeout 400 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 3 --launch-count 1 \
    -o $O/g8_$k -f python tools/profile_scene.py final 64 > $O/g8_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls $O | grep g8_
