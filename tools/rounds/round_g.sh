#!/bin/bash
# One gpurun call of round 1-g: tests, smoke, bench, smem-stack A/B, ncu of the wavefront stages.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/g_gpus.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_output_multi.py -x -q -m gpu > $O/g_pytest_new.log 2>&1; echo "new tests rc=$?"
tail -3 $O/g_pytest_new.log
timeout 900 python -m pytest tests -x -q -m gpu --durations=12 > $O/g_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 $O/g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/g_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/g_bench_cornell.json 2> $O/g_bench_cornell.err; echo "bench rc=$?"
timeout 120 python tools/encode_probe.py > $O/g_encode_probe.txt 2>&1
for v in base ss16 ss24; do
  d=variants_build/$v; [ $v = base ] && d=raytracinginrust_b200/lib
  echo "== $v" >> $O/g_smem_stack_ab.txt
  RTB200_LIB_DIR=$d timeout 300 python tools/wf_probe2.py mesh:16 random:128 final:64 >> $O/g_smem_stack_ab.txt 2>&1
done
cat $O/g_smem_stack_ab.txt
# ncu: launch list of a wavefront render (host-driven round loop so every kernel is a plain launch), then one
# full capture each of the extend and shade stages in steady state
RTB200_WF_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/g_launches_wavefront_final.csv python tools/profile_scene.py final 64 > $O/g_ncu_list.log 2>&1
for k in wf_extend_simple_kernel wf_shade_kernel; do
  RTB200_WF_GRAPH=0 timeout 400 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 8 --launch-count 1 \
    -o $O/g_$k -f python tools/profile_scene.py final 64 > $O/g_ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la $O
