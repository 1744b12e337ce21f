#!/bin/bash
# 2-GPU call of round 1-g: the single-process multi-GPU entry over real NVLink peers, next to the torchrun + NCCL path.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/g2_gpus.txt 2>&1
nvidia-smi topo -m > $O/g2_topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_output_multi.py tests/test_multi_gpu_cpu.py -x -q -m gpu > $O/g2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/g2_pytest.log
timeout 400 python tools/multi_probe.py > $O/g2_multi_probe.jsonl 2> $O/g2_multi_probe.err; echo "probe rc=$?"; cat $O/g2_multi_probe.jsonl; tail -3 $O/g2_multi_probe.err
for w in cornell mesh; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --workload $w --no-cpu-baseline > $O/g2_bench_$w.json 2> $O/g2_bench_$w.err; echo "bench $w rc=$?"
done
timeout 200 raytracinginrust_b200/lib/rtb200_render --scene mesh --spp 8 --gpus 0 --assets assets > $O/g2_mesh.ppm 2> $O/g2_cli.err; echo "cli rc=$?"; tail -1 $O/g2_cli.err; head -c 20 $O/g2_mesh.ppm | head -2; md5sum $O/g2_mesh.ppm > $O/g2_mesh_ppm.md5; rm -f $O/g2_mesh.ppm
