#!/bin/bash
# r2-x (2 GPUs): stdout of the bench under torchrun is the JSON line and nothing else (NCCL's banner used to precede it);
# the compile phases of the mesh scene on the box's host cores; GPU suite of the final tree on one of the GPUs.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
RTB200_COMPILE_TIMING=1 timeout 300 python tools/compile_probe.py mesh 4 > $O/x_compile_probe.txt 2>&1; tail -22 $O/x_compile_probe.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 > $O/x_bench.json 2> $O/x_bench.err; echo "bench N=$N rc=$?"
echo "stdout lines: $(wc -l < $O/x_bench.json)"; head -c 60 $O/x_bench.json; echo; grep -c "NCCL version" $O/x_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/x_bench_reference.json 2> $O/x_bench_reference.err; echo "reference rc=$?"; echo "stdout lines: $(wc -l < $O/x_bench_reference.json)"
timeout 1500 python -m pytest tests -x -q -m gpu > $O/x_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/x_pytest.log
