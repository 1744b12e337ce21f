#!/bin/bash
# r2-q (N GPUs): the bench line under torchrun at the box's full GPU count, both arms, as the driver launches them.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
N=$(nvidia-smi -L | wc -l)
echo "gpus: $N" | tee $O/q_gpus.txt; nproc >> $O/q_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/q_bench_reference.json 2> $O/q_bench_reference.err; echo "reference rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 > $O/q_bench.json 2> $O/q_bench.err; echo "bench N=$N rc=$?"; tail -3 $O/q_bench.err
