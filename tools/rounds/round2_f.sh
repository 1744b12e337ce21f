#!/bin/bash
# r2-f (2 GPUs): bench.py under torchrun at N = 2 (compile once on rank 0 + ncclBroadcast of the table blob, sample
# blocks, ncclReduce), the reference arm launched the same way, and the multi-GPU tests on a box that has peers.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/f_gpus.txt 2>&1; nproc >> $O/f_gpus.txt
timeout 600 python -m pytest tests/test_gpu_output_multi.py -x -q -m gpu > $O/f_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -2 $O/f_pytest_multi.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/f_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/f_bench_2gpu.json 2> $O/f_bench_2gpu.err; echo "bench N=2 rc=$?"; tail -3 $O/f_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/f_bench_reference_2gpu.json 2> $O/f_bench_reference_2gpu.err; echo "reference N=2 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > $O/f_bench_1gpu.json 2> $O/f_bench_1gpu.err; echo "bench N=1 rc=$?"
ls -la $O | tail -8
