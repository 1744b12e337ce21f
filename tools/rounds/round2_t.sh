#!/bin/bash
# r2-t: the tree after the host-side changes of the afternoon (radix sort in the reference order, blob without zero fill,
# four-lane checksum): GPU suite, smoke, both bench arms as the driver runs them.  Kernels unchanged since r2-p.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/t_gpus.txt 2>&1; nproc >> $O/t_gpus.txt
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > $O/t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/t_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/t_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/t_bench_reference.json 2> $O/t_bench_reference.err; echo "reference rc=$?"; cat $O/t_bench_reference.json | cut -c1-400
timeout 900 python bench.py > $O/t_bench.json 2> $O/t_bench.err; echo "bench rc=$?"; tail -2 $O/t_bench.err; cat $O/t_bench.json | cut -c1-1500
