#!/bin/bash
# tools/scaling_sweep.sh N "workloads" — bench.py at N GPUs for each workload, JSON lines to gpurun_out/scale_N.jsonl
N=$1; WL=${2:-"cornell final mesh"}
cd "$(dirname "$0")/.."; mkdir -p gpurun_out; : > gpurun_out/scale_$N.jsonl
for w in $WL; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --no-cpu-baseline --workload $w >> gpurun_out/scale_$N.jsonl 2> gpurun_out/scale_${N}_$w.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline --workload $w >> gpurun_out/scale_$N.jsonl 2> gpurun_out/scale_${N}_$w.err
  fi
  tail -1 gpurun_out/scale_$N.jsonl | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['scene'], 'N=%d'%d['n_gpus'], '%.1f Mpaths/s'%d['value'], '%.1f Mrays/s'%d['mrays_per_s'], 'ms/step %.1f'%d['ms_per_step'], 'e2e %.1f'%d['e2e']['value'], d['clocks'])" 2>&1 | tail -1
done
