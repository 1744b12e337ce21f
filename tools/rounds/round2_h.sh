#!/bin/bash
# r2-h: BVH::new on the GPU (gpu_bvh.cu): memcheck of the build on a small mesh, the bit-identity tests, and
# create / render times of both trees on the bench's mesh; then the bench line with all five configs.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gpu_built or first_hit or mesh" > $O/h_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/h_pytest.log
timeout 300 python tools/gpu_bvh_probe.py mesh 16 0 2>&1 | tee $O/h_gpu_bvh_probe.txt
timeout 900 python bench.py > $O/h_bench.json 2> $O/h_bench.err; echo "bench rc=$?"; tail -2 $O/h_bench.err
ls -la $O | tail -6
