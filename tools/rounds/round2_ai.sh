#!/bin/bash
# r2-ai: with one sample per item, do RTiOW / mesh change their mind about the pipeline?  And tile-sorted stages at 1 per item for the mesh.
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
{
echo "--- default pipelines"; timeout 300 python tools/wf_probe2.py random:800 mesh:16 final:512
echo "--- wavefront forced"; RTB200_PIPELINE=wavefront timeout 300 python tools/wf_probe2.py random:800 mesh:16
echo "--- wavefront forced, mesh at 1 sample per item (RTB200_CHUNKS=16)"; RTB200_CHUNKS=16 RTB200_PIPELINE=wavefront timeout 300 python tools/wf_probe2.py mesh:16
echo "--- megakernel forced, final"; RTB200_PIPELINE=megakernel timeout 300 python tools/wf_probe2.py final:512
} 2>&1 | tee $O/ai_pipelines.txt
