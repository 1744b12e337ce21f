#!/bin/bash
# r2-aj: megakernel against wavefront at full image sizes (the probe of r2-ai printed two CRCs for the final scene at 512 spp)
cd "$(dirname "$0")/../.."
O=gpurun_out
mkdir -p $O
timeout 600 python tools/pipeline_identity_probe.py final:8 final:64 final:512 random:100 cornell:64 2>&1 | tee $O/aj_identity.txt
echo "--- with 8 samples per item (RTB200_CHUNKS=64)"; RTB200_CHUNKS=64 timeout 600 python tools/pipeline_identity_probe.py final:512 2>&1 | tee -a $O/aj_identity.txt
