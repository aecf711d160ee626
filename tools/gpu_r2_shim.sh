#!/bin/bash
# Round 2 (1 GPU): the .Call shim suite (new: rmincgpu_per_voxel_anova on files against the reference's per_voxel_anova).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_shim.log) 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_rshim.py tests/test_rshim.py -q -p no:cacheprovider 2>&1 | tail -5
echo "=== done"
