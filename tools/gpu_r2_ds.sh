#!/bin/bash
# Round 2 (1 GPU): resident-dataset suite (new: every mincFDR method and the pi0 grid on the resident column).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ds.log) 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_dataset.py tests/test_gpu_parity.py -q -p no:cacheprovider -k "dataset or fdr" 2>&1 | tail -5
echo "=== done"
