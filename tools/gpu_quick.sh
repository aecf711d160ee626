#!/bin/bash
# One gpurun call: a named subset of the GPU tests (pattern in $1).
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/quick.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -k "$1" --durations=5 2>&1 | tail -40
echo "=== done"
