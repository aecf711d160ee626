#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/smoke_all.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== pytest -m gpu"
timeout -k 10 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 2>&1 | tail -25
echo "=== done"
