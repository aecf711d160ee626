#!/bin/bash
# Short gpurun call: GPU test suite and smoke.  (compute-sanitizer is refused by this GPU pool, so no memcheck step.)
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/verify.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -15
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== done"
