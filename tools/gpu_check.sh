#!/bin/bash
# gpurun call: the full GPU suite + smoke after a change.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/check.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -12
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== done"
