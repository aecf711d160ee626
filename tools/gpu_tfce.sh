#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/tfce.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout -k 10 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "tfce or TFCE" 2>&1 | tail -8
export RMG_BENCH_SPIN_S=0
timeout -k 10 900 python bench.py --small --steps 3 --warmup 3 --parts tfce | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['vertexTFCE'], indent=1))"
timeout -k 10 900 python bench.py --small --steps 3 --warmup 3 --parts tfce --tfce-perms 256 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps(d['vertexTFCE'], indent=1))"
echo "=== launch list"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_tfce.csv python bench.py --small --steps 3 --warmup 3 --parts tfce --tfce-perms 16 > /dev/null 2>&1
echo "rc=$?"
echo "=== done"
