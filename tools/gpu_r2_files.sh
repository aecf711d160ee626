#!/bin/bash
# Round 2 (1 GPU): the from-files leg with the staging matrix kept across calls, both decoders.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_files.log) 2>&1
export RMG_BENCH_SPIN_S=0
for dec in builtin zlib; do
RMG_INFLATE=$([ $dec = zlib ] && echo zlib || echo "") timeout -k 10 900 python bench.py --steps 3 --warmup 3 --parts files --skip-cpu 2>gpurun_out/r2files_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); json.dump(d['from_files'], open('gpurun_out/r2f_from_files_$dec.json','w'), indent=1); f=d['from_files']
print('$dec', {k: (round(v,4) if isinstance(v,float) else v) for k,v in f.items() if k not in ('workload','timing')})" || tail -5 gpurun_out/r2files_err.txt
done
echo "=== done"
