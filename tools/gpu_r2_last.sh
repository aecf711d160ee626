#!/bin/bash
# Round 2, last call: the default bench command on the final code.
set -u
mkdir -p gpurun_out
( time timeout -k 10 900 python bench.py > gpurun_out/r2i_bench_n1.json 2>gpurun_out/r2i_err.txt ) 2>&1 | grep real
tail -2 gpurun_out/r2i_err.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'roofline', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']/1e6,2))
r=d['randomize']; print('randomize', round(r['ms'],1), r['e2e']['seconds_each_call'], r['e2e_stored_u16']['seconds_each_call'])
print('stored', round(d['stored_u16']['ms_per_step'],3), round(d['stored_u16']['masked']['ms_per_step'],3), 'masked', round(d['masked_variant']['ms_per_step'],3), 'vertex', round(d['vertexLm']['ms_per_step'],4), 'tfce', round(d['vertexTFCE']['perms_per_s']))
PY
