#!/bin/bash
# Round 2 (1 GPU): pageable host input (R's heap) through the pinned ring -- non-temporal staging copy against plain memcpy.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_pageable.log) 2>&1
export RMG_BENCH_SPIN_S=0
echo "=== pytest -m gpu -k 'pageable or staging or chunk or host'"
timeout -k 10 900 python -m pytest tests -m gpu -q -p no:cacheprovider -k "pageable or staging or chunk or host or dataset" 2>&1 | tail -3
run() { timeout -k 10 900 python bench.py --steps 2 --warmup 3 --parts e2e_pageable --skip-cpu 2>gpurun_out/r2pg_err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']['pageable_host']; print('$1', {k:(round(v,3) if isinstance(v,float) else v) for k,v in e.items() if k!='randomize'}, 'randomize s', round(e['randomize']['seconds'],3), 'gemm window ms', round(e['randomize']['gemm_window_ms'],1))" || tail -5 gpurun_out/r2pg_err.txt; }
run nt_stores
RMG_STAGING_MEMCPY=1 run memcpy
run nt_stores_again
echo "=== done"
