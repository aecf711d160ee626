#!/bin/bash
# Round 2, final N = 1 record: the driver's commands in the driver's order -- GPU suite, smoke, reference arm, bench -- plus the
# launch list of the bench command (ncu gpu__time_duration pass) for profiles/.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_final1.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -4
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "=== reference arm"
( time timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2>gpurun_out/r2f_ref_err.txt ) 2>&1 | grep real
tail -c 600 gpurun_out/r2f_bench_reference.json
echo "=== bench (default flags)"
( time timeout -k 10 1500 python bench.py > gpurun_out/r2f_bench_n1.json 2>gpurun_out/r2f_err.txt ) 2>&1 | grep real
tail -3 gpurun_out/r2f_err.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench_n1.json').read().strip().splitlines()[-1])
keep={k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}
print(json.dumps(keep)); print('roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e'].get('stored_u16'))
for k in ('masked_variant','vertexLm','covariate','stored_u16','randomize'):
    v=d.get(k) or {}
    print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_step','ms','achieved_gbs','frac_of_hbm_peak','perms_per_s','tensor','masked','frac_of_dgemm','frac_of_dgemm_executed','e2e','e2e_stored_u16')})
t=d.get('vertexTFCE') or {}
print('tfce', t.get('perms_per_s'), (t.get('from_pinned') or {}).get('perms_per_s'), (t.get('mincTFCE') or {}).get('seconds'))
PY
echo "=== from files"
timeout -k 10 900 python bench.py --steps 3 --warmup 3 --parts files --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); json.dump(d['from_files'], open('gpurun_out/r2f_from_files.json','w'), indent=1); print(d['from_files'])"
echo "=== launch list of the bench command"
timeout -k 10 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 3 --parts masked,vertex,cov,stored --skip-cpu > /dev/null 2>&1
python tools/summarize_profiles.py launches gpurun_out/r02f_launches_bench.csv "bench.py --steps 2 --warmup 3 --parts masked,vertex,cov,stored (round 2 final)" > gpurun_out/r02f_launches_bench.md 2>/dev/null || true
head -30 gpurun_out/r02f_launches_bench.md
echo "=== done"
