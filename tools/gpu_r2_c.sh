#!/bin/bash
# Round 2, call C (1 GPU): full GPU suite (incl. the executed .Call shim), smoke, bench legs for the skewed-warp
# permutation GEMM, ncu raw + source pages of it.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_c.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
export RMG_BENCH_SPIN_S=0
echo "=== bench N=1: randomize + multi"
timeout -k 10 1200 python bench.py --steps 5 --warmup 3 --parts randomize,multi > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
echo "rc=$?"; tail -5 gpurun_out/r2c_bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c_bench_n1.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("randomize") or {}
        m = d.get("single_process_multi") or {}
        print(f, json.dumps({"value": d["value"], "rand_ms": r.get("ms"), "rand_e2e_multi": (m.get("randomize_e2e") or {}).get("seconds"),
                             "dataset": m.get("resident_dataset"), "parity": m.get("multi_parity"),
                             "frac_of_dgemm": r.get("frac_of_dgemm"), "dgemm": r.get("dgemm_peak_tflops")}, indent=1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "=== ncu perm"
CMD="python bench.py --steps 1 --warmup 3 --perms 256 --parts randomize"
$CMD > gpurun_out/plain_perm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"perm_gemm_kernel" -s 1 -c 1 -o /tmp/r02c_perm -f $CMD > gpurun_out/ncu_perm.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02c_perm.ncu-rep --page raw --csv > gpurun_out/r02c_perm_raw.csv 2>/dev/null
ncu -i /tmp/r02c_perm.ncu-rep --page source --csv > gpurun_out/r02c_perm_source.csv 2>/dev/null
echo "=== done"
