#!/bin/bash
# Round 2, call B (gpurun --gpus 2): full GPU suite, smoke, N = 2 torchrun bench (randomize + multi legs), N = 1 bench legs,
# masked-fit L2 fetch granularity experiment, then the ncu raw + source pages of the reworked permutation GEMM.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_b.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu (2 GPUs visible)"
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
export RMG_BENCH_SPIN_S=0
echo "=== bench N=2 (torchrun): randomize + e2e + multi"
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 5 --warmup 3 --parts randomize,e2e,multi > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
echo "rc=$?"; tail -5 gpurun_out/r2b_bench_n2.err
echo "=== bench N=1: randomize + e2e + multi"
RMG_TIMING=1 timeout -k 10 1200 python bench.py --steps 5 --warmup 3 --parts randomize,e2e,multi,masked > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
echo "rc=$?"; tail -24 gpurun_out/r2b_bench_n1.err
for g in 32 128; do
  echo "=== masked fit, L2 fetch granularity $g"
  RMG_L2_FETCH=$g timeout -k 10 600 python bench.py --steps 5 --warmup 3 --parts masked 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['masked_variant'], d['ms_per_step'])"
done
python - <<'PY'
import json
for f in ("gpurun_out/r2b_bench_n2.json", "gpurun_out/r2b_bench_n1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("randomize") or {}
        m = d.get("single_process_multi") or {}
        print(f, json.dumps({"value": d["value"], "rand_ms": r.get("ms"), "rand_e2e": (r.get("e2e") or {}).get("seconds"),
                             "rand_e2e_multi": (m.get("randomize_e2e") or {}).get("seconds"), "fit_e2e_multi": (m.get("fit_e2e") or {}).get("value"),
                             "dataset": m.get("resident_dataset"), "parity": m.get("multi_parity"), "masked": d.get("masked_variant"),
                             "frac_of_dgemm": r.get("frac_of_dgemm")}, indent=1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "=== ncu perm"
CMD="python bench.py --steps 1 --warmup 3 --perms 128 --parts randomize"
$CMD > gpurun_out/plain_perm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"perm_gemm_kernel" -s 1 -c 1 -o /tmp/r02b_perm -f $CMD > gpurun_out/ncu_perm.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02b_perm.ncu-rep --page raw --csv > gpurun_out/r02b_perm_raw.csv 2>/dev/null
ncu -i /tmp/r02b_perm.ncu-rep --page source --csv > gpurun_out/r02b_perm_source.csv 2>/dev/null
echo "=== done"
