#!/bin/bash
# One gpurun call: parity tests, benches, then (only after the plain runs exited 0) ncu captures.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/session.log) 2>&1
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,power.limit --format=csv
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu"
timeout -k 10 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== TMA tile-shape sweep (headline only)"
for c in 0 1 2 3 4; do
  echo "--- RMG_TMA_CFG=$c"
  RMG_FIT_VARIANT=tma RMG_TMA_CFG=$c RMG_BENCH_SPIN_S=0 timeout -k 10 600 python bench.py --steps 5 --warmup 3 --skip-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'])"
done
echo "=== bench full"
timeout -k 10 1500 python bench.py --steps 10 --warmup 3 | tee gpurun_out/bench_full.json
echo "=== randomize refit path"
RMG_PERM_PATH=refit RMG_BENCH_SPIN_S=0 timeout -k 10 600 python bench.py --steps 2 --warmup 3 --perms 32 --parts randomize | tee gpurun_out/bench_refit.json
echo "=== reference arm"
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 | tee gpurun_out/bench_reference.json
echo "=== ncu: launch list of the bench command"
export RMG_BENCH_SPIN_S=0
python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
echo "=== ncu: full capture of the fit kernel (LDG, default)"
ncu --set full --clock-control none --import-source on -k regex:lm_fit -s 3 -c 2 -o gpurun_out/prof_fit_ldg -f \
    python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/ncu_fit_ldg.log 2>&1
echo "ncu fit ldg rc=$?"
echo "=== ncu: full capture of the TMA variant"
RMG_FIT_VARIANT=tma python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/plain_tma.log 2>&1 &&
RMG_FIT_VARIANT=tma ncu --set full --clock-control none --import-source on -k regex:lm_fit -s 3 -c 1 -o gpurun_out/prof_fit_tma -f \
    python bench.py --steps 3 --warmup 3 --skip-extras > gpurun_out/ncu_fit_tma.log 2>&1
echo "ncu fit tma rc=$?"
echo "=== ncu: full capture of the permutation GEMM"
python bench.py --steps 1 --warmup 3 --perms 64 --parts randomize > gpurun_out/plain_perm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:perm_gemm_kernel -s 1 -c 1 -o gpurun_out/prof_perm_gemm -f \
    python bench.py --steps 1 --warmup 3 --perms 64 --parts randomize > gpurun_out/ncu_perm.log 2>&1
echo "ncu perm rc=$?"
ls -la gpurun_out
echo "=== done"
