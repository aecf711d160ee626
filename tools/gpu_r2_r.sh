#!/bin/bash
# Round 2, call R (1 GPU): the build as committed -- whole GPU suite, smoke, the default bench line, then ncu --set full of
# the stored-type fit (one FMA per sample, 16 subjects per stage) and of the lazy TFCE sweep.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_r.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
echo "=== pytest -m gpu (all)"
timeout -k 10 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
echo "=== smoke"
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default command)"
timeout -k 10 1500 python bench.py > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err
echo "rc=$?"; tail -c 3000 gpurun_out/r02n_bench_n1.json
export RMG_BENCH_SPIN_S=0
prof() {   # name, kernel regex, launches to keep, matching launches to skip, command...
  local name=$1 regex=$2 count=$3 skip=$4; shift 4
  "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c "$count" -o /tmp/r02n_$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name rc=$?"
  ncu -i /tmp/r02n_$name.ncu-rep --page raw --csv > gpurun_out/r02n_prof_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/r02n_$name.ncu-rep --page source --csv > gpurun_out/r02n_${name}_source.csv 2>/dev/null
  python tools/ncu_source_summary.py gpurun_out/r02n_${name}_source.csv 20 > gpurun_out/r02n_${name}_source_summary.md
  ls -la gpurun_out/r02n_prof_${name}_raw.csv
}
prof stored "lm_fit_packed" 1 3 python bench.py --steps 1 --warmup 3 --parts stored
prof tfce_lazy "tfce_(link_lazy|settle|flow|finish|jump|combine_lazy)" 40 60 python bench.py --steps 1 --warmup 3 --tfce-perms 64 --parts tfce --skip-cpu
echo "=== launch list of the tfce leg"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02n_launches_tfce.csv python bench.py --steps 1 --warmup 1 --tfce-perms 64 --parts tfce --skip-cpu > /dev/null 2>&1
echo "rc=$?"; wc -l gpurun_out/r02n_launches_tfce.csv
echo "=== done"
