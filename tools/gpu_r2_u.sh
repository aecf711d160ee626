#!/bin/bash
# Round 2, call U (1 GPU): mincTFCE on volumes, lazy sweep vs level sweep -- wall clock and launch lists.
set -u
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_u.log) 2>&1
python -c "import __graft_entry__ as g; g.build()" || exit 1
for sw in 0 1; do
  echo "=== RMG_TFCE_SWEEP=$sw"
  RMG_TFCE_SWEEP=$sw RMG_TIMING=1 python tools/vol_tfce_probe.py 3 2>&1 | tail -12
  RMG_TFCE_SWEEP=$sw ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2u_launches_vol_$sw.csv python tools/vol_tfce_probe.py 1 > /dev/null 2>&1
  python - <<PY
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/r2u_launches_vol_$sw.csv")) if len(r) > 10]
ci = {h: i for i, h in enumerate(rows[0])}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]); name = re.sub(r"^void |rmg::|<unnamed>::", "", name)[:60]
    v = float(r[ci["Metric Value"]]); u = r[ci["Metric Unit"]]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[name][0] += 1; agg[name][1] += v
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {k:62s} {n:5d} {v:10.0f} us")
PY
done
echo "=== done"
