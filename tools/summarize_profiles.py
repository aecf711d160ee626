#!/usr/bin/env python
"""Turn ncu CSV exports into the markdown / JSON summaries committed under profiles/.

  python tools/summarize_profiles.py launches <launches.csv> <title> > profiles/xx_launches.md
  python tools/summarize_profiles.py raw <raw.csv> [kernel-regex]     > table of the key metrics per kernel
  python tools/summarize_profiles.py traffic <raw.csv> <source-label>  > JSON: DRAM bytes per launch per kernel
"""
import csv
import json
import re
import sys
from collections import OrderedDict

KEY = [
    ("duration [ms]", "gpu__time_duration.sum"),
    ("DRAM read [GB]", "dram__bytes_read.sum"),
    ("DRAM write [GB]", "dram__bytes_write.sum"),
    ("DRAM read, % of ncu peak", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed"),
    ("FP64 pipe active %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue slots active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("registers/thread", "launch__registers_per_thread"),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
]
UNIT = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0,
        "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "").replace("rmg::", "").replace("<unnamed>::", "")
    return name.strip()


def load_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def val(row, hdr, units, metric):
    if metric not in hdr:
        return None
    i = hdr.index(metric)
    try:
        x = float(row[i].replace(",", ""))
    except ValueError:
        return None
    return x * UNIT.get(units[i], 1.0) if ("bytes" in metric or "duration" in metric) else x


def cmd_raw(path, pattern=".*"):
    hdr, units, rows = load_raw(path)
    seen = OrderedDict()
    for r in rows:
        k = short(r[hdr.index("Kernel Name")])
        if re.search(pattern, k) and k not in seen:
            seen[k] = r
    names = list(seen)
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for label, metric in KEY:
        cells = []
        for n in names:
            x = val(seen[n], hdr, units, metric)
            cells.append("n/a" if x is None else (f"{x:.6g}"))
        print(f"| {label} | " + " | ".join(cells) + " |")


def cmd_traffic(path, source):
    hdr, units, rows = load_raw(path)
    out = OrderedDict()
    for r in rows:
        k = re.sub(r"<.*", "", short(r[hdr.index("Kernel Name")]))
        if k in out:
            continue
        rd, wr, du = (val(r, hdr, units, m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
        out[k] = {"dram_bytes_read": rd * 1e9, "dram_bytes_write": wr * 1e9, "duration_ms_under_ncu": du, "source": source}
    print(json.dumps(out, indent=1))


def cmd_launches(path, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        try:
            ms = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        except ValueError:
            continue
        k = short(r[ki])
        k = k if len(k) <= 70 else "..." + k[-67:]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    MINE = r"^(lm_|mask_|perm_|pack_|fdr_|tfce_|extract_maps|keys_|colmax)"
    ours = sum(a[1] for k, a in agg.items() if re.search(MINE, k))
    print(f"# ncu launch list: {title}\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none --csv`; per-launch times are cold-cache and serialised, "
          "compare SHARES.  `at::` / `native::` kernels are torch generating the synthetic inputs (outside every timed region).\n")
    print("| kernel | launches | total ms | share of all | share of this library's kernels |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, ms) in agg.items():
        mine = re.search(MINE, k)
        print(f"| `{k}` | {c} | {ms:.3f} | {100 * ms / total:.1f}% | {(f'{100 * ms / ours:.1f}%' if mine else '-')} |")


if __name__ == "__main__":
    {"raw": cmd_raw, "traffic": cmd_traffic, "launches": cmd_launches}[sys.argv[1]](*sys.argv[2:])
