"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into per-kernel totals.
    python tools/summarize_launches.py profiles/<launches>.csv [first_id last_id]
Prints a markdown table: kernel, launches, total ms, share.  The optional id range restricts the summary to one
step of the hot path (ids are the first CSV column)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else -1
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.reader(lines):
    if r[0] == "ID":
        continue
    try:
        i = int(r[0])
    except ValueError:
        continue
    if not (lo <= i <= hi):
        continue
    name = re.sub(r"\(.*", "", r[4])
    name = re.sub(r"^void ", "", name)
    rows.append((i, name, r[8], r[7], float(r[-1]) / 1e6))
tot = sum(x[-1] for x in rows)
agg = defaultdict(lambda: [0, 0.0])
for _, name, grid, block, ms in rows:
    agg[name][0] += 1
    agg[name][1] += ms
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f%% |" % (name, c, ms, 100 * ms / tot))
print("| **all** | %d | %.3f | 100%% |" % (len(rows), tot))
