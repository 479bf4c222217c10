"""Hot SASS lines of one kernel of an ncu --set full report (run here, no GPU needed):
    python tools/ncu_sass_hot.py <report.ncu-rep> <kernel name> [launch index among that kernel = 0] [top = 20]
Prints total samples, warp / thread instructions, average active threads per warp, and the top lines by stall samples."""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if r and r[0].startswith("0x") and len(r) > ix["stall_long_sb"]]
seen, uniq = set(), []
for r in data:                      # the page repeats the listing once per view
    if r[0] in seen:
        continue
    seen.add(r[0])
    uniq.append(r)
data = uniq
S = lambda r, k: int(r[ix[k]] or 0)
tot = sum(S(r, "# Samples") for r in data)
wi = sum(S(r, "Instructions Executed") for r in data)
ti = sum(S(r, "Thread Instructions Executed") for r in data)
print("%s: %d SASS lines, %d samples, %.3g warp instructions, %.3g thread instructions, %.1f active threads / warp" % (kern, len(data), tot, wi, ti, ti / max(1, wi)))
for i, r in sorted(enumerate(data), key=lambda x: -S(x[1], "# Samples"))[:top]:
    print("%5d %6.2f%%  %-58s exec %-11s thr %-3s long_sb %s" % (i, 100.0 * S(r, "# Samples") / max(1, tot), r[ix["Source"]].strip()[:58], r[ix["Instructions Executed"]],
                                                                 r[ix["Avg. Threads Executed"]], r[ix["stall_long_sb"]]))
