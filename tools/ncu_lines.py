"""Per-source-line instruction and stall-sample shares of one kernel of an `ncu --set full` report (run here, no GPU needed).
The report's SASS page has counts per instruction but no line numbers in CSV form; nvdisasm's line info of the same cubin
(built with -lineinfo) supplies them -- both list the function's instructions in the same order.

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <lib .so> <cubin name, e.g. join> [top = 30]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern, lib, cub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
name = rows[0][1]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
seen, sass = set(), []
for r in rows[2:]:
    if r and r[0].startswith("0x") and r[0] not in seen and len(r) > ix["# Samples"]:
        seen.add(r[0])
        sass.append(r)
work = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=work, capture_output=True)
cubin = [f for f in os.listdir(work) if f.startswith(cub + ".")][0]
short = re.sub(r"\(.*", "", name).split("::")[-1]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(work, cubin)], capture_output=True, text=True).stdout
lines, cur_line, cur_file, in_fun = [], None, None, False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        in_fun = short in m.group(1) and (in_fun is False)
        continue
    if not in_fun:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append((cur_file, cur_line))
if len(lines) != len(sass):
    print("warning: %d disassembled instructions, %d in the report" % (len(lines), len(sass)))
agg = {}
S = lambda r, k: int(r[ix[k]] or 0)
for (f, l), r in zip(lines, sass):
    a = agg.setdefault((f, l), [0, 0, 0, 0])
    a[0] += S(r, "Instructions Executed"); a[1] += S(r, "# Samples"); a[2] += S(r, "Thread Instructions Executed"); a[3] += S(r, "stall_long_sb")
ti, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("%s: %.3g warp instructions, %d samples" % (name, ti, ts))
src_cache = {}
def src(f, l):
    for d in ("cgx_b200/csrc", "."):
        p = os.path.join(d, f or "")
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][l - 1].strip()[:100] if l and l <= len(src_cache[p]) else ""
    return ""
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5.1f%% samp %5.1f%% instr  %4.1f thr/warp  %s:%s  %s" % (100.0 * a[1] / max(1, ts), 100.0 * a[0] / max(1, ti), a[2] / max(1, a[0]), f, l, src(f, l)))
