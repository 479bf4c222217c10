"""Summarise an `ncu --set full` report into the per-kernel table committed under profiles/ (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/r1/prof.ncu-rep [profiles/out.md] [profiles/dram_traffic.json]
One row per captured launch: duration, DRAM read/write bytes, DRAM / L2 / L1 / SM throughput as % of peak, occupancy,
issue utilisation, registers, top stall reasons (from the source page).  With a third argument the per-launch DRAM bytes
(dram__bytes_read.sum + dram__bytes_write.sum) of each kernel are merged into that JSON (bench.py's roofline.traffic)."""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict

rep = sys.argv[1]
out_md = sys.argv[2] if len(sys.argv) > 2 else None
out_json = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = {}, None          # source pages by kernel name (first captured launch of each)
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        key = re.sub(r"\(.*", "", r[1]).replace("void ", "").replace("cgx::", "")
        cur = [] if key not in blocks else None
        if cur is not None:
            blocks[key] = cur
        continue
    if cur is not None:
        cur.append(r)


def stalls(block):
    if not block:
        return ""
    h, data = block[0], block[1:]
    cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not" not in c]
    i_s = h.index("Warp Stall Sampling (All Samples)")
    tot = sum(int(r[i_s] or 0) for r in data) or 1
    agg = sorted(((sum(int(r[i] or 0) for r in data), h[i][6:]) for i in cols), reverse=True)[:3]
    return ", ".join("%s %.0f%%" % (k, 100.0 * v / tot) for v, k in agg)


def to_bytes(v, u):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


def to_ms(v, u):
    return float(v) * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)


lines = ["| kernel | grid | ms | DRAM rd GB | DRAM wr GB | DRAM % | L2 % | L1 % | SM % | occupancy % | issue % | regs | top stalls |",
         "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|"]
traffic = OrderedDict()
alg = {}                     # algorithmic bytes of the captured launch, where the launch shape gives them (onesweep: 4096 keys per CTA)
for n, r in enumerate(rows[2:]):
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
    g = lambda k: r[idx[k]] if k in idx else "nan"
    ms = to_ms(g("gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]])
    rd = to_bytes(g("dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]])
    wr = to_bytes(g("dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]])
    traffic.setdefault(name, []).append(rd + wr)
    if "rs_onesweep_kernel" in name:
        per_key = 2 * ((8 if "unsigned long" in name.split(",")[0] else 4) + (4 if "(bool)1" in name or ", 1," in name else 0))
        if rd + wr >= max(traffic[name]):
            alg[name] = float(g("launch__grid_size")) * 4096 * per_key
    lines.append("| `%s` | %s | %.3f | %.3f | %.3f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s | %s |" % (
        name, g("launch__grid_size"), ms, rd / 1e9, wr / 1e9, float(g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
        float(g("lts__throughput.avg.pct_of_peak_sustained_elapsed")), float(g("l1tex__throughput.avg.pct_of_peak_sustained_elapsed")),
        float(g("sm__throughput.avg.pct_of_peak_sustained_elapsed")), float(g("sm__warps_active.avg.pct_of_peak_sustained_active")),
        float(g("smsp__issue_active.avg.pct_of_peak_sustained_active")), g("launch__registers_per_thread"), stalls(blocks.get(name.replace("cgx::", "")))))
text = "\n".join(lines)
print(text)
if out_md:
    open(out_md, "a").write(text + "\n")
if out_json:
    try:
        d = json.load(open(out_json))
    except (OSError, ValueError):
        d = {}
    alias = {"j1_scan_kernel": "join_onegap", "j1_pos_kernel": "join_onegap", "j1_pos_ordered_kernel": "join_onegap", "j2_scan_kernel": "join_twogap",
             "j2_ordered_kernel": "join_twogap", "agg_group_kernel": "agg_group", "agg_rules_kernel": "agg_rules",
             "agg_hash_kernel": "agg_hash", "extract_onegap_kernel": "extract_onegap", "extract_contig_kernel": "extract_contig",
             "extract_twogap_kernel": "extract_twogap", "rs_onesweep_kernel": "radix_onesweep", "lookup_kernel": "lookup"}
    for name, v in traffic.items():
        base = name.split("::")[-1].split("<")[0]
        key = alias.get(base, base)
        if key in d and key == "radix_onesweep" and d[key].get("report") == rep.split("/")[-1] and d[key]["dram_bytes_per_launch"] >= max(v):
            continue             # several template instances in one report: keep the largest launch
        d[key] = {"dram_bytes_per_launch": max(v), "launches_captured": len(v), "report": rep.split("/")[-1], "kernel": name}
        if name in alg:
            d[key]["alg_bytes_of_captured_launch"] = alg[name]
    json.dump(d, open(out_json, "w"), indent=1)
