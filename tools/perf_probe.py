"""GPU box: stage timings of the hot path at a given scale (development aid)."""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgx_b200 import synth
from cgx_b200.extractor import GrammarExtractor

ns, nq = int(sys.argv[1]), int(sys.argv[2])
v = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
bq = int(sys.argv[5]) if len(sys.argv) > 5 else nq          # queries per cgx_extract batch
t0 = time.time(); c = synth.generate(ns, nq, v_src=v, v_tgt=v); t1 = time.time(); lay = synth.text_layout(c); t2 = time.time()
print("gen %.1fs layout %.1fs n=%d m=%d T=%d lex=%d" % (t1 - t0, t2 - t1, lay["n"], lay["m"], len(lay["qry_tok"]), len(lay["lex_f"])), flush=True)
ex = GrammarExtractor(0)
t0 = time.time(); info = ex.build_index(lay); print("index build wall %.2fs" % (time.time() - t0), info, flush=True)
ex.profile(True)
for r in range(reps):
    t0 = time.time()
    infos = ex.extract_stream(lay["qry_tok"], lay["qry_off"], bq)
    dt = time.time() - t0
    if len(infos) > 1:
        for i in infos: print("  batch", json.dumps(i), flush=True)
    print("rep %d wall %.3fs -> %.0f q/s (%d batches, device ms %.1f)" % (r, dt, nq / dt, len(infos), sum(i["ms_total"] for i in infos)), json.dumps(infos[-1]), flush=True)
rep = ex.profile_report()
tot = sum(v["ms"] for v in rep.values())
print("profiled kernels: %.1f ms over %d reps" % (tot, reps))
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print("  %-18s launches %5d  ms/rep %9.3f  share %5.1f%%  GB/s %8.1f" % (k, v["launches"], v["ms"] / reps, 100 * v["ms"] / tot, v["bytes"] / 1e6 / max(v["ms"], 1e-9)))
