"""GPU box: SURVEY.md 8(d)'s two extra baselines, both from the reference's own code run on the same B200 host.

  (i)  "ExtractPair.c" CPU leg: createLexiconGappyFast / createLexiconTwoGapFast / createLexiconFast (ExtractPair.c:664,939,515),
       single thread, timed by wall-clock brackets around the three calls (oracle/build_ref_dump.sh inserts them into an
       out-of-tree copy; no reference logic changes);
  (ii) the reference GPU binary's own stderr stage timers (SuffixArray.cu:1339-2243, ExtractPair.cu:3393-3991, SuffixArray.c:238).

    python tools/ref_timers.py <name> <n_sent> <n_qry> <v> [seed=1234] [qry_seed=4321]  ->  gpurun_out/ref_timers_<name>.json

The product CLI runs on the same six files right after it (its per-stage CUDA-event times from the -v log) so that the two
appear side by side.  TEST / MEASUREMENT INFRASTRUCTURE ONLY."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_DUMP_BIN = os.path.join(ROOT, "oracle", "_ref", "strmatchcuda_dump")
CLI = os.path.join(ROOT, "bin", "strmatchcuda")

STAGES = [  # (label, regex on the reference's stderr)
    ("sa_construction_cpu_s", r"SA Construction ([0-9.]+) sec"),
    ("precomputation_s", r"-> Precmputation time: ([0-9.]+)"),
    ("lookup_kernels_s", r"Total Continous Phrase Kernel time: ([0-9.]+) second"),
    ("onegap_enumeration_s", r"-> One Gap Enumeration time: ([0-9.]+)"),
    ("onegap_enumeration_sort_s", r"-> One Gap Enumeration Sorting time: ([0-9.]+)"),
    ("onegap_enumeration_cpu_s", r"-> One Gap Enumeration CPU Processing time: ([0-9.]+)"),
    ("onegap_lookup_kernel_s", r"-> One Gap Look up on SA Kernel time: ([0-9.]+)"),
    ("onegap_lookup_sort_s", r"-> One Gap on SA Sorting Thrust time: ([0-9.]+)"),
    ("onegap_lookup_cpu_s", r"-> One Gap on SA Processing time: ([0-9.]+)"),
    ("twogap_enumeration_s", r"-> Two Gap Enumeration time: ([0-9.]+)"),
    ("twogap_enumeration_sort_s", r"-> Two Gap Enumeration sort time: ([0-9.]+)"),
    ("twogap_enumeration_cpu_s", r"-> Two Gap Enumeration CPU Processing time: ([0-9.]+)"),
    ("twogap_lookup_kernel_s", r"-> Two Gap Look Up on SA Kernel: ([0-9.]+)"),
    ("twogap_lookup_sort_s", r"-> Two Gap on SA Sorting Kernel: ([0-9.]+)"),
    ("twogap_lookup_cpu_s", r"-> Two Gap on SA CPU processing time: ([0-9.]+)"),
    ("extract_contig_kernel_s", r"-> Kernel extractConsistentPairs_Gappy ab, abX, Xab time: ([0-9.]+)"),
    ("extract_twogap_kernel_s", r"-> Kernel extractConsistentPairs Two Gap Seeds aXbXc time: ([0-9.]+)"),
    ("extract_onegap_kernel_s", r"-> Kernel extractConsistentPairs One Gap Seeds aXb time: ([0-9.]+)"),
    ("lexical_task_s", r"-> Lexical Task CPU/GPU time: ([0-9.]+)"),
]


def parse_reference_stderr(text):
    out = {}
    for label, rx in STAGES:
        m = re.findall(rx, text)
        if m:
            out[label] = sum(float(x) for x in m)
    sorts = re.findall(r"-> Gappy Results Sorting Thrust time: ([0-9.]+)", text)
    if sorts:
        out["extract_result_sorts_s"] = sum(float(x) for x in sorts)
    lex = {}
    for name, sec, rec in re.findall(r"cgx_ref_timer (\w+) ([0-9.]+) s (\d+) records", text):
        lex[name] = {"seconds": float(sec), "records": int(rec)}
    if lex:
        out["extractpair_c"] = lex
        out["extractpair_c_total_s"] = sum(v["seconds"] for v in lex.values())
        out["extractpair_c_records"] = sum(v["records"] for v in lex.values())
    return out


def run(name, n_sent, n_qry, v, seed=1234, qry_seed=4321, keep_dir=None):
    from cgx_b200 import synth
    work = keep_dir or tempfile.mkdtemp(prefix="cgx_reft_")
    t0 = time.time()
    c = synth.generate(n_sent, n_qry, v_src=v, v_tgt=v, seed=seed, qry_seed=qry_seed)
    p = synth.write_text(c, work, "corpus")
    gen_s = time.time() - t0
    args = [p["f"], p["q"], p["e"], p["a"], p["lex"]]
    res = {"name": name, "sentence_pairs": c.n_sent, "source_tokens": int(len(c.src_words)), "queries": c.n_qry, "vocabulary": v,
           "host_cores": os.cpu_count(), "text_files_s": gen_s}
    if os.path.exists(REF_DUMP_BIN):
        os.makedirs(os.path.join(work, "ref"), exist_ok=True)
        env = {k: v_ for k, v_ in os.environ.items() if k != "CGX_DUMP_DIR"}
        t0 = time.time()
        r = subprocess.run([REF_DUMP_BIN] + args + [os.path.join(work, "ref")], capture_output=True, text=True, cwd=work, env=env)
        wall = time.time() - t0
        ok = "Start Printing Gappy Phrases" in r.stderr
        ref = parse_reference_stderr(r.stderr + "\n" + r.stdout)
        ref.update({"wall_s": wall, "completed": ok, "binary": "oracle/_ref/strmatchcuda_dump (reference sources + fwrite / timer hooks, CGX_DUMP_DIR unset)"})
        if not ok:
            ref["stderr_tail"] = r.stderr[-1500:]
        res["reference"] = ref
    else:
        res["reference"] = {"unavailable": "oracle/_ref/strmatchcuda_dump not built"}
    os.makedirs(os.path.join(work, "mine"), exist_ok=True)
    t0 = time.time()
    r = subprocess.run([CLI] + args + [os.path.join(work, "mine")], capture_output=True, text=True)
    res["product_cli"] = {"wall_s": time.time() - t0, "rc": r.returncode, "stderr_tail": r.stderr[-1200:]}
    return res


if __name__ == "__main__":
    name = sys.argv[1]
    ns, nq, v = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    seed = int(sys.argv[5]) if len(sys.argv) > 5 else 1234
    qseed = int(sys.argv[6]) if len(sys.argv) > 6 else 4321
    out = run(name, ns, nq, v, seed, qseed)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_timers_%s.json" % name)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))
