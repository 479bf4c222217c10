#!/usr/bin/env python
"""The drop-in command line end to end at the C2 corpus: text files in, grammar files out, file writing included.

    python tools/cli_c2.py [sentence_pairs=1000000] [queries=2000] [writer threads ...=1 8 32]

Writes the synthetic corpus as strmatchcuda text files once, then runs bin/strmatchcuda over it once per writer-thread count
(the first run builds and saves the index with -i, the later ones load it) and once more with -z 1 at the largest count.  Every run's
stderr stage line (loading / index / match+extract / grammar writing / total, cgx_b200/host/run.c) is parsed into
gpurun_out/r2/cli_c2.json together with the wall time, the bytes of grammar text written and the writer's rules per second.
The output directory is removed after each run (a 10 k-query C2 batch is ~5 GB of grammar text)."""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLI = os.path.join(ROOT, "bin", "strmatchcuda")
LINE = re.compile(r"loading ([0-9.]+) s, index ([0-9.]+) s, match\+extract ([0-9.]+) s .*?grammar writing ([0-9.]+) s, total ([0-9.]+) s; (\d+) rules; ([0-9.]+) query")


def tree_bytes(d):
    return sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(d) for f in fs)


def main():
    from cgx_b200 import synth
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    writers = [int(x) for x in sys.argv[3:]] or [1, 8, 32]
    work = tempfile.mkdtemp(prefix="cgx_cli_c2_")
    out = {"sentence_pairs": ns, "queries": nq, "host_cores": os.cpu_count(), "runs": []}
    try:
        t0 = time.time()
        c = synth.generate(ns, nq)
        out["generate_s"] = time.time() - t0
        t0 = time.time()
        p = synth.write_text(c, work, "corpus")
        out["text_files_s"] = time.time() - t0
        out["source_tokens"] = int(len(c.src_words))
        out["input_bytes"] = {k: os.path.getsize(v) for k, v in p.items()}
        del c
        files = [p["f"], p["q"], p["e"], p["a"], p["lex"]]
        index = os.path.join(work, "corpus.idx")
        for w, extra in [(w, []) for w in writers] + [(max(writers), ["-z", "1"])]:
            dest = os.path.join(work, "out")
            os.makedirs(dest)
            had_index = os.path.exists(index)
            t0 = time.time()
            try:
                r = subprocess.run([CLI, "-q", "-i", index, "-w", str(w)] + extra + files + [dest], capture_output=True, text=True, timeout=240)
                rc, err = r.returncode, r.stderr
            except subprocess.TimeoutExpired as e:
                rc, err = -1, "timeout: " + str(e)
            run = {"writer_threads": w, "options": extra, "index_loaded_from_file": had_index, "rc": rc, "wall_s": time.time() - t0,
                   "grammar_bytes": tree_bytes(dest), "grammar_files": sum(len(fs) for _, _, fs in os.walk(dest))}
            m = LINE.search(err)
            if m:
                load, idx, mx, wr, tot, rules, qps = (float(x) for x in m.groups())
                run.update({"loading_s": load, "index_s": idx, "match_extract_s": mx, "grammar_writing_s": wr, "total_s": tot, "rules": int(rules),
                            "queries_per_s_total": nq / tot, "queries_per_s_extract_and_write": qps, "queries_per_s_after_loading": nq / max(mx + wr, 1e-9),
                            "writer_rules_per_s": rules / wr if wr > 0 else None, "writer_bytes_per_s": run["grammar_bytes"] / wr if wr > 0 else None})
            else:
                run["stderr_tail"] = err[-800:]
            out["runs"].append(run)
            print(json.dumps(run), flush=True)
            shutil.rmtree(dest, ignore_errors=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out", "r2"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2", "cli_c2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
