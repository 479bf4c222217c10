#!/usr/bin/env python
"""bench.py -- query sentences/s of the grammar-extraction hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (default `c2` = BASELINE.json configs[1]): synthetic 1 M sentence-pair Zipfian parallel corpus
(V = 50 k, word alignments with 10 % dropped links), 10 k query sentences per GPU, max phrase length 5, rule
shapes of the reference (one- and two-gap), sample sizes 300/65/70.  One *step* = one pass of the hot path
(contiguous lookup -> pattern enumeration -> gappy band joins -> extraction -> aggregation + lexical
scoring) over the rank's query batch against the resident index.

  value  : whole-job query sentences/s, queries resident in HBM, results left in HBM (cgx_extract_dev),
           CUDA-event time of the K steps, max over ranks.
  e2e    : the same through the C ABI with HOST buffers, as a caller with a stream of batches uses it
           (cgx_extract_begin + cgx_result_at, the path bin/strmatchcuda takes): every step copies its queries H2D
           from pinned memory and every rule / pattern table D2H; the D2H tail of step i overlaps the kernels of
           step i+1 (results rotate over three buffer sets), the last step's tail is waited for inside the timed
           region.  Wall clock over the K steps with a device synchronise on both sides, max over ranks.
           e2e.single_batch_ms is the unpipelined latency of one batch (cgx_extract returns when the last
           result byte is on the host).
  roofline : the dominant kernel of the step (per-kernel CUDA events on the launching stream,
           cgx_profile_*), algorithmic bytes as defined in DESIGN.md, peak from MEASURED_PEAKS.json.
  cpu_baseline : the CPU oracle port (oracle/cgx_oracle.c, single thread) on a bounded sample of the same
           queries against the same corpus, SA from the reference's own SuffixArray.c (oracle/_ref/libref_sa.so).
  --impl reference : the CPU implementation of the path on all host cores (the reference has no CPU
           matcher/extractor -- SURVEY.md section 0 -- so this is the oracle port, one process per core over a
           bounded query sample; its suffix array is built by the reference's SuffixArray.c).

  strong : a FIXED set of 80 k queries of the same corpus, cut into contiguous token-balanced shards (cgx_b200.dist.shard_queries),
           every rank streaming its shard through cgx_extract_begin in 10 k-query batches; whole-set wall clock, max over ranks.
           The weak-scaling `value` keeps 10 k queries per GPU; this block shows what a fixed job gains from N GPUs.
  c3     : (N = 1 only) BASELINE.json configs[2] -- 10 M sentence pairs (~260 M tokens), 100 k queries -- as a second block: index
           build, then all queries streamed in batches the way bin/strmatchcuda cuts them (a batch whose hit lists outgrow the
           31-bit result indices is refused and halved); device-resident and end-to-end rates, per-kernel table.  --no-c3 skips it.

  sweep  : BASELINE.json configs[4] -- the first 1 k / 10 k / 100 k / 1 M queries of one fixed set of the same corpus, sharded
           over the ranks like `strong` and streamed in 10 k-query batches; end-to-end wall clock (host buffers, every result
           array copied D2H), max over ranks.  --no-sweep skips it.
  gpu_reference : (N = 1 only) SURVEY.md 8(d)'s two reference legs, from the reference's own code on this box (tools/ref_timers.py,
           oracle/_ref/strmatchcuda_dump = reference sources + timer hooks): its GPU binary's stage timers and its host
           aggregation createLexicon*Fast (ExtractPair.c:515,664,939; one thread) on the largest configuration it survives here
           (100 k sentence pairs, 60 queries: it faults at C1 and C2 sizes), with the product's CUDA-event stage times and the
           drop-in CLI's wall clock (text files in, grammar files out) on the same six files.  --no-gpu-reference skips it.

Multi-GPU (weak scaling): the corpus and every rank's queries are generated once, on rank 0; the index is built once on rank 0
and broadcast over NVLink with NCCL; every rank then processes its own 10 k-query batch; there is no data-path collective.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (sentence pairs, queries per GPU, vocabulary, description)
    "c2": (1_000_000, 10_000, 50_000, "synthetic 1M sentence-pair Zipfian corpus (V=50k), 10k queries, max phrase len 5 (BASELINE.json configs[1])"),
    "c1": (10_000, 100, 2_000, "synthetic toy-scale corpus: 10k sentence pairs (V=2k), 100 queries (stand-in for the unshipped toy/ corpus, configs[0])"),
    "mid": (200_000, 2_000, 30_000, "synthetic 200k sentence-pair corpus (V=30k), 2k queries (development size)"),
    "c3": (10_000_000, 100_000, 50_000, "synthetic 10M sentence-pair corpus (~260M tokens, V=50k), 100k queries, one- and two-gap rules (BASELINE.json configs[2])"),
}
STRONG_QUERIES = 80_000          # fixed query set of the strong-scaling block (8 shards of one 10 k batch at N = 8)
STRONG_SEED = 8765
BATCH_QUERIES = 10_000           # queries per batch of a stream (run.c's CGXH_DEFAULT_BATCH)
SWEEP_QUERIES = (1_000, 10_000, 100_000, 1_000_000)      # BASELINE.json configs[4]: query sweep 1 k .. 1 M sentences
SWEEP_SEED = 2468


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_inputs(workload: str, rank: int):
    from cgx_b200 import synth
    ns, nq, v, _ = WORKLOADS[workload]
    t0 = time.time()
    c = synth.generate(ns, nq, v_src=v, v_tgt=v, seed=1234, qry_seed=4321 + rank)
    lay = synth.text_layout(c)
    log("[rank %d] synthetic corpus: n=%d source tokens, m=%d target tokens, Q=%d queries (T=%d tokens), lex=%d entries, %.1f s"
        % (rank, lay["n"], lay["m"], len(lay["qry_off"]) - 1, len(lay["qry_tok"]), len(lay["lex_f"]), time.time() - t0))
    return lay


def make_query_sets(workload: str, lay, world: int):
    """Rank 0: the query batch of every rank (rank r: seed 4321 + r, as make_inputs(workload, r) would give it) and the fixed
    strong-scaling set, as ids of the corpus vocabulary.  The corpus itself is not generated again."""
    from cgx_b200 import synth
    ns, nq, v, _ = WORKLOADS[workload]
    weak = [(np.ascontiguousarray(lay["qry_tok"], dtype=np.int32), np.ascontiguousarray(lay["qry_off"], dtype=np.int32))]
    for r in range(1, world):
        w, off = synth.generate_queries(ns, nq, v_src=v, v_tgt=v, seed=1234, qry_seed=4321 + r)
        weak.append((synth.query_ids(lay["src_names"], w), off.astype(np.int32)))
    w, off = synth.generate_queries(ns, STRONG_QUERIES, v_src=v, v_tgt=v, seed=1234, qry_seed=STRONG_SEED)
    strong = (synth.query_ids(lay["src_names"], w), off.astype(np.int32))
    w, off = synth.generate_queries(ns, max(SWEEP_QUERIES), v_src=v, v_tgt=v, seed=1234, qry_seed=SWEEP_SEED)
    sweep = (synth.query_ids(lay["src_names"], w), off.astype(np.int32))
    return {"weak": weak, "strong": strong, "sweep": sweep, "n": int(lay["n"])}


def config_of(workload: str, n_tokens: int, Q: int, T: int, world: int):
    """`config` of the JSON line -- the same dictionary from both arms (the driver compares them)."""
    ns, nq, v, desc = WORKLOADS[workload]
    return {"workload": desc, "sentence_pairs": ns, "source_tokens": int(n_tokens), "queries_per_gpu": int(Q), "query_tokens_per_gpu": int(T),
            "rule_shapes": "ab, Xab, abX, XabX, aXb, XaXb, aXbX, aXbXc (the reference's full set; superset of the quoted 1-gap)",
            "samples": [300, 65, 70], "parallelism": "query-sharded x%d, index broadcast once" % world,
            "l2": "explicit 256 MB flush between steps; the index working set (88 B per token pair) is far larger than the 126 MB L2"}


# ----------------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], None, set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                mhz, mx = float(p[0]), float(p[1])
            except ValueError:
                continue
            smax = mx
            if t_begin - 0.05 <= ts <= t_end + 0.05:
                sm.append(mhz)
                try:
                    power.append(float(p[2]))
                except ValueError:
                    pass
                for nm, val in zip(names, p[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (test infrastructure) timed as the reported baseline
# ----------------------------------------------------------------------------------------------------------
def cpu_suffix_array(lay):
    """Suffix array on the CPU by the reference's own SuffixArray.c (oracle/_ref/libref_sa.so); falls back to the
    oracle's prefix-doubling port when that library was not built.  Returns (sa, seconds, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import REF_SA_PATH
    s = np.ascontiguousarray(lay["str"], dtype=np.int32)
    n = int(lay["n"])
    if os.path.exists(REF_SA_PATH):
        L = C.CDLL(REF_SA_PATH)
        L.ref_sa_build.restype = C.c_double
        L.ref_sa_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        sa = np.empty(n, dtype=np.int32)
        devnull = os.open(os.devnull, os.O_WRONLY)      # SuffixArray.c prints progress on stderr/stdout
        so, se = os.dup(1), os.dup(2)
        os.dup2(devnull, 1), os.dup2(devnull, 2)
        try:
            sec = L.ref_sa_build(s.ctypes.data, n, int(s[n - 1]), sa.ctypes.data, None)
        finally:
            os.dup2(so, 1), os.dup2(se, 2)
            os.close(devnull), os.close(so), os.close(se)
        return sa, float(sec), "reference"
    return None, None, "port"


def make_oracle(lay, sa):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle import Oracle
    o = Oracle.from_layout(lay)
    t0 = time.time()
    if sa is not None:
        o.set_sa(sa)
        sa_s = None
    else:
        o.build_sa()
        sa_s = time.time() - t0
    # one-time index-side work of the oracle (frequent-pair precomputation, SuffixArray.cu:1132-1340): untimed
    o.run(lay["qry_tok"][:1], np.array([0, 1], dtype=np.int32))
    return o, sa_s


def oracle_time_queries(o, lay, q0, q1):
    off, tok = lay["qry_off"], lay["qry_tok"]
    t0 = time.perf_counter()
    o.run(tok[off[q0]:off[q1]], (off[q0:q1 + 1] - off[q0]).astype(np.int32))
    return time.perf_counter() - t0


def cpu_baseline_single(lay, budget_s=25.0):
    """Single-thread oracle on the first queries of the batch, growing the sample until ~budget_s is spent."""
    sa, sa_sec, sa_kind = cpu_suffix_array(lay)
    o, port_sa_s = make_oracle(lay, sa)
    Q = len(lay["qry_off"]) - 1
    nq, spent, done_q, done_t = 1, 0.0, 0, 0.0
    while spent < budget_s and nq <= Q:
        dt = oracle_time_queries(o, lay, 0, nq)
        spent += dt
        done_q, done_t = nq, dt
        if dt * 2.2 + spent > budget_s * 1.6:
            break
        nq *= 2
    o.close()
    return {"value": done_q / done_t, "unit": "query sentences/s", "cores": 1, "kind": "port",
            "sample": "first %d of %d queries as one batch, %.1f s of single-thread oracle work (oracle/cgx_oracle.c); per-query CPU cost "
                      "falls with batch size because patterns are de-duplicated across queries" % (done_q, Q, done_t),
            "sa_build_s": sa_sec if sa_sec is not None else port_sa_s, "sa_build_kind": sa_kind}


def reference_arm(args):
    """bench.py --impl reference: the CPU implementation of the path on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lay = make_inputs(args.workload, 0)
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    per_worker = args.cpu_queries_per_core
    sa, sa_sec, sa_kind = cpu_suffix_array(lay)
    log("CPU suffix array (%s): %s s" % (sa_kind, "%.2f" % sa_sec if sa_sec else "n/a"))
    o, port_sa_s = make_oracle(lay, sa)
    Q = len(lay["qry_off"]) - 1
    workers = min(workers, max(1, Q // per_worker))
    sample_q = workers * per_worker

    def one_step():
        """fork one process per core; each runs the oracle on its own slice of the sample (index shared copy-on-write)."""
        t0 = time.perf_counter()
        pids = []
        for w in range(workers):
            pid = os.fork()
            if pid == 0:
                try:
                    oracle_time_queries(o, lay, w * per_worker, (w + 1) * per_worker)
                    os._exit(0)
                except BaseException:
                    os._exit(1)
            pids.append(pid)
        ok = True
        for pid in pids:
            _, st = os.waitpid(pid, 0)
            ok &= (st == 0)
        if not ok:
            raise RuntimeError("an oracle worker failed")
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        one_step()
    times = [one_step() for _ in range(args.steps)]
    total = sum(times)
    value = sample_q * args.steps / total
    sample = ("%d queries per step (%d processes x %d), each process one oracle batch against the full corpus index; "
              "the reference has no CPU matcher/extractor, so the path is the oracle port; suffix array by the reference's SuffixArray.c"
              % (sample_q, workers, per_worker))
    print(json.dumps({
        "impl": "reference", "metric": "query sentences/sec (grammar extraction)", "value": value, "unit": "query sentences/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_of(args.workload, int(lay["n"]), Q, len(lay["qry_tok"]), args.gpus),
        "cpu_baseline": {"value": value, "unit": "query sentences/s", "cores": workers, "kind": "port", "sample": sample, "queries_per_step": sample_q,
                         "sa_build_s": sa_sec if sa_sec is not None else port_sa_s, "sa_build_kind": sa_kind},
        "e2e": {"value": value, "unit": "query sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def result_bytes(ex):
    """D2H bytes of one cgx_extract batch (every array cgx_result exposes)."""
    from cgx_b200._lib import Result
    r = Result()
    ex.L.cgx_result(ex.h, C.byref(r))
    return result_bytes_of(r)


def result_bytes_of(r):
    q1 = int(np.ctypeslib.as_array(r.q1_off, shape=(r.Q + 1,))[-1]) if r.Q else 0
    q2 = int(np.ctypeslib.as_array(r.q2_off, shape=(r.Q + 1,))[-1]) if r.Q else 0
    b = 4 * (r.T * 5 + r.G * 4 + r.D1 * 4 + r.D2 * 2 + 2 * (r.Q + 1) + q1 + q2)
    for k in range(3):
        b += 16 * r.n_rules[k] + (4 + 4) * r.n_ids[k]
    return int(b)


def touch(r):
    """Device->host read of a step's result: the rule counts and the last rule of every kind, from the host mirrors."""
    tot = 0
    for k in range(3):
        n = int(r.n_rules[k])
        tot += n
        if n:
            last = (C.c_char * 16).from_address(r.rules[k] + 16 * (n - 1))
            tot += last.raw[0] & 0
    return tot


def run_c3(args, device):
    """Second block: BASELINE.json configs[2].  10 M sentence pairs (~260 M source tokens), 100 k queries streamed in 10 k-query
    batches through cgx_extract_begin (a batch whose hit lists exceed the 31-bit result indices is refused and halved, as
    bin/strmatchcuda does).  One untimed warm-up batch (buffer growth), then one timed pass over all queries."""
    import torch
    from cgx_b200 import synth
    from cgx_b200.extractor import GrammarExtractor
    ns, nq, v, desc = WORKLOADS["c3"]
    nq = min(nq, args.c3_queries)
    t0 = time.time()
    c = synth.generate(ns, nq, v_src=v, v_tgt=v, seed=1234, qry_seed=4321)
    lay = synth.text_layout(c)
    del c
    gen_s = time.time() - t0
    tok = np.ascontiguousarray(lay["qry_tok"], dtype=np.int32)
    off = np.ascontiguousarray(lay["qry_off"], dtype=np.int32)
    Q, T = len(off) - 1, len(tok)
    log("c3: n=%d source tokens, Q=%d queries (T=%d tokens), lex=%d entries, generated in %.1f s" % (lay["n"], Q, T, len(lay["lex_f"]), gen_s))
    ex = GrammarExtractor(device)
    t0 = time.time()
    info = ex.build_index(lay)
    index_wall = time.time() - t0
    n_tokens, src_max = int(lay["n"]), int(lay["str"][: lay["n"]].max())
    log("c3 index: SA %.1f ms (%d rounds), auxiliary %.1f ms, %.2f GB resident, wall %.2f s" % (info["sa_build_ms"], info["sa_rounds"], info["aux_build_ms"],
                                                                                           info["index_bytes"] / 1e9, index_wall))
    del lay
    # warm-up: the device buffers and the pinned mirrors of all three result sets grow to batch size (pinning ~2 GB of host
    # memory per set takes over a second); the first 10 k-query batch is refused and the advice settles on the batch size
    w1 = min(Q, 2 * BATCH_QUERIES)
    ex.extract_stream(tok[: off[w1]], off[: w1 + 1], batch_queries=BATCH_QUERIES)
    acc = {"rules": 0, "d2h": 0, "batches": 0}

    def on_batch(a, b, r):
        acc["rules"] += touch(r)
        acc["d2h"] += result_bytes_of(r)
        acc["batches"] += 1
    ex.profile(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    infos = ex.extract_stream(tok, off, batch_queries=BATCH_QUERIES, on_batch=on_batch)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    prof = ex.profile_report()
    ex.profile(False)
    ex.close()
    dev_ms = sum(i["ms_total"] for i in infos)
    kern = {k: {"ms": round(x["ms"], 3), "launches": x["launches"], "alg_bytes": x["bytes"], "gbs": (x["bytes"] / 1e9) / (x["ms"] / 1e3) if x["ms"] > 0 else None}
            for k, x in prof.items()}
    tot = {k: int(sum(i[k] for i in infos)) for k in ("hits1", "hits2", "samples", "n_ab", "n_1gap", "n_2gap", "launches")}
    stages = {k: round(float(sum(i[k] for i in infos)), 3) for k in ("ms_lookup", "ms_enum", "ms_join", "ms_extract", "ms_aggregate")}
    tot["rules"] = [int(sum(i["rules"][k] for i in infos)) for k in range(3)]
    return {"workload": desc, "sentence_pairs": ns, "source_tokens": n_tokens, "source_vocabulary": src_max - 2, "queries": Q, "query_tokens": T,
            "value": Q / (dev_ms / 1e3), "unit": "query sentences/s", "device_ms_total": dev_ms,
            "e2e": {"value": Q / wall, "unit": "query sentences/s", "wall_s": wall, "h2d_bytes": 4 * (2 * T + Q + len(infos)), "d2h_bytes": acc["d2h"],
                    "rules_read_back": acc["rules"], "api": "cgx_extract_begin + cgx_result_at per batch, host buffers"},
            "batches": len(infos), "batch_queries": BATCH_QUERIES, "largest_batch_queries": max(i["q1"] - i["q0"] for i in infos),
            "smallest_batch_queries": min(i["q1"] - i["q0"] for i in infos),
            "sa_build": {"gpu_ms": info["sa_build_ms"], "rounds": info["sa_rounds"], "key_bits": info["sa_key_bits"], "aux_index_ms": info["aux_build_ms"]},
            "index_bytes": int(info["index_bytes"]), "index_wall_s": index_wall, "corpus_generation_s": gen_s, "totals": tot, "stage_ms": stages,
            "per_batch": [{k: i[k] for k in ("q0", "q1", "hits1", "hits2", "ms_total", "ms_join", "ms_extract", "ms_aggregate")} for i in infos], "kernels": kern}


REFERENCE_SURVIVES = ("mid100k", 100_000, 60, 20_000)      # name, sentence pairs, queries, vocabulary


def gpu_reference_block(device):
    """The reference's own binary and host aggregation next to the product, same files, same B200 (see the module docstring)."""
    import re
    import shutil
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_timers
    from cgx_b200.extractor import GrammarExtractor
    from cgx_b200.host import HostCorpus
    name, ns, nq, v = REFERENCE_SURVIVES
    work = tempfile.mkdtemp(prefix="cgx_gpuref_")
    try:
        r = ref_timers.run(name, ns, nq, v, keep_dir=work)
        ref = r.get("reference", {})
        out = {"config": "synthetic %d sentence pairs (V=%d), %d queries: the largest configuration the unmodified reference survives on this box "
                         "(it faults in oneGapLookUpSA at the C1 stand-in and cannot hold a C2 batch in its 60 M-hit preallocation, ComTypes.h:56)" % (ns, v, nq),
               "source_tokens": r.get("source_tokens"), "queries": r.get("queries"), "reference": ref}
        cli = r.get("product_cli", {})
        m = re.search(r"loading ([0-9.]+) s, index ([0-9.]+) s, match\+extract ([0-9.]+) s .*?grammar writing ([0-9.]+) s, total ([0-9.]+) s; (\d+) rules", cli.get("stderr_tail", ""))
        prod = {"cli_wall_s": cli.get("wall_s"), "cli_rc": cli.get("rc")}
        if m:
            prod.update({"cli_loading_s": float(m.group(1)), "cli_index_s": float(m.group(2)), "cli_match_extract_s": float(m.group(3)),
                         "cli_grammar_writing_s": float(m.group(4)), "cli_total_s": float(m.group(5)), "cli_rules": int(m.group(6))})
            if float(m.group(4)) > 0:
                prod["cli_writer_rules_per_s"] = int(m.group(6)) / float(m.group(4))
        # the product's per-stage CUDA-event times on the same files (second batch: buffers warm)
        hc = HostCorpus(*(os.path.join(work, "corpus." + e) for e in ("f", "q", "e", "a", "lex")))
        lay = hc.layout()
        ex = GrammarExtractor(device)
        info = ex.build_index(lay)
        ex.extract(lay["qry_tok"], lay["qry_off"], fetch=False)
        b = ex.extract(lay["qry_tok"], lay["qry_off"], fetch=False)
        ex.close()
        prod.update({"sa_build_ms": info["sa_build_ms"], "aux_index_ms": info["aux_build_ms"]})
        prod.update({k: b[k] for k in ("ms_total", "ms_lookup", "ms_enum", "ms_join", "ms_extract", "ms_aggregate", "hits1", "hits2", "samples", "rules")})
        out["product"] = prod
        if ref.get("completed"):
            g = lambda *ks: sum(ref.get(k, 0.0) for k in ks)
            pairs = {"sa_build": (g("sa_construction_cpu_s"), info["sa_build_ms"] / 1e3),
                     "lookup": (g("lookup_kernels_s"), b["ms_lookup"] / 1e3),
                     "enumerate+join (kernels, sorts, host scans)": (g("precomputation_s", "onegap_enumeration_s", "onegap_enumeration_sort_s", "onegap_enumeration_cpu_s",
                                                                       "onegap_lookup_kernel_s", "onegap_lookup_sort_s", "onegap_lookup_cpu_s", "twogap_enumeration_s",
                                                                       "twogap_enumeration_sort_s", "twogap_enumeration_cpu_s", "twogap_lookup_kernel_s", "twogap_lookup_sort_s",
                                                                       "twogap_lookup_cpu_s"), (b["ms_enum"] + b["ms_join"]) / 1e3),
                     "extract (kernels + result sorts)": (g("extract_contig_kernel_s", "extract_twogap_kernel_s", "extract_onegap_kernel_s", "extract_result_sorts_s"), b["ms_extract"] / 1e3),
                     "aggregate + lexical scoring (ExtractPair.c createLexicon*Fast + lexical task)": (g("extractpair_c_total_s", "lexical_task_s"), b["ms_aggregate"] / 1e3)}
            out["stages_reference_s_vs_product_s"] = {k: {"reference_s": a, "product_s": c, "ratio": (a / c if c > 0 else None)} for k, (a, c) in pairs.items()}
            if ref.get("wall_s") and prod.get("cli_wall_s"):
                out["cli_wall_ratio"] = ref["wall_s"] / prod["cli_wall_s"]
        return out
    finally:
        shutil.rmtree(work, ignore_errors=True)


def bind_to_gpu_numa_node(local, rank):
    """Pin this rank's host threads to the CPUs next to its GPU (NVML's ideal affinity) before anything is allocated, so that
    the pinned result mirrors (1.8 GB per batch and set) live on the GPU's own NUMA node.  On the single-socket 8-GPU VM
    of this pool it changes nothing (one NUMA node: 8 ranks share ~90 GB/s of host ingest either way); it matters on
    two-socket hosts."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        cpus = sorted(os.sched_getaffinity(0))
        log("[rank %d] bound to %d CPUs next to GPU %d (%d..%d)" % (rank, len(cpus), local, cpus[0], cpus[-1]))
    except Exception as e:  # binding is an optimisation, never a requirement
        log("[rank %d] CPU affinity not set: %r" % (rank, e))


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from cgx_b200 import dist as cdist
    from cgx_b200.extractor import GrammarExtractor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch --gpus %d through torch.distributed.run (one rank per GPU)" % args.gpus)
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind_to_gpu_numa_node(local, rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # corpus and every rank's queries: generated once, on rank 0 (round 1 regenerated the 1 M-pair corpus on every rank)
    lay, qsets = None, None
    if rank == 0:
        lay = make_inputs(args.workload, 0)
        t0 = time.time()
        qsets = make_query_sets(args.workload, lay, world)
        log("query sets for %d rank(s) + %d-query strong-scaling set: %.1f s" % (world, STRONG_QUERIES, time.time() - t0))
    if world > 1:
        box = [qsets]
        dist.broadcast_object_list(box, src=0)
        qsets = box[0]
    q_tok, q_off = qsets["weak"][rank]
    Q, T = len(q_off) - 1, len(q_tok)
    ex = GrammarExtractor(local)             # raises when the CUDA library / device is missing: there is no fallback
    bcast_ms, bcast_bytes = 0.0, 0
    if rank == 0:
        t0 = time.time()
        info = ex.build_index(lay)
        log("index: SA %.1f ms (%d doubling rounds, %d-bit keys), auxiliary %.1f ms, %.2f GB resident, wall %.2f s"
            % (info["sa_build_ms"], info["sa_rounds"], info["sa_key_bits"], info["aux_build_ms"], info["index_bytes"] / 1e9, time.time() - t0))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bcast_bytes = cdist.broadcast_index(ex, src=0)
        torch.cuda.synchronize()
        dist.barrier()
        bcast_ms = 1e3 * (time.perf_counter() - t0)
        info = ex.index_info()
        log("[rank %d] index broadcast over NCCL: %.1f MB in %.1f ms" % (rank, bcast_bytes / 1e6, bcast_ms))

    # ---- SA build ms (second headline metric): tokens resident in HBM -> sa[] in HBM, warm (1 untimed + 3 timed builds)
    sa_ms = None
    if rank == 0:
        str_d = torch.from_numpy(np.ascontiguousarray(lay["str"], dtype=np.int32)).to(dev)
        sa_d = torch.empty(int(lay["n"]), dtype=torch.int32, device=dev)
        ex.sa_build_dev(str_d.data_ptr(), int(lay["n"]), int(lay["str"][: lay["n"]].max()), sa_d.data_ptr())
        runs = [ex.sa_build_dev(str_d.data_ptr(), int(lay["n"]), int(lay["str"][: lay["n"]].max()), sa_d.data_ptr())[1] for _ in range(3)]
        sa_ms = float(np.mean(runs))
        log("SA build (warm, tokens in HBM): %.2f ms mean of %s" % (sa_ms, ["%.2f" % r for r in runs]))
        del str_d, sa_d

    # device-resident query batch (value) and pinned host copies (e2e)
    qt_h = torch.from_numpy(np.ascontiguousarray(q_tok, dtype=np.int32)).pin_memory()
    qo_h = torch.from_numpy(np.ascontiguousarray(q_off, dtype=np.int32)).pin_memory()
    t2q = np.repeat(np.arange(Q, dtype=np.int32), np.diff(q_off))
    qt_d, qo_d, t2q_d = qt_h.to(dev), qo_h.to(dev), torch.from_numpy(t2q).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    qt_np, qo_np = qt_h.numpy(), qo_h.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def step_dev():
        flush.zero_()                                    # evict L2 between steps (outside the timed events)
        torch.cuda.synchronize()
        return ex.extract_dev(qt_d.data_ptr(), qo_d.data_ptr(), t2q_d.data_ptr(), Q, T)

    def step_e2e():
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        binfo = ex.extract(qt_np, qo_np, fetch=False)    # cgx_extract: returns after every result array is on the host
        return time.perf_counter() - t0, binfo

    def run_e2e_pipelined(steps):
        """K batches through cgx_extract_begin: step i's D2H tail travels while step i+1 computes."""
        n_rules = 0
        for i in range(steps):
            flush.zero_()                                # L2 eviction stays between the steps, inside the timed region
            torch.cuda.current_stream().synchronize()    # (only torch's stream: the copy stream keeps travelling)
            ex.extract_begin(qt_np, qo_np)
            if i > 0:
                n_rules += touch(ex.result_at(1, raw=True))
        n_rules += touch(ex.result_at(0, raw=True))
        return n_rules

    for _ in range(args.warmup):
        step_dev()
    for _ in range(args.warmup):
        step_e2e()
    run_e2e_pipelined(max(args.warmup, 3))               # all three result sets allocated (device + pinned) before timing

    sampler = ClockSampler(local) if rank == 0 else None
    # ---- timed: device-resident -------------------------------------------------------------------
    ex.profile(True)
    barrier()
    t_begin = time.time()
    w0 = time.perf_counter()
    dev_ms, binfo = 0.0, None
    for _ in range(args.steps):
        binfo = step_dev()
        dev_ms += binfo["ms_total"]
    barrier()
    wall_dev = time.perf_counter() - w0
    prof = ex.profile_report()
    ex.profile(False)
    # ---- timed: end to end through the C ABI with host buffers ------------------------------------------
    barrier()
    w0 = time.perf_counter()
    e2e_rules = run_e2e_pipelined(args.steps)
    barrier()
    e2e_s = time.perf_counter() - w0
    d2h = result_bytes(ex)
    lat_s = 0.0
    for _ in range(args.steps):
        dt, einfo = step_e2e()
        lat_s += dt
    barrier()
    # ---- timed: strong scaling -- the fixed 80 k-query set, sharded (one warm pass, then `strong_passes` timed ones) ----------
    stok, soff = qsets["strong"]
    sq0, sq1 = cdist.shard_queries(soff, world, rank)
    my_tok = np.ascontiguousarray(stok[soff[sq0]:soff[sq1]])
    my_off = np.ascontiguousarray(soff[sq0:sq1 + 1] - soff[sq0])
    strong_bytes = [0, 0]

    def strong_pass():
        rules = [0]

        def on_batch(a, b, r):
            rules[0] += touch(r)
            strong_bytes[0] += result_bytes_of(r)
            strong_bytes[1] += 1
        barrier()
        t0 = time.perf_counter()
        infos = ex.extract_stream(my_tok, my_off, batch_queries=BATCH_QUERIES, on_batch=on_batch)
        torch.cuda.synchronize()
        return time.perf_counter() - t0, sum(i["ms_total"] for i in infos), len(infos), rules[0]

    strong_pass()
    strong_bytes = [0, 0]
    strong_runs = [strong_pass() for _ in range(args.strong_passes)]
    strong_s = sum(r[0] for r in strong_runs) / len(strong_runs)
    strong_dev_ms = sum(r[1] for r in strong_runs) / len(strong_runs)
    barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    # ---- query sweep (configs[4]): prefixes of one fixed set, sharded like the strong block
    sweep_rows = []
    if not args.no_sweep and "sweep" in qsets:
        wtok, woff = qsets["sweep"]
        for nq in SWEEP_QUERIES:
            nq = min(nq, len(woff) - 1)             # (the generator drops empty sentences: the 1 M set is a few short)
            sub_off = woff[: nq + 1]
            a0, a1 = cdist.shard_queries(sub_off, world, rank)
            mt = np.ascontiguousarray(wtok[sub_off[a0]:sub_off[a1]])
            mo = np.ascontiguousarray(sub_off[a0:a1 + 1] - sub_off[a0])
            seen = [0]

            def on_sweep_batch(a, b, r, seen=seen):
                seen[0] += touch(r)
            barrier()
            t0 = time.perf_counter()
            infos = ex.extract_stream(mt, mo, batch_queries=BATCH_QUERIES, on_batch=on_sweep_batch) if a1 > a0 else []
            torch.cuda.synchronize()
            tw = torch.tensor([time.perf_counter() - t0, sum(i["ms_total"] for i in infos) / 1e3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            sweep_rows.append({"queries": nq, "value": nq / float(tw[0]), "unit": "query sentences/s", "wall_s": float(tw[0]), "device_s": float(tw[1]),
                               "rank0_batches": len(infos)})
        barrier()

    tv = torch.tensor([dev_ms, e2e_s, wall_dev, lat_s, strong_s, strong_dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_s_max, wall_dev_max, lat_s_max, strong_s_max, strong_dev_ms_max = (float(x) for x in tv.cpu())
    strong = {"queries": STRONG_QUERIES, "query_tokens": int(soff[-1]), "sharding": "cgx_b200.dist.shard_queries: contiguous, token-balanced, one shard per rank",
              "batch_queries": BATCH_QUERIES, "passes": args.strong_passes, "value": STRONG_QUERIES / strong_s_max, "unit": "query sentences/s",
              "ms_per_pass": 1e3 * strong_s_max, "device_value": STRONG_QUERIES / (strong_dev_ms_max / 1e3), "device_ms_per_pass": strong_dev_ms_max,
              "rank0_shard_queries": int(sq1 - sq0), "rank0_batches_per_pass": strong_runs[0][2],
              "rank0_d2h_bytes_per_pass": strong_bytes[0] // max(1, args.strong_passes),
              "api": "cgx_extract_begin + cgx_result_at per batch (host buffers, every result array copied D2H), wall clock of the whole set, max over ranks"}

    if rank == 0:
        hbm_peak, peak_src = peaks()
        kern = {}
        for name, v in prof.items():
            per = v["ms"] / args.steps
            kern[name] = {"ms_per_step": round(per, 4), "launches_per_step": v["launches"] / args.steps,
                          "alg_bytes_per_step": v["bytes"] / args.steps,
                          "gbs": (v["bytes"] / 1e9) / (v["ms"] / 1e3) if v["ms"] > 0 else None}
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        tn, tvv = top
        ach = (tvv["bytes"] / 1e9) / (tvv["ms"] / 1e3) if tvv["ms"] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")       # per-launch DRAM bytes from the committed ncu --set full capture
        if os.path.exists(tpath):
            tr = json.load(open(tpath)).get(tn, {})
            traffic = tr.get("dram_bytes_per_launch")
            if traffic and tr.get("alg_bytes_of_captured_launch"):     # scale the captured launch to this run's average launch
                traffic = traffic * (tvv["bytes"] / max(1, tvv["launches"])) / tr["alg_bytes_of_captured_launch"]
        roofline = {"bound": "hbm", "kernel": tn, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                    "peak_source": peak_src, "share_of_step": tvv["ms"] / max(1e-9, sum(v["ms"] for v in prof.values())),
                    "avg_launch_ms": tvv["ms"] / max(1, tvv["launches"]), "alg_bytes_per_launch": tvv["bytes"] / max(1, tvv["launches"])}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_baseline_single(lay, args.cpu_budget)
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": "query sentences/s", "cores": 1, "kind": "port", "sample": "failed: %r" % (e,)}
        c3 = None
        if world == 1 and not args.no_c3 and args.workload == "c2":
            ex.close()                                   # the C2 index leaves the GPU before the 260 M-token one is built
            ex = None
            del flush, qt_d, qo_d, t2q_d
            torch.cuda.empty_cache()
            try:
                c3 = run_c3(args, local)
            except Exception as e:  # the second block never costs the first its line
                c3 = {"error": repr(e)}
        gpu_ref = None
        if world == 1 and not args.no_gpu_reference:
            if ex is not None:
                ex.close()
                ex = None
            try:
                gpu_ref = gpu_reference_block(local)
                if cpu is not None and gpu_ref.get("reference", {}).get("extractpair_c"):
                    cpu["reference_leg"] = {"kind": "reference", "what": "createLexiconFast / GappyFast / TwoGapFast of the reference's ExtractPair.c (:515,664,939), one thread, "
                                            "on the rule records its own GPU kernels produced for " + gpu_ref["config"].split(":")[0],
                                            "cores": 1, "records": gpu_ref["reference"].get("extractpair_c_records"), "seconds": gpu_ref["reference"].get("extractpair_c_total_s"),
                                            "product_aggregate_ms_same_files": gpu_ref.get("product", {}).get("ms_aggregate")}
            except Exception as e:  # a reported baseline, never a requirement
                gpu_ref = {"error": repr(e)}
        out = {
            "metric": "query sentences/sec (grammar extraction)", "value": world * Q * args.steps / (dev_ms_max / 1e3), "unit": "query sentences/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_of(args.workload, int(info["n"]), Q, T, world),
            "clocks": clocks,
            "e2e": {"value": world * Q * args.steps / e2e_s_max, "unit": "query sentences/s", "h2d_bytes_per_step": 4 * (2 * T + Q + 1),
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s_max / args.steps,
                    "api": "cgx_extract_begin + cgx_result_at: D2H of step i overlaps the kernels of step i+1, last tail inside the timed region",
                    "single_batch_ms": 1e3 * lat_s_max / args.steps, "rules_read_back": int(e2e_rules)},
            "gpu_launches": int(binfo["launches"]) * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "sa_build": {"gpu_ms": sa_ms, "gpu_ms_first_call": info["sa_build_ms"], "rounds": info["sa_rounds"], "key_bits": info["sa_key_bits"],
                         "alg_bytes": 16.0 * info["n"] * info["sa_rounds"],
                         "gbs": 16.0 * info["n"] * info["sa_rounds"] / 1e9 / (sa_ms / 1e3) if sa_ms else None,
                         "aux_index_ms": info["aux_build_ms"], "cpu_reference_s": cpu.get("sa_build_s") if cpu else None},
            "index_broadcast": {"ms": bcast_ms, "bytes": bcast_bytes} if world > 1 else None,
            "batch": {k: binfo[k] for k in ("G", "D1", "hits1", "D2", "hits2", "samples", "n_ab", "n_1gap", "n_2gap", "rules", "launches", "ms_lookup",
                                           "ms_enum", "ms_join", "ms_extract", "ms_aggregate")},
            "kernels": kern,
            "wall_ms_per_step_incl_flush": 1e3 * wall_dev_max / args.steps,
            "index_bytes": int(info["index_bytes"]),
            "strong": strong,
            "sweep": {"workload": "BASELINE.json configs[4] on the c2 corpus: the first N queries of one fixed 1 M-query set, sharded over the ranks "
                                  "(cgx_b200.dist.shard_queries), 10 k-query batches, end to end (host buffers, all results D2H); sample sizes 300/65/70",
                      "rows": sweep_rows} if sweep_rows else None,
            "c3": c3,
            "gpu_reference": gpu_ref,
        }
        print(json.dumps(out), flush=True)
    if ex is not None:
        ex.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cgx_b200", choices=["cgx_b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of single-thread CPU oracle work for cpu_baseline")
    ap.add_argument("--cpu-queries-per-core", type=int, default=1, help="--impl reference: queries per process and step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the second block (BASELINE.json configs[2]: 10 M sentence pairs, 100 k queries; N = 1 only)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the query-count sweep (BASELINE.json configs[4])")
    ap.add_argument("--only-c3", action="store_true", help="development: run the c3 block alone and print it")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference-binary block (stage timers of oracle/_ref/strmatchcuda_dump; N = 1 only)")
    ap.add_argument("--c3-queries", type=int, default=WORKLOADS["c3"][1], help="queries of the c3 block (default: all 100 k)")
    ap.add_argument("--strong-passes", type=int, default=2, help="timed passes over the fixed strong-scaling query set")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cgx_b200" else max(args.warmup, 0)
    if args.only_c3:                        # development: the second block alone
        print(json.dumps(run_c3(args, 0)), flush=True)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
