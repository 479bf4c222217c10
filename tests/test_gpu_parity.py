"""GPU (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same seeded inputs
(bit-exact for every integer / index result, 1e-5 relative for the two lexical float features), against the
committed reference fixtures, against the reference binary run live next to it, and -- at larger sizes --
through size-independent properties."""
import collections
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cgx_b200 import grammar_compare as gc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


from _parity import assert_full_parity, expand_marker_hits, ms, remap_ids  # noqa: E402,F401


@pytest.fixture(scope="module")
def ex_micro(micro):
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    ex = GrammarExtractor(0)
    ex.build_index(lay)
    res = ex.extract(lay["qry_tok"], lay["qry_off"])
    return ex, res


def test_suffix_array_bit_exact(ex_micro, micro_oracle, golden):
    ex, _ = ex_micro
    sa = ex.suffix_array()
    assert np.array_equal(sa, micro_oracle.sa())
    assert np.array_equal(sa, golden[0]["sa"])          # the reference's own DC3 output


def test_occurrence_lists(ex_micro, micro):
    ex, _ = ex_micro
    _, lay = micro
    s, n = lay["str"].astype(np.int64), lay["n"]
    for m in (1, 2, 3):
        key = np.zeros(n, dtype=np.int64)
        for j in range(m):
            key = key * (int(s.max()) + 1) + s[np.arange(n) + j]
        assert np.array_equal(ex.occurrence_list(m), np.lexsort((np.arange(n), key)))
    assert np.array_equal(ex.frequent_tokens(), np.asarray(golden_freq(micro)))


def golden_freq(micro):
    # frequency descending, ties by ascending id, top 100, stored ascending (SuffixArray.cu:1148-1198)
    _, lay = micro
    s = lay["str"][:lay["n"]]
    ids, cnt = np.unique(s[s >= 2], return_counts=True)
    order = np.lexsort((ids, -cnt))[:100]
    return sorted(ids[order].tolist())


def test_lookup(ex_micro, micro_oracle):
    ex, res = ex_micro
    T = res.T
    assert np.array_equal(ex.debug_fetch("longest", T), np.minimum(micro_oracle.longest(), 5))
    assert np.array_equal(ex.debug_fetch("intervals", T * 10).reshape(T, 5, 2), micro_oracle.intervals(5))


def test_patterns_and_hits(ex_micro, micro_oracle, micro):
    ex, res = ex_micro
    o, oc = micro_oracle, micro_oracle.counts()
    s = micro[1]["str"]
    assert (res.G, res.D1, res.D2, res.info["enu1"], res.info["enu2"]) == (oc.G, oc.D1, oc.D2, oc.enu1, oc.enu2)
    assert ms(res.phrases) == ms(o.blocks())

    def toks(p):
        a, ls, b, le = (int(x) for x in p[:4])
        t = list(s[a:a + ls]) + [-1] + list(s[b:b + le])
        return t + [-2] * (5 - len(t))
    assert np.array_equal(np.array([toks(p) for p in res.pat1], dtype=np.int32), o.onegap_patterns()[:, :5])   # same ids, same order
    assert np.array_equal(res.pat2[:, :2], o.twogap_patterns()[:, :2])
    assert np.array_equal(ex.debug_fetch("hits1", int(res.info["hits1"]) * 3, 3), expand_marker_hits(o))
    assert np.array_equal(ex.debug_fetch("hits2", int(res.info["hits2"]) * 4, 4), o.twogap_hits())
    miss = o.feature_missing()
    pf = ex.debug_fetch("pat1_full", res.D1 * 8, 8)             # device records: + hit_start, hit_count, marker_pair, fs_extra
    assert np.array_equal(pf[:, :4], res.pat1) and np.array_equal(ex.debug_fetch("pat2_full", res.D2 * 4, 4)[:, :2], res.pat2)
    for d in range(res.D1):
        if pf[d, 6] >= 0 and pf[d, 5] > 0:
            assert int(pf[d, 7]) == int(miss[pf[d, 6]])


def test_extraction_records_bit_exact(ex_micro, micro_oracle):
    ex, res = ex_micro
    for k, (name, cnt) in enumerate((("rec_ab", "n_ab"), ("rec_1", "n_1gap"), ("rec_2", "n_2gap"))):
        mine = ex.debug_fetch(name, int(res.info[cnt]) * 7, 7)
        orc = remap_ids(micro_oracle.records(k), res, micro_oracle, k)
        assert ms(mine) == ms(orc), name


def test_rules_bit_exact(ex_micro, micro_oracle):
    """Distinct rules: identical (id, paircount, f, all_suffix_fsample) multisets.  The two float features are
    checked with the 1e-5 tolerance on the grammar lines (next test)."""
    ex, res = ex_micro
    for k in range(3):
        mine = res.rules[k]
        ref = micro_oracle.rules(k)
        rid = remap_ids(ref["id"].reshape(-1, 1).astype(np.int64), res, micro_oracle, k)[:, 0]
        key_m = collections.Counter(zip(mine["id"].tolist(), mine["pc"].tolist(), mine["f"].tolist(), mine["fs"].tolist()))
        key_r = collections.Counter(zip(rid.tolist(), ref["pc"].tolist(), ref["f"].tolist(), ref["fs"].tolist()))
        assert key_m == key_r, k


def test_grammar_files_equal_oracle_and_match_reference(micro, micro_files, golden, tmp_path):
    """Through the drop-in boundary: bin/strmatchcuda (C host + C ABI) on the six input files."""
    out = tmp_path / "mine"
    out.mkdir()
    r = subprocess.run([os.path.join(ROOT, "bin", "strmatchcuda"), "-q", micro_files["f"], micro_files["q"], micro_files["e"], micro_files["a"],
                        micro_files["lex"], str(out)], capture_output=True, text=True)
    assert r.returncode == 0 and "Start Printing Gappy Phrases" in r.stderr, r.stderr[-2000:]
    from _oracle import Oracle
    o = Oracle.from_files(micro_files["f"], micro_files["e"], micro_files["a"], micro_files["lex"])
    o.build_sa()
    o.run_query_file(micro_files["q"])
    orc = tmp_path / "orc"
    orc.mkdir()
    o.write_grammars(str(orc))
    c = gc.compare_dirs(str(out), str(orc), rtol=1e-5, atol=2e-6)
    assert c["files"] == micro[0].n_qry and c["only_a"] == 0 and c["only_b"] == 0 and c["float_mismatch"] == 0, c
    # same rule-GROUP order as the reference writer (PrintResults.c:451-570): the sequence of source sides agrees;
    # the order of target sides inside one source pattern is unspecified in the reference (first-sighting order
    # of atomically appended records) and hash order here
    def groups(path):
        seq = []
        for line in open(path):
            src = line.split(" ||| ")[1]
            if not seq or seq[-1] != src:
                seq.append(src)
        return seq
    assert groups(out / "grammar.0.s") == groups(orc / "grammar.0.s")
    ref = tmp_path / "ref"
    ref.mkdir()
    for q, lines in golden[1].items():
        (ref / ("grammar.%d.s" % q)).write_text("".join(lines))
    c = gc.compare_dirs(str(out), str(ref))
    assert c["frac_equal"] >= 0.97, c                    # north-star bar: >= 0.90


def test_live_reference_binary(tmp_path):
    """Run the unmodified reference binary and the product on a fresh corpus side by side."""
    from _oracle import REF_BIN
    from cgx_b200 import synth
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/strmatchcuda not present")
    c = synth.generate(3000, 24, v_src=600, v_tgt=600, n_phrases=1200, seed=99, qry_seed=77)
    p = synth.write_text(c, str(tmp_path), "c")
    args = [p["f"], p["q"], p["e"], p["a"], p["lex"]]
    (tmp_path / "ref").mkdir()
    (tmp_path / "mine").mkdir()
    for attempt in range(3):          # the reference reads memory it never wrote: a run that dies is repeated
        r1 = subprocess.run([REF_BIN] + args + [str(tmp_path / "ref")], capture_output=True, text=True, cwd=str(tmp_path))
        if "Start Printing Gappy Phrases" in r1.stderr:
            break
    assert "Start Printing Gappy Phrases" in r1.stderr, r1.stderr[-1500:]
    r2 = subprocess.run([os.path.join(ROOT, "bin", "strmatchcuda"), "-q"] + args + [str(tmp_path / "mine")], capture_output=True, text=True)
    assert r2.returncode == 0, r2.stderr[-1500:]
    cmp_ = gc.compare_dirs(str(tmp_path / "mine"), str(tmp_path / "ref"))
    assert cmp_["files"] == 24 and cmp_["frac_equal"] >= 0.95, cmp_


def test_edge_cases(micro):
    """Empty batch, empty queries, all-OOV queries, a one-token query, OOV inside a query."""
    from _oracle import Oracle
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    ex = GrammarExtractor(0)
    ex.build_index(lay)
    o = Oracle.from_layout(lay)
    o.build_sa()
    q = lay["qry_tok"]
    cases = [
        (np.zeros(0, np.int32), np.array([0], np.int32)),
        (np.zeros(0, np.int32), np.array([0, 0, 0], np.int32)),
        (np.array([-1, -1, -1], np.int32), np.array([0, 3], np.int32)),
        (q[:1].copy(), np.array([0, 1], np.int32)),
        (np.concatenate([q[:6], [-1], q[6:14]]).astype(np.int32), np.array([0, 0, 15], np.int32)),
    ]
    for tok, off in cases:
        res = ex.extract(tok, off)
        if len(off) - 1 == 0:
            assert res.G == 0 and sum(len(r) for r in res.rules) == 0
            continue
        oc = o.run(tok, off)
        assert (res.G, res.D1, res.D2) == (oc.G, oc.D1, oc.D2)
        for k in range(3):
            assert len(res.rules[k]) == len(o.rules(k))


def test_queries_longer_than_128_tokens(micro):
    """SURVEY.md 8f, lifted limit: the reference's lookup kernel gives a query 128 threads, one per token (SuffixArray.cu:402-419),
    so it never sees token 129 of a sentence.  Neither the product nor the oracle has that bound: a 300-token query (the batch's
    queries concatenated, OOV tokens included) goes through every stage bit-exact, and its tail -- beyond token 128 -- yields rules."""
    from _oracle import Oracle
    from _parity import assert_full_parity
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    q = lay["qry_tok"]
    long_q = np.tile(np.concatenate([q, q[::-1]]), 3)[:300].astype(np.int32)
    assert len(long_q) == 300
    tok = np.concatenate([long_q, q[:9]]).astype(np.int32)
    off = np.array([0, 300, 309], np.int32)
    ex = GrammarExtractor(0)
    try:
        ex.build_index(lay)
        res = ex.extract(tok, off)
        o = Oracle.from_layout(lay)
        o.build_sa()
        o.run(tok, off)
        assert_full_parity(ex, res, lay, o)
        tail = ex.extract(long_q[128:].copy(), np.array([0, 300 - 128], np.int32))
        assert tail.G > 0 and len(tail.rules[1]) > 0
        # phrases of the tail are phrases of the long query: nothing past token 128 was dropped
        up_len = {(int(p[0]), int(p[2])) for p in res.phrases}
        assert all((int(p[0]), int(p[2])) in up_len for p in tail.phrases)
    finally:
        ex.close()


def test_batching_is_transparent(micro):
    """Per-query output does not depend on batch composition: one batch == two batches, query by query."""
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    ex = GrammarExtractor(0)
    ex.build_index(lay)
    off = lay["qry_off"]
    full = ex.extract(lay["qry_tok"], off)
    lines_full = [sorted(full.grammar_lines(q, lay)) for q in range(full.Q)]
    h = full.Q // 2
    a = ex.extract(lay["qry_tok"][:off[h]], off[:h + 1])
    la = [sorted(a.grammar_lines(q, lay)) for q in range(a.Q)]
    b = ex.extract(lay["qry_tok"][off[h]:], off[h:] - off[h])
    lb = [sorted(b.grammar_lines(q, lay)) for q in range(b.Q)]
    assert la + lb == lines_full


def test_pipelined_batches_equal_synchronous(micro):
    """cgx_extract_begin / cgx_result_at: five batches through the three-set pipeline; every batch read back at age 1 (or
    2, or 0 for the last) equals the same batch through the synchronous cgx_extract, array by array."""
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    ex = GrammarExtractor(0)
    ex.build_index(lay)
    off, tok = lay["qry_off"], lay["qry_tok"]
    Q = len(off) - 1
    cuts = [0, Q // 7, Q // 3, Q // 3, Q // 2, Q]          # ragged batches, one of them empty
    parts = [(tok[off[a]:off[b]], off[a:b + 1] - off[a]) for a, b in zip(cuts[:-1], cuts[1:])]
    want = [ex.extract(t, o) for t, o in parts]

    def same(x, y):
        assert (x.Q, x.T, x.G, x.D1, x.D2) == (y.Q, y.T, y.G, y.D1, y.D2)
        for name in ("phrase_id", "phrases", "pat1", "pat2", "q1_off", "q1_ids", "q2_off", "q2_ids"):
            assert np.array_equal(getattr(x, name), getattr(y, name)), name
        for k in range(3):
            assert x.rules[k].tobytes() == y.rules[k].tobytes(), k
            assert np.array_equal(x.updown[k], y.updown[k]), k

    got = [None] * len(parts)
    for i, (t, o) in enumerate(parts):
        ex.extract_begin(t, o)
        if i >= 2:                                           # the batch two begins ago is still there
            same(ex.result_at(2, *parts[i - 2]), want[i - 2])
        if i >= 1:
            got[i - 1] = ex.result_at(1, *parts[i - 1])
    got[-1] = ex.result_at(0, *parts[-1])
    for g, w in zip(got, want):
        same(g, w)
    with pytest.raises(RuntimeError):
        ex.result_at(3)


def test_cli_pipelined_batches_write_the_same_grammars(micro, micro_files, tmp_path):
    """bin/strmatchcuda -b: many small batches through the begin / writer-thread pipeline == one batch, file by file."""
    outs = []
    for name, extra in (("one", []), ("many", ["-b", "7", "-w", "2"])):
        out = tmp_path / name
        out.mkdir()
        r = subprocess.run([os.path.join(ROOT, "bin", "strmatchcuda"), "-q"] + extra + [micro_files["f"], micro_files["q"], micro_files["e"], micro_files["a"],
                            micro_files["lex"], str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(out)
    c = gc.compare_dirs(str(outs[0]), str(outs[1]), rtol=0, atol=0)
    assert c["files"] == micro[0].n_qry and c["only_a"] == 0 and c["only_b"] == 0 and c["float_mismatch"] == 0, c


def test_cli_server_mode_and_gzip(micro, micro_files, tmp_path):
    """strmatchcuda -S keeps the index resident and serves "<query file> <output dir>" requests from stdin (the
    reference reloads the corpus and rebuilds its suffix array for every query file); -z writes grammar.<qid>.s.gz.  Every
    request's grammar files equal those of a one-shot run on the same query file."""
    import gzip
    cli = os.path.join(ROOT, "bin", "strmatchcuda")
    q_lines = open(micro_files["q"]).read().splitlines()
    half = tmp_path / "half.q"                                   # a second, different query file: the last half, reversed
    half.write_text("".join(x + "\n" for x in q_lines[len(q_lines) // 2:][::-1]))
    args = [micro_files["e"], micro_files["a"], micro_files["lex"]]
    ref_full, ref_half = tmp_path / "ref_full", tmp_path / "ref_half"
    for out, qf in ((ref_full, micro_files["q"]), (ref_half, str(half))):
        out.mkdir()
        r = subprocess.run([cli, "-q", micro_files["f"], qf] + args + [str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    s0, s1, s2 = tmp_path / "s0", tmp_path / "s1", tmp_path / "s2"
    for d in (s0, s1, s2):
        d.mkdir()
    req = "%s %s\n\n%s %s\n%s\nquit\n" % (half, s1, micro_files["q"], s2, "lonely_field")
    r = subprocess.run([cli, "-q", "-S", "-z", "6", micro_files["f"], micro_files["q"]] + args + [str(s0)], input=req, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    answers = r.stdout.strip().splitlines()
    assert len(answers) == 3 and answers[0].startswith("done %s %d queries" % (half, len(q_lines) - len(q_lines) // 2)), answers
    assert answers[1].startswith("done %s %d queries" % (micro_files["q"], len(q_lines))) and answers[2].startswith("error lonely_field"), answers
    assert r.stderr.count("SA Construction") == 1 and r.stderr.count("Start Printing Gappy Phrases") == 3      # one index, three query files
    for served, want in ((s0, ref_full), (s1, ref_half), (s2, ref_full)):
        plain = tmp_path / (served.name + "_plain")
        plain.mkdir()
        names = sorted(os.listdir(served))
        assert names and all(n.endswith(".s.gz") for n in names)
        for n in names:
            (plain / n[:-3]).write_bytes(gzip.open(served / n).read())
        c = gc.compare_dirs(str(plain), str(want), rtol=0, atol=0)
        assert c["files"] == len(os.listdir(want)) and c["only_a"] == 0 and c["only_b"] == 0 and c["float_mismatch"] == 0 and c["missing_files"] == 0, c


def test_oversized_batches_are_split(micro, micro_files, tmp_path, monkeypatch):
    """A batch whose hit lists exceed the index width of the result tables is refused with CGX_E_BATCH_TOO_LARGE (the
    limit is lowered here through CGX_HIT_LIMIT); the callers -- extract_stream and bin/strmatchcuda -- cut it in two,
    the pipeline state survives the refusal, and the grammars equal those of the unsplit run."""
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    ex = GrammarExtractor(0)
    ex.build_index(lay)
    off, tok = lay["qry_off"], lay["qry_tok"]
    full = ex.extract(tok, off)
    want = [sorted(full.grammar_lines(q, lay)) for q in range(full.Q)]
    single = 0                                               # the largest single query must still fit
    for q in range(full.Q):
        i = ex.extract(tok[off[q]:off[q + 1]], off[q:q + 2] - off[q], fetch=False)
        single = max(single, int(i["hits1"]), int(i["hits2"]))
    total = max(int(full.info["hits1"]), int(full.info["hits2"]))
    limit = max(single + 1, total // 3)
    assert limit < total
    monkeypatch.setenv("CGX_HIT_LIMIT", str(limit))
    assert ex.L.cgx_extract(ex.h, tok.ctypes.data_as(C.POINTER(C.c_int32)), np.ascontiguousarray(off, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int32)), full.Q) == 3
    got = {}

    def on_batch(q0, q1, r):
        from cgx_b200.extractor import BatchResult
        from cgx_b200._lib import BatchInfo
        br = BatchResult(r, BatchInfo(), tok[off[q0]:off[q1]], off[q0:q1 + 1] - off[q0])
        for q in range(q0, q1):
            got[q] = sorted(br.grammar_lines(q - q0, lay))

    infos = ex.extract_stream(tok, off, batch_queries=full.Q, on_batch=on_batch)
    assert len(infos) >= 3 and [got[q] for q in range(full.Q)] == want
    # cgx_batch_advice: the refusals of that stream are remembered -- the same queries again, asked for in one batch, are cut to a
    # size expected to fit before anything runs, so the second stream pays for fewer refused scans (none, when the queries are alike)
    first_refusals = ex.stream_refusals
    assert first_refusals >= 1 and ex.L.cgx_batch_advice(ex.h, full.Q) < full.Q
    got.clear()
    infos2 = ex.extract_stream(tok, off, batch_queries=full.Q, on_batch=on_batch)
    assert ex.stream_refusals <= max(0, first_refusals - 1) and len(infos2) >= 2 and [got[q] for q in range(full.Q)] == want
    outs = []
    for name in ("plain", "split"):
        out = tmp_path / name
        out.mkdir()
        env = dict(os.environ)
        if name == "plain":
            env.pop("CGX_HIT_LIMIT", None)
        r = subprocess.run([os.path.join(ROOT, "bin", "strmatchcuda"), micro_files["f"], micro_files["q"], micro_files["e"], micro_files["a"],
                            micro_files["lex"], str(out)], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        assert ("splitting" in r.stderr) == (name == "split")
        outs.append(out)
    c = gc.compare_dirs(str(outs[0]), str(outs[1]), rtol=0, atol=0)
    assert c["files"] == micro[0].n_qry and c["only_a"] == 0 and c["only_b"] == 0 and c["float_mismatch"] == 0, c


def test_persisted_index_round_trip(micro, micro_files, tmp_path):
    """cgx_index_save / cgx_index_load: a context that loads the file answers a batch exactly like the one that built the
    index (suffix array and every result array bit-identical); garbage and truncated files are refused; bin/strmatchcuda -i
    builds + saves on the first run, loads on the second, and writes the same grammars."""
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    a = GrammarExtractor(0)
    a.build_index(lay)
    path = tmp_path / "micro.cgxidx"
    a.save_index(path)
    b = GrammarExtractor(0)
    info = b.load_index(path)
    assert (info["n"], info["m"]) == (int(lay["n"]), int(lay["m"]))
    assert np.array_equal(a.suffix_array(), b.suffix_array())
    ra, rb = a.extract(lay["qry_tok"], lay["qry_off"]), b.extract(lay["qry_tok"], lay["qry_off"])
    assert (ra.G, ra.D1, ra.D2) == (rb.G, rb.D1, rb.D2)
    for name in ("phrase_id", "phrases", "pat1", "pat2", "q1_ids", "q2_ids"):
        assert np.array_equal(getattr(ra, name), getattr(rb, name)), name
    for k in range(3):
        assert ra.rules[k].tobytes() == rb.rules[k].tobytes(), k
    bad = tmp_path / "bad.cgxidx"
    bad.write_bytes(b"not an index")
    c = GrammarExtractor(0)
    with pytest.raises(RuntimeError):
        c.load_index(bad)
    trunc = tmp_path / "trunc.cgxidx"
    trunc.write_bytes(path.read_bytes()[: path.stat().st_size // 2])
    with pytest.raises(RuntimeError):
        c.load_index(trunc)
    outs = []
    idx = tmp_path / "cli.cgxidx"
    for name in ("first", "second"):
        out = tmp_path / name
        out.mkdir()
        r = subprocess.run([os.path.join(ROOT, "bin", "strmatchcuda"), "-i", str(idx), micro_files["f"], micro_files["q"], micro_files["e"], micro_files["a"],
                            micro_files["lex"], str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        assert ("index saved to" in r.stderr) == (name == "first") and ("index loaded from" in r.stderr) == (name == "second"), r.stderr[-1500:]
        outs.append(out)
    cmp = gc.compare_dirs(str(outs[0]), str(outs[1]), rtol=0, atol=0)
    assert cmp["files"] == micro[0].n_qry and cmp["only_a"] == 0 and cmp["only_b"] == 0 and cmp["float_mismatch"] == 0, cmp
    # an index file of ANOTHER corpus with the same token counts (one source word replaced by its neighbour) is refused: the
    # file records checksums of the token arrays it was built from (cgx_index_matches)
    words = open(micro_files["f"]).read().split("\n")
    first = words[0].split()
    k = next(i for i in range(len(first) - 1) if first[i] != first[i + 1])
    first[k + 1] = first[k]
    other = tmp_path / "other.f"
    other.write_text("\n".join([" ".join(first)] + words[1:]))
    out = tmp_path / "third"
    out.mkdir()
    r = subprocess.run([os.path.join(ROOT, "bin", "strmatchcuda"), "-i", str(idx), str(other), micro_files["q"], micro_files["e"], micro_files["a"],
                        micro_files["lex"], str(out)], capture_output=True, text=True)
    assert r.returncode != 0 and "was built from another corpus" in r.stderr, r.stderr[-1500:]
    assert not os.path.exists(str(idx) + ".tmp")                       # saves go through <file>.tmp + rename


def test_medium_scale_properties():
    """A corpus the oracle would need minutes for: size-independent properties only.
    SA is a permutation whose adjacent suffixes are in order; occurrence lists are position-sorted inside every
    bucket; sharded extraction equals unsharded extraction."""
    from cgx_b200 import synth
    from cgx_b200.extractor import GrammarExtractor
    c = synth.generate(120000, 300, v_src=20000, v_tgt=20000, seed=5, qry_seed=6)
    lay = synth.text_layout(c)
    ex = GrammarExtractor(0)
    info = ex.build_index(lay)
    n, s = lay["n"], lay["str"]
    sa = ex.suffix_array()
    assert np.array_equal(np.sort(sa), np.arange(n, dtype=np.int32))
    rng = np.random.default_rng(0)
    for k in rng.integers(0, n - 1, size=3000):
        a, b = int(sa[k]), int(sa[k + 1])
        j = 0
        while s[a + j] == s[b + j]:
            j += 1
        assert s[a + j] < s[b + j]
    inv = ex.occurrence_list(2)
    key = s[inv].astype(np.int64) * (int(s.max()) + 1) + s[inv + 1]
    assert np.all(np.diff(key) >= 0)
    same = np.diff(key) == 0
    assert np.all(np.diff(inv.astype(np.int64))[same] > 0)
    assert info["sa_rounds"] >= 3
    off = lay["qry_off"]
    full = ex.extract(lay["qry_tok"], off)
    tot_full = sum(len(full.grammar_lines(q, lay)) for q in (0, 149, 150, 299))
    h = 150
    a = ex.extract(lay["qry_tok"][:off[h]], off[:h + 1])
    b = ex.extract(lay["qry_tok"][off[h]:], off[h:] - off[h])
    assert sorted(a.grammar_lines(149, lay)) == sorted(full.grammar_lines(149, lay))
    assert sorted(b.grammar_lines(0, lay)) == sorted(full.grammar_lines(150, lay))
    assert tot_full > 0


@pytest.mark.parametrize("n,bits,with_vals", [(1, 8, 0), (4097, 17, 1), (300_000, 52, 0), (1_000_003, 64, 1), (2_500_000, 23, 1)])
def test_radix_sort_matches_stable_reference(n, bits, with_vals):
    """The hand-written onesweep radix sort (every stage of the path sorts with it) against numpy's stable sort:
    keys bit-exact and payloads equal to the stable permutation, ragged sizes and partial key widths included."""
    import ctypes as C
    import torch
    from cgx_b200 import _lib
    L = _lib.load()
    h = C.c_void_p()
    assert L.cgx_create(0, C.byref(h)) == 0
    try:
        rng = np.random.default_rng(n)
        src = rng.integers(0, 1 << 63, size=n, dtype=np.uint64)
        if n > 1000:
            src[: n // 3] &= np.uint64(0xFFFF)            # heavy duplicates: stability matters
        if bits < 64:
            src &= np.uint64((1 << bits) - 1)
        keys = torch.from_numpy(src.view(np.int64).copy()).cuda()
        vals = torch.arange(n, dtype=torch.int32).cuda() if with_vals else None
        ms, passes = C.c_float(), C.c_int()
        rc = L.cgx_debug_sort_u64(h, keys.data_ptr(), vals.data_ptr() if with_vals else None, n, 0, bits, C.byref(ms), C.byref(passes))
        assert rc == 0, L.cgx_last_error(h)
        perm = np.argsort(src, kind="stable")
        assert np.array_equal(keys.cpu().numpy().view(np.uint64), src[perm])
        if with_vals:
            assert np.array_equal(vals.cpu().numpy().astype(np.int64), perm)
    finally:
        L.cgx_destroy(h)


@pytest.mark.parametrize("mode,ordered", [("phrase", "1"), ("position", "1"), ("position", "0"), ("phrase", "0"), ("position", "direct")])
def test_join_variants_agree_with_oracle(mode, ordered, micro, micro_oracle, monkeypatch):
    """Every join variant -- walk of the first phrases' occurrence lists / one streamed pass over the corpus, hits emitted in
    position order through the tile look-back (sort on the pattern bits only) / appended unordered (full sort) -- must give the
    oracle's hit list, two-gap hits and missing counts bit for bit."""
    from cgx_b200.extractor import GrammarExtractor
    monkeypatch.setenv("CGX_JOIN_MODE", mode)
    if ordered == "direct":                                 # a 40-hit stage: most warps flush it several times (many chunks per warp)
        monkeypatch.setenv("CGX_JOIN_STAGE_CAP", "40")
        ordered = "1"
    monkeypatch.setenv("CGX_JOIN_ORDERED", ordered)
    _, lay = micro
    ex = GrammarExtractor(0)
    try:
        ex.build_index(lay)
        res = ex.extract(lay["qry_tok"], lay["qry_off"])
        o = micro_oracle
        assert np.array_equal(ex.debug_fetch("hits1", int(res.info["hits1"]) * 3, 3), expand_marker_hits(o))
        assert np.array_equal(ex.debug_fetch("hits2", int(res.info["hits2"]) * 4, 4), o.twogap_hits())
        miss = o.feature_missing()
        pf = ex.debug_fetch("pat1_full", res.D1 * 8, 8)
        for d in range(res.D1):
            if pf[d, 6] >= 0 and pf[d, 5] > 0:
                assert int(pf[d, 7]) == int(miss[pf[d, 6]])
        for k in range(3):
            assert len(res.rules[k]) == len(o.rules(k))
    finally:
        ex.close()


LONG = dict(n_sent=400, n_qry=8, v_src=300, v_tgt=300, n_phrases=600, mean_len=150.0, sd_len=90.0, max_len=600, qry_mean_len=12.0, seed=5, qry_seed=6)


def test_forced_wide_fields_equal_narrow(micro, monkeypatch):
    """SURVEY.md 8f, lifted limit: the 16-bit alignment layout (align_fields.cuh, cgx_index_build_wide) on a corpus that fits
    the reference's 8 bits -- CGX_FORCE_WIDE=1 widens the input of cgx_index_build -- gives byte-identical results."""
    from cgx_b200.extractor import GrammarExtractor
    _, lay = micro
    a = GrammarExtractor(0)
    a.build_index(lay)
    monkeypatch.setenv("CGX_FORCE_WIDE", "1")
    b = GrammarExtractor(0)
    b.build_index(lay)
    monkeypatch.delenv("CGX_FORCE_WIDE")
    try:
        ra, rb = a.extract(lay["qry_tok"], lay["qry_off"]), b.extract(lay["qry_tok"], lay["qry_off"])
        assert ra.info["hits1"] == rb.info["hits1"] and ra.info["hits2"] == rb.info["hits2"] and ra.info["hits1"] > 0
        for k in range(3):
            assert len(ra.rules[k]) > 0 and ra.rules[k].tobytes() == rb.rules[k].tobytes(), k
        for name, cnt, cols in (("rec_ab", "n_ab", 7), ("rec_1", "n_1gap", 7), ("rec_2", "n_2gap", 7), ("hits1", "hits1", 3), ("hits2", "hits2", 4)):
            assert np.array_equal(a.debug_fetch(name, int(ra.info[cnt]) * cols, cols), b.debug_fetch(name, int(rb.info[cnt]) * cols, cols)), name
    finally:
        a.close()
        b.close()


def test_sentences_of_255_tokens_and_more(tmp_path):
    """SURVEY.md 8f, lifted limit.  The reference exits on a sentence of 255 tokens ("Not possible, too long sentence",
    ExtractPair.cu:2683; 8-bit position counters, Start.cu:269).  Here the C loaders switch to the 16-bit alignment layout, the
    index is built with cgx_index_build_wide, and every stage equals the wide build of the oracle (the same algorithm on 16-bit
    fields): stage by stage through the C ABI, and grammar file by grammar file through the drop-in CLI -- also from a persisted
    index (-i), whose file records the layout."""
    from _oracle import Oracle
    from _parity import assert_full_parity
    from cgx_b200 import synth
    from cgx_b200.extractor import GrammarExtractor
    from cgx_b200.host import HostCorpus
    files = synth.write_text(synth.generate(**LONG), str(tmp_path), "corpus")
    assert max(len(l.split()) for l in open(files["f"])) >= 300
    hc = HostCorpus(files["f"], files["q"], files["e"], files["a"], files["lex"])
    lay = hc.layout()
    assert lay["wide"]
    o = Oracle.from_layout(lay)
    assert o.wide
    o.build_sa()
    o.run(lay["qry_tok"], lay["qry_off"])
    ex = GrammarExtractor(0)
    try:
        ex.build_index(lay)
        res = ex.extract(lay["qry_tok"], lay["qry_off"])
        assert res.info["hits1"] > 10_000 and len(res.rules[2]) > 1000
        assert_full_parity(ex, res, lay, o)
    finally:
        ex.close()
    # rules that lie beyond token 255 of their sentence exist (the part the 8-bit layout could not address)
    r1 = o.records(1)
    sent_start = np.concatenate([[0], np.cumsum([len(l.split()) + 1 for l in open(files["e"])])])
    rel = r1[:, 1] - sent_start[np.searchsorted(sent_start, r1[:, 1], side="right") - 1]
    assert (rel >= 255).sum() > 100
    # the drop-in CLI: text files in, grammar files out; then the same from the persisted index
    of = Oracle.from_files(files["f"], files["e"], files["a"], files["lex"], wide=True)
    of.build_sa()
    of.run_query_file(files["q"])
    orc = tmp_path / "orc"
    orc.mkdir()
    of.write_grammars(str(orc))
    cli = os.path.join(ROOT, "bin", "strmatchcuda")
    idx = str(tmp_path / "corpus.idx")
    for name in ("first", "second"):
        out = tmp_path / name
        out.mkdir()
        r = subprocess.run([cli, "-q", "-i", idx, files["f"], files["q"], files["e"], files["a"], files["lex"], str(out)], capture_output=True, text=True)
        assert r.returncode == 0 and "Start Printing Gappy Phrases" in r.stderr, r.stderr[-2000:]
        assert ("index loaded from" in r.stderr) == (name == "second")
        c = gc.compare_dirs(str(out), str(orc), rtol=1e-5, atol=2e-6)
        assert c["files"] == LONG["n_qry"] and c["only_a"] == 0 and c["only_b"] == 0 and c["float_mismatch"] == 0, c
