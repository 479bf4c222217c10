"""CPU: the C-ABI library loads and exports every symbol include/cgx_b200.h declares (no compute calls without
a GPU), the product fails loudly without a device (no fallback), and the C host loaders reproduce the reference
layouts (checked against the independent numpy construction in cgx_b200/synth.py and the oracle's own loader)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(built):
    from cgx_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cgx_b200.h")).read()
    declared = set(re.findall(r"\b(cgx_[a-z0-9_]+)\s*\(", hdr))
    lib = _lib.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.cgx_version() >= 100
    assert set(_lib.EXPORTED_SYMBOLS) <= declared


def test_no_cpu_fallback(built):
    """Without a CUDA device every entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cgx_b200.extractor import GrammarExtractor
    with pytest.raises(RuntimeError):
        GrammarExtractor(0)


def test_product_never_touches_the_oracle():
    """cgx_b200/ (kernels, ABI, host) must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "cgx_b200")
    for dirpath, _, files in os.walk(pkg):
        if "lib" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    if re.search(r"#\s*include.*oracle|import\s+_oracle|from\s+_oracle|cgx_oracle|-loracle|oracle/_build|oracle/_ref", line):
                        if "oracle/cgx_oracle.c" in line and line.strip().startswith(("//", "*", "#", "/*")):
                            continue   # a comment pointing at the restatement
                        pytest.fail("%s references the oracle: %s" % (os.path.join(dirpath, f), line.strip()))
    out = subprocess.run(["ldd", os.path.join(pkg, "lib", "libcgx_host.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_host_loaders_match_reference_layout(micro, micro_files, built):
    from cgx_b200.host import HostCorpus
    _, lay = micro
    hc = HostCorpus(micro_files["f"], micro_files["q"], micro_files["e"], micro_files["a"], micro_files["lex"])
    h = hc.layout()
    assert h["n"] == lay["n"] and h["m"] == lay["m"]
    for k in ("str", "tgt", "P", "L_tar", "R_tar", "qry_tok", "qry_off"):
        assert np.array_equal(h[k], lay[k]), k
    # RLP[n-1] is never written by the reference (ExtractPair.cu:2721 stops at toklen-1)
    assert np.array_equal(h["RLP"][:-1], lay["RLP"][:-1])
    # lexical table: same multiset of (f, e, v1, v2)
    a = sorted(zip(h["lex_f"].tolist(), h["lex_e"].tolist(), h["lex_v1"].tolist(), h["lex_v2"].tolist()))
    b = sorted(zip(lay["lex_f"].tolist(), lay["lex_e"].tolist(), lay["lex_v1"].tolist(), lay["lex_v2"].tolist()))
    assert a == b
    assert hc.src_name(2) == "s%d" % lay["src_names"][0]
    assert h["src_last"] == lay["src_last"]


def test_loader_edge_cases(tmp_path, built):
    """Empty lines, OOV query words, ragged alignment lines, a token that starts with a tab (Start.cu:279)."""
    from cgx_b200.host import HostCorpus
    (tmp_path / "f").write_text("a b c\n\nd a\n")
    (tmp_path / "e").write_text("x y\n\nz\n")
    (tmp_path / "a").write_text("0-0 2-1\n\n0-0 1-0\n")
    (tmp_path / "q").write_text("a zzz b\n\nc \tq d\n")
    (tmp_path / "lex").write_text("a x 0.5 0.25\nNULL x 0.1 0.2\nb NULL 0.3 0.4\nunk x 1 1\nNULL NULL 1 1\n")
    hc = HostCorpus(*(str(tmp_path / k) for k in ("f", "q", "e", "a", "lex")))
    h = hc.layout()
    assert h["str"].tolist() == [2, 3, 4, 1, 1, 5, 2, 1, 1, 6, 0, 0, 0]
    assert h["tgt"].tolist() == [2, 3, 1, 1, 4, 1, 1, 5, 0, 0, 0]
    assert h["qry_tok"].tolist() == [2, -1, 3, 4] and h["qry_off"].tolist() == [0, 3, 3, 4]
    assert h["P"].tolist() == [0, 1, 2, 0, 0, 0, 1, 0, 0, 0]
    L = (h["RLP"] >> 24) & 255
    assert L[:3].tolist() == [0, 255, 1] and h["RLP"][3] == 3 and h["RLP"][4] == 4
    assert h["L_tar"].tolist()[:2] == [0, 2] and h["L_tar"][4] == 0 and h["R_tar"][4] == 1
    assert sorted(zip(h["lex_f"].tolist(), h["lex_e"].tolist())) == [(-1, -1), (-1, 2), (2, 2), (3, -1)]


def _strtok_walk(line, delims=" "):
    """The reference's walk over a line (Start.cu:270-310): one trailing newline stripped, tokens = maximal runs of characters outside
    `delims`, the walk ends at the first token that begins with white space."""
    if line.endswith("\n"):
        line = line[:-1]
    out = []
    for t in re.split("[" + re.escape(delims) + "]+", line):
        if t == "":
            continue
        if t[0] in " \t\r\v\f\n":
            break
        out.append(t)
    return out


def _atoi(t):
    m = re.match(r"[ \t\r\v\f\n]*\+?([0-9]*)", t)
    return int(m.group(1)) if m.group(1) else 0


def test_loaders_on_ragged_text_equal_a_python_restatement(tmp_path, built):
    """Leading / doubled / trailing blanks, tabs and carriage returns inside tokens, a token that begins with a tab (the rest of the line
    is dropped), empty lines, files without a final newline, alignment lines with stray blanks, lexical fields split over lines:
    the C loaders (two threads at a time, cgxh_load_files) against a line-by-line Python restatement of Start.cu:50-132, :240-380,
    ExtractPair.cu:2463-2519 and :2639-2739."""
    import random
    from cgx_b200.host import HostCorpus
    rnd = random.Random(11)
    for final_newline in (True, False):
        src, tgt, al = [], [], []
        for k in range(300):
            ns, nt = rnd.randint(1, 30), rnd.randint(1, 30)
            s, t = ["s%d" % rnd.randint(0, 40) for _ in range(ns)], ["t%d" % rnd.randint(0, 40) for _ in range(nt)]
            mode = k % 10
            if mode == 3 and ns > 3:
                s[2] = "a\tb"
            if mode == 4 and ns > 3:
                s[rnd.randrange(1, ns)] = "\tzz"
            sl, tl = " ".join(s), " ".join(t)
            if mode == 1:
                sl = "  " + sl.replace(" ", "   ", 2) + "  "
            if mode == 2:
                sl, tl = sl + "\r", tl + "\r"
            if mode == 7:
                tl = ""
            ns_eff, nt_eff = len(_strtok_walk(sl)), len(_strtok_walk(tl))
            links = sorted({(rnd.randrange(ns_eff), rnd.randrange(nt_eff)) for _ in range(rnd.randint(0, 40))}) if ns_eff and nt_eff else []
            a = " ".join("%d-%d" % l for l in links)
            if mode == 5:
                a = a.replace(" ", "  ") + " "
            if mode == 6:
                a = " " + a
            if mode == 8:
                a = ""
            if mode == 9 and links:
                a = a.replace("-", " - ", 1)
            src.append(sl); tgt.append(tl); al.append(a)
        lexl = []
        for i in range(2000):
            f = rnd.choice(["NULL", "s%d" % rnd.randint(0, 60), "unk%d" % i])
            e = rnd.choice(["NULL", "t%d" % rnd.randint(0, 60), "unk"])
            lexl.append(f + rnd.choice([" ", "  ", "\t", " \n"]) + e + " %.6g %g" % (rnd.random(), rnd.random() * 1e-3))
        qry = [" ".join(rnd.choice(["s%d" % rnd.randint(0, 50), "oov"]) for _ in range(rnd.randint(0, 40))) for _ in range(40)]
        qry[3], qry[5], qry[7] = "  " + qry[3] + "  ", qry[5] + "\r", "s1 \ts2 s3"
        end = "\n" if final_newline else ""
        for name, lines in (("f", src), ("e", tgt), ("a", al), ("lex", lexl), ("q", qry)):
            (tmp_path / name).write_text("\n".join(lines) + end)
        h = HostCorpus(*(str(tmp_path / k) for k in ("f", "q", "e", "a", "lex"))).layout()

        def side(lines):
            ids, tok, sent, pos = {}, [], [0], []
            for line in lines:
                for j, w in enumerate(_strtok_walk(line)):
                    tok.append(ids.setdefault(w, len(ids) + 2))
                    pos.append(j & 255)
                tok.append(1); pos.append(0); sent.append(len(tok))
            last = len(ids) + 2                          # one past the id added last
            return ids, tok + [1, last, 0, 0, 0], sent, pos + [0, 0]
        sid, stok, ssent, spos = side(src)
        tid, ttok, tsent, _ = side(tgt)
        assert h["str"].tolist() == stok and h["tgt"].tolist() == ttok
        assert h["src_sentenceind"].tolist() == ssent and h["tgt_sentenceind"].tolist() == tsent and h["P"].tolist() == spos
        qt, qo = [], [0]
        for line in qry:
            qt += [sid.get(w, -1) for w in _strtok_walk(line)]
            qo.append(len(qt))
        assert h["qry_tok"].tolist() == qt and h["qry_off"].tolist() == qo
        n, m = h["n"], h["m"]
        Ls, Rs, Lt, Rt = [255] * n, [255] * n, [255] * m, [255] * m
        for k, line in enumerate(al):
            nums = [_atoi(t) for t in _strtok_walk(line, " -")]
            for s_no, t_no in zip(nums[0::2], nums[1::2]):
                si, ti = ssent[k] + s_no, tsent[k] + t_no
                Ls[si], Rs[si] = (t_no, t_no) if Ls[si] == 255 else (min(Ls[si], t_no), max(Rs[si], t_no))
                Lt[ti], Rt[ti] = (s_no, s_no) if Lt[ti] == 255 else (min(Lt[ti], s_no), max(Rt[ti], s_no))
        assert h["L_tar"].tolist() == Lt and h["R_tar"].tolist() == Rt
        eos = set(x - 1 for x in ssent[1:])
        got = h["RLP"]
        for i in range(n - 1):
            if i not in eos:
                assert (int(got[i]) >> 24, (int(got[i]) >> 16) & 255, (int(got[i]) >> 8) & 255) == (Ls[i], Rs[i], spos[i]), i
        fields = " ".join(lexl).split()
        want = []
        for f, e, x, y in zip(fields[0::4], fields[1::4], fields[2::4], fields[3::4]):
            fi, ei = sid.get(f, -1), tid.get(e, -1)
            if (fi < 0 and f != "NULL") or (ei < 0 and e != "NULL"):
                continue
            want.append((fi, ei, np.float32(x), np.float32(y)))
        assert list(zip(h["lex_f"].tolist(), h["lex_e"].tolist(), h["lex_v1"], h["lex_v2"])) == want


def test_writer_prints_the_oracles_results_exactly_like_the_oracle(micro, micro_files, tmp_path, built):
    """cgx_b200/host/writer.c on the CPU: a cgx_result_t assembled from the oracle's own arrays (distinct phrases, pattern tables,
    per-query id lists, distinct rules packed into the 16-byte wire records + first / idinfo words) must print, through the C
    writer and the C loaders' vocabularies, the grammar files the oracle's writer prints -- every line, every feature digit
    (PrintResults.c:339-577).  Plain and gzip'd, one and three writer threads."""
    import gzip
    from _oracle import Oracle
    from cgx_b200 import grammar_compare as gc
    from cgx_b200._lib import RULE_WIRE_DTYPE, Result
    from cgx_b200.host import HostCorpus
    o = Oracle.from_files(micro_files["f"], micro_files["e"], micro_files["a"], micro_files["lex"])
    o.build_sa()
    oc = o.run_query_file(micro_files["q"])
    hc = HostCorpus(*(micro_files[k] for k in ("f", "q", "e", "a", "lex")))
    lay = hc.layout()
    s, qoff = lay["str"], lay["qry_off"]
    Q, T, G, D1, D2 = len(qoff) - 1, int(qoff[-1]), oc.G, oc.D1, oc.D2
    blocks = o.blocks()                                                        # {up, down, len, corpus position}, ids = first appearance
    bid = {(int(b[0]), int(b[2])): g for g, b in enumerate(blocks)}
    iv = o.intervals(5)
    phrase_id = np.full((T, 5), -1, dtype=np.int32)
    for t in range(T):
        for m in range(5):
            if iv[t, m, 0] >= 0:
                phrase_id[t, m] = bid[(int(iv[t, m, 0]), m + 1)]
    # a corpus position spelling every 1..3-gram (the writer spells a pattern's halves from the text)
    where = {}
    for m in (1, 2, 3):
        for p in range(lay["n"] - m, -1, -1):
            where[tuple(s[p:p + m].tolist())] = p
    op1 = o.onegap_patterns()
    pat1 = np.zeros((D1, 4), dtype=np.int32)
    for d in range(D1):
        ls, le = int(op1[d, 6]), int(op1[d, 7])
        pat1[d] = (where[tuple(op1[d, :ls].tolist())], ls, where[tuple(op1[d, ls + 1:ls + 1 + le].tolist())], le)
    pat2 = np.ascontiguousarray(o.twogap_patterns()[:, :2], dtype=np.int32)
    lists = [[o.query_list(w, q) for q in range(Q)] for w in (0, 1, 2)]
    # GenerateBlocks order = the order the writer derives from phrase_id
    for q in range(Q):
        seen, order = set(), []
        for g in phrase_id[qoff[q]:qoff[q + 1]].ravel().tolist():
            if g >= 0 and g not in seen:
                seen.add(g)
                order.append(g)
        assert order == lists[0][q].tolist(), q
    offs = [np.concatenate([[0], np.cumsum([len(x) for x in lists[w]])]).astype(np.int32) for w in (1, 2)]
    ids = [np.concatenate(lists[w] + [np.zeros(0, np.int32)]).astype(np.int32) for w in (1, 2)]
    n_ids = [G, 2 * G + D1, G + D2 + 2 * D1]
    keep = []                                                                  # ctypes holds raw pointers
    res = Result()
    res.Q, res.T, res.G, res.D1, res.D2 = Q, T, G, D1, D2
    ptr = lambda a, t=C.c_int32: a.ctypes.data_as(C.POINTER(t))
    for name, a in (("phrase_id", phrase_id), ("phrases", np.ascontiguousarray(blocks, dtype=np.int32)), ("pat1", pat1), ("pat2", pat2),
                    ("q1_off", offs[0]), ("q1_ids", ids[0]), ("q2_off", offs[1]), ("q2_ids", ids[1])):
        a = np.ascontiguousarray(a)
        keep.append(a)
        setattr(res, name, ptr(a))
    total = 0
    for k in range(3):
        r = o.rules(k)
        r = r[np.argsort(r["id"], kind="stable")]
        rec = r["rec"].reshape(-1, 6)                                          # tgt_start, end, gap1, gap1_1, gap2, gap2_1 (-1: no such gap)
        g = lambda v: np.where(v < 0, 15, v).astype(np.uint32)
        wire = np.zeros(len(r), dtype=RULE_WIRE_DTYPE)
        wire["tgt_start"], wire["mlfe"], wire["mlef"] = rec[:, 0], r["mlfe"], r["mlef"]
        assert rec[:, 1].max(initial=0) < 15 and rec[:, 2:].max(initial=0) < 15
        wire["span"] = (rec[:, 1].astype(np.uint32) | g(rec[:, 2]) << 4 | g(rec[:, 3]) << 8 | g(rec[:, 4]) << 12 | g(rec[:, 5]) << 16 |
                        r["pc"].astype(np.uint32) << 20)
        first = np.full(n_ids[k], -1, dtype=np.int32)
        idinfo = np.zeros(n_ids[k], dtype=np.uint32)
        uid, start, cnt = np.unique(r["id"], return_index=True, return_counts=True)
        first[uid] = start
        idinfo[uid] = r["f"][start].astype(np.uint32) | r["fs"][start].astype(np.uint32) << 9 | cnt.astype(np.uint32) << 18
        assert cnt.max(initial=0) <= 300 and r["f"].max(initial=0) <= 300 and r["fs"].max(initial=0) <= 300
        keep += [wire, first, idinfo]
        res.rules[k], res.n_rules[k], res.n_ids[k] = wire.ctypes.data, len(wire), n_ids[k]
        res.first[k], res.idinfo[k] = ptr(first), ptr(idinfo, C.c_uint32)
        total += len(wire)
    assert total > 5000
    orc = tmp_path / "orc"
    orc.mkdir()
    o.write_grammars(str(orc))
    qo = np.ascontiguousarray(qoff, dtype=np.int32)
    hc.H.cgxh_write_grammars_ex.argtypes = [C.c_char_p, C.POINTER(Result), C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    for threads, gz in ((1, 0), (3, 0), (2, 6)):
        out = tmp_path / ("mine_%d_%d" % (threads, gz))
        out.mkdir()
        assert hc.H.cgxh_write_grammars_ex(str(out).encode(), C.byref(res), ptr(qo), 0, C.byref(hc.src), C.byref(hc.tgt), threads, gz) == 0
        if gz:
            for q in range(Q):
                (out / ("grammar.%d.s" % q)).write_bytes(gzip.open(out / ("grammar.%d.s.gz" % q)).read())
                os.remove(out / ("grammar.%d.s.gz" % q))
        c = gc.compare_dirs(str(out), str(orc), rtol=0, atol=0)
        assert c["files"] == Q and c["only_a"] == 0 and c["only_b"] == 0 and c["float_mismatch"] == 0, (threads, gz, c)
        # same sequence of rule groups (source sides) in every file
        for q in range(Q):
            seq = lambda path: [k for k, _ in __import__("itertools").groupby(line.split(" ||| ")[1] for line in open(path))]
            assert seq(out / ("grammar.%d.s" % q)) == seq(orc / ("grammar.%d.s" % q)), q


def test_threaded_loader_equals_the_four_loaders_one_by_one(micro_files, tmp_path, built):
    """cgxh_load_files (source || target, then alignment || lexical file) fills the same arrays as the four loaders called in turn;
    a missing file is reported, not survived."""
    from cgx_b200.host import Align, HostCorpus, Lex, Side, load
    H = load()
    hc = HostCorpus(*(micro_files[k] for k in ("f", "q", "e", "a", "lex")))
    par = hc.layout()
    src, tgt, al, lex = Side(), Side(), Align(), Lex()
    assert H.cgxh_corpus_load(micro_files["f"].encode(), 1, C.byref(src)) == 0
    assert H.cgxh_corpus_load(micro_files["e"].encode(), 0, C.byref(tgt)) == 0
    assert H.cgxh_lex_load(micro_files["lex"].encode(), C.byref(src), C.byref(tgt), C.byref(lex)) == 0
    assert H.cgxh_alignment_load(micro_files["a"].encode(), C.byref(src), C.byref(tgt), C.byref(al)) == 0
    seq = HostCorpus.__new__(HostCorpus)
    seq.H, seq.src, seq.tgt, seq.al, seq.lex, seq.qry = H, src, tgt, al, lex, hc.qry
    one = seq.layout()
    for k, v in par.items():
        assert np.array_equal(v, one[k]) if isinstance(v, np.ndarray) else v == one[k], k
    for missing in ("f", "e", "a", "lex"):
        paths = dict(micro_files)
        paths[missing] = str(tmp_path / "absent")
        with pytest.raises(RuntimeError):
            HostCorpus(*(paths[k] for k in ("f", "q", "e", "a", "lex")))


def test_cli_usage_contract(built):
    """Exactly six positionals, otherwise help and exit 0 (Main.c:46-48)."""
    exe = os.path.join(ROOT, "bin", "strmatchcuda")
    r = subprocess.run([exe, "a", "b"], capture_output=True, text=True)
    assert r.returncode == 0 and "Please check your input arguments" in r.stdout
    r = subprocess.run([exe, "-t", "11", "a", "b", "c", "d", "e", "f"], capture_output=True, text=True)
    assert r.returncode == 0 and "finger length must be between 1 and 10" in r.stderr


def test_synth_is_deterministic_and_within_reference_limits(micro):
    from cgx_b200 import synth
    from conftest import MICRO
    c, lay = micro
    c2 = synth.generate(**MICRO)
    assert np.array_equal(c.src_words, c2.src_words) and np.array_equal(c.link_t, c2.link_t) and np.array_equal(c.qry_words, c2.qry_words)
    assert np.diff(c.src_off).max() < 255 and np.diff(c.tgt_off).max() < 255          # ExtractPair.cu:2683
    assert len(np.unique(c.src_words)) >= 100                                          # SuffixArray.cu:1175
    assert lay["str"][lay["n"] - 1] == lay["str"].max() and lay["str"][lay["n"]:].tolist() == [0, 0, 0]


def test_rule_wire_format_macros_agree_with_python_decode(tmp_path):
    """cgx_rule_t travels packed (16 B): the accessor macros of include/cgx_b200.h (what a C caller such as the grammar
    writer uses) and cgx_b200.extractor.decode_rules (what the parity tests use) must read the same fields."""
    from cgx_b200._lib import RULE_WIRE_DTYPE
    from cgx_b200.extractor import decode_rules
    rng = np.random.default_rng(5)
    n_ids, rows = 40, []
    updown = np.full((n_ids, 2), -1, dtype=np.int32)
    idinfo = np.zeros(n_ids, dtype=np.uint32)
    for cid in range(n_ids):
        cnt = int(rng.integers(0, 4))
        if cnt == 0:
            continue
        updown[cid] = (len(rows), len(rows) + cnt - 1)
        idinfo[cid] = int(rng.integers(1, 301)) | (int(rng.integers(1, 301)) << 9) | (cnt << 18)
        for _ in range(cnt):
            end = int(rng.integers(0, 15))
            g = sorted(rng.integers(0, end + 1, size=4).tolist())
            gaps = [15, 15, 15, 15]
            kind = int(rng.integers(0, 3))
            if kind >= 1:
                gaps[0], gaps[1] = g[0], g[1]
            if kind == 2:
                gaps[2], gaps[3] = g[2], g[3]
            pc = int(rng.integers(1, 301))
            span = end | gaps[0] << 4 | gaps[1] << 8 | gaps[2] << 12 | gaps[3] << 16 | pc << 20
            rows.append((int(rng.integers(0, 1 << 30)), span, float(rng.random()), float(rng.random())))
    wire = np.array(rows, dtype=RULE_WIRE_DTYPE)
    mine = decode_rules(wire, updown, idinfo)
    src = tmp_path / "dump.c"
    src.write_text(r"""
#include <stdio.h>
#include "cgx_b200.h"
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb");
    cgx_rule_t r;
    _Static_assert(sizeof(cgx_rule_t) == 16, "cgx_rule_t is 16 bytes on the wire");
    while (fread(&r, sizeof r, 1, f) == 1)
        printf("%d %d %d %d %d %d %d\n", r.tgt_start, CGX_RULE_END(&r), CGX_RULE_GAP1(&r), CGX_RULE_GAP1_END(&r), CGX_RULE_GAP2(&r),
               CGX_RULE_GAP2_END(&r), CGX_RULE_PC(&r));
    unsigned w = 7u | 300u << 9 | 211u << 18;
    printf("%d %d %d %d\n", CGX_ID_F(w), CGX_ID_FS(w), CGX_ID_RULES(w), CGX_RULE_NOGAP);
    return 0;
}
""")
    exe, raw = tmp_path / "dump", tmp_path / "rules.bin"
    raw.write_bytes(wire.tobytes())
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe), str(raw)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[len(wire)].split() == ["7", "300", "211", "15"]
    for i, line in enumerate(out[:len(wire)]):
        ts, end, g1, g1e, g2, g2e, pc = (int(x) for x in line.split())
        m = mine[i]
        none = lambda v: 255 if v == 15 else v
        assert (ts, end, none(g1), none(g1e) if g1 != 15 else 255, none(g2), none(g2e) if g2 != 15 else 255, pc) == \
               (int(m["tgt_start"]), int(m["end"]), int(m["gap1"]), int(m["gap1_1"]), int(m["gap2"]), int(m["gap2_1"]), int(m["pc"])), i
    ids = np.repeat(np.arange(n_ids), np.where(updown[:, 0] >= 0, updown[:, 1] - updown[:, 0] + 1, 0))
    assert np.array_equal(mine["id"], ids)
    assert np.array_equal(mine["f"], idinfo[ids] & 0x1FF) and np.array_equal(mine["fs"], (idinfo[ids] >> 9) & 0x1FF)



def test_writer_float_format_equals_printf(built):
    """The grammar writer prints its five float features with a hand-rolled fixed-6 formatter (writer.c put_f6) instead of
    snprintf("%f"): it must give glibc's digits for every float -- (double)x * 1e6 is exact for a 24-bit significand, so the
    only rounding is rint()'s nearest-even, the one printf applies to the exact decimal expansion."""
    from cgx_b200._lib import HOST_LIB_PATH
    L = C.CDLL(HOST_LIB_PATH)
    L.cgxh_format_f6.argtypes = [C.c_float, C.c_char_p]
    L.cgxh_format_f6.restype = C.c_int
    buf = C.create_string_buffer(64)
    rng = np.random.default_rng(3)
    vals = [0.0, -0.0, 1.0, -1.0, 99.0, 198.0, 0.5e-6, 1.5e-6, 2.5e-6, -0.4e-6, -0.5e-6, -0.6e-6, 0.0000005, 0.9999995, 0.9999994, 1e-7, 123456.789,
            2.4771213, 0.30103, 16777216.0, 3.4e38, float("inf"), -float("inf")]
    vals += list(rng.random(20000).astype(np.float32) * 5)
    vals += list(-np.log10(rng.integers(1, 301, size=20000) / rng.integers(1, 301, size=20000), dtype=np.float32))
    vals += list((rng.integers(0, 1 << 24, size=20000) * 5e-7).astype(np.float32))           # many values next to a half-way point
    vals += list(np.frombuffer(rng.bytes(4 * 20000), dtype=np.float32))                      # arbitrary bit patterns
    for v in vals:
        x = float(np.float32(v))
        if x != x:
            continue
        n = L.cgxh_format_f6(C.c_float(x), buf)
        assert buf.value.decode() == "%f" % x and n == len(buf.value), (x, buf.value)


LONG = dict(n_sent=400, n_qry=8, v_src=300, v_tgt=300, n_phrases=600, mean_len=150.0, sd_len=90.0, max_len=600, qry_mean_len=12.0, seed=5, qry_seed=6)


def test_wide_oracle_equals_narrow_on_a_corpus_that_fits_8_bits(micro):
    """SURVEY.md 8f, lifted limit.  oracle/cgx_oracle.c built with -DORC_WIDE runs the same algorithm on 16-bit alignment fields
    (sentences of 255 tokens and more, which the reference refuses: ExtractPair.cu:2683).  On a corpus the 8-bit layout holds,
    both builds must give the same records and rules."""
    from _oracle import Oracle
    _, lay = micro
    outs = []
    for wide in (False, True):
        o = Oracle.from_layout(lay, wide=wide)
        o.build_sa()
        c = o.run(lay["qry_tok"], lay["qry_off"])
        outs.append(((c.G, c.D1, c.D2, c.hits1, c.hits2), [o.records(k).tobytes() for k in range(3)], [o.rules(k).tobytes() for k in range(3)]))
        o.close()
    assert outs[0] == outs[1]


def test_long_sentences_load_into_the_16_bit_layout(tmp_path, built):
    """The C loaders on a corpus with sentences of up to ~400 tokens: the 16-bit alignment layout (wide = 1) is chosen, its fields
    equal a numpy parse of the same alignment file, and the narrow oracle refuses the corpus while the wide one runs it."""
    from _oracle import Oracle
    from cgx_b200 import synth
    from cgx_b200.host import HostCorpus
    files = synth.write_text(synth.generate(**LONG), str(tmp_path), "corpus")
    src_len = [len(l.split()) for l in open(files["f"])]
    tgt_len = [len(l.split()) for l in open(files["e"])]
    assert max(src_len) >= 300
    lay = HostCorpus(files["f"], files["q"], files["e"], files["a"], files["lex"]).layout()
    assert lay["wide"] and lay["RLP"].dtype == np.uint64 and lay["L_tar"].dtype == np.uint16
    s_off = np.concatenate([[0], np.cumsum(np.array(src_len) + 1)])
    t_off = np.concatenate([[0], np.cumsum(np.array(tgt_len) + 1)])
    n, m = lay["n"], lay["m"]
    L = np.full(n, 65535, np.int64); R = np.full(n, 65535, np.int64); Lt = np.full(m, 65535, np.int64); Rt = np.full(m, 65535, np.int64)
    for q, line in enumerate(open(files["a"])):
        for pair in line.split():
            a, b = (int(x) for x in pair.split("-"))
            si, ti = s_off[q] + a, t_off[q] + b
            L[si], R[si] = (b, b) if L[si] == 65535 else (min(L[si], b), max(R[si], b))
            Lt[ti], Rt[ti] = (a, a) if Lt[ti] == 65535 else (min(Lt[ti], a), max(Rt[ti], a))
    assert np.array_equal(lay["L_tar"], Lt.astype(np.uint16)) and np.array_equal(lay["R_tar"], Rt.astype(np.uint16))
    rlp = lay["RLP"]
    for q in range(len(src_len)):
        a, b = s_off[q], s_off[q + 1] - 1                                           # tokens a..b-1, EOS at b
        w = rlp[a:b]
        assert np.array_equal((w >> 48) & 0xFFFF, L[a:b]) and np.array_equal((w >> 32) & 0xFFFF, R[a:b])
        assert np.array_equal((w >> 16) & 0xFFFF, np.arange(b - a))                  # position in the sentence, not wrapped at 256
        assert int(rlp[b]) == t_off[q + 1]                                          # the word at the EOS: offset of the next target sentence
    with pytest.raises(RuntimeError):
        Oracle.from_files(files["f"], files["e"], files["a"], files["lex"])          # 8-bit fields: "sentence too long", like the reference
    o = Oracle.from_files(files["f"], files["e"], files["a"], files["lex"], wide=True)
    o.build_sa()
    o.run_query_file(files["q"])
    assert o.counts().G > 100 and len(o.rules(1)) > 1000
    o.close()
