"""CPU: pin the oracle against outputs of the reference itself (tests/golden/, made by tools/make_golden.py
from oracle/_ref/strmatchcuda_dump on a B200).  Deterministic stages must agree bit-for-bit; the stages the
reference computes nondeterministically (unstable comparator sorts, atomics) are compared on the subsets
that are deterministic, with the known reference defects spelled out."""
import collections
import os

import numpy as np
import pytest

from cgx_b200 import grammar_compare as gc


def ms(a):
    return collections.Counter(map(tuple, np.asarray(a).tolist()))


def test_params_match(golden):
    from conftest import MICRO
    assert eval(str(golden[0]["params"])) == MICRO


def test_suffix_array_bit_exact(micro_oracle, golden):
    assert np.array_equal(micro_oracle.sa(), golden[0]["sa"])


def test_builtin_sa_equals_reference_dc3(micro, micro_oracle):
    """The oracle's own prefix-doubling SA against SuffixArray.c (oracle/_ref/libref_sa.so) when present."""
    import ctypes as C
    from _oracle import REF_SA_PATH
    if not os.path.exists(REF_SA_PATH):
        pytest.skip("oracle/_ref/libref_sa.so not built (no /root/reference)")
    lay = micro[1]
    L = C.CDLL(REF_SA_PATH)
    L.ref_sa_build.restype = C.c_double
    s = np.ascontiguousarray(lay["str"], dtype=np.int32)
    sa = np.empty(lay["n"], dtype=np.int32)
    L.ref_sa_build(s.ctypes.data_as(C.c_void_p), C.c_int(lay["n"]), C.c_int(lay["src_last"]), sa.ctypes.data_as(C.c_void_p), None)
    assert np.array_equal(sa, micro_oracle.sa())


def test_lookup_bit_exact(micro_oracle, golden):
    g = golden[0]
    o = micro_oracle
    lg = o.longest()
    assert np.array_equal(lg, g["result_two"]["longestmatch"])
    iv = o.intervals(cap=int(lg.max()))
    for t in range(len(lg)):
        if lg[t] >= 1:
            assert (iv[t, 0, 0], iv[t, 0, 1]) == (g["result_two"]["up"][t], g["result_two"]["down"][t])
        for m in range(2, int(lg[t]) + 1):
            rc = g["result_connect"][g["connectoffset"][t] + m - 2]
            assert (iv[t, m - 1, 0], iv[t, m - 1, 1]) == (rc["up"], rc["down"])


def test_blocks_and_frequent_tokens(micro_oracle, golden):
    g = golden[0]
    assert np.array_equal(micro_oracle.blocks(), np.stack([g["blocks"][k] for k in ("start", "end", "matchlen", "string_start")], 1))
    assert np.array_equal(micro_oracle.frequent(), g["frequentList"])
    pi = micro_oracle.precomp_index()
    assert np.array_equal(pi[:, 0], g["precomp_index"]["start"].astype(np.int32))
    assert np.array_equal(pi[:, 1], g["precomp_index"]["end"].astype(np.int32))
    pl = micro_oracle.precomp_list()
    assert np.array_equal(pl[:, 0], g["precomp_onegap"]["start"].astype(np.int32))
    assert np.array_equal(pl[:, 1], g["precomp_onegap"]["length"].astype(np.int32))
    # featureMissingCount: the reference resets the counter from every thread of the CTA while others already
    # count (GappyLook.cu:759) -- it can only lose increments
    assert np.all(micro_oracle.feature_missing() >= g["featureMissingCount"])
    assert np.mean(micro_oracle.feature_missing() == g["featureMissingCount"]) > 0.99


def _scrambled(o_pat, ref_search, ref_pattern, ref_inst):
    """Patterns whose (key, value) pair the reference's thrust::sort_by_key mis-paired (its comparator returns
    true on equality, SuffixArray.cu:51-67 -- not a strict weak order): the representative instance stored in
    oneGapSearch no longer spells the pattern it is filed under."""
    return set()


def test_onegap_patterns_and_hits(micro_oracle, golden):
    g = golden[0]
    o = micro_oracle
    p1 = o.onegap_patterns()
    rs, rp = g["oneGapSearch"], g["onegapPattern"]
    assert len(p1) == len(rs)
    assert np.array_equal(rp["pattern"][rs["position"]], p1[:, :5])          # same distinct patterns, same ids
    h, rh = o.onegap_hits(), g["oneGapSA"]
    ref = np.stack([rh["position"].astype(np.int64), rh["str_position"].astype(np.int64), rh["length"].astype(np.int64)], 1)
    ref = ref[np.lexsort((ref[:, 2], ref[:, 1], ref[:, 0]))]
    mine_cnt = np.bincount(h[:, 0], minlength=len(p1))
    ref_cnt = np.bincount(ref[:, 0], minlength=len(p1))
    bad = set(np.nonzero(mine_cnt != ref_cnt)[0].tolist())
    # the reference's invalid comparator may mis-pair a handful of patterns per run; everything else is exact
    assert len(bad) <= max(2, len(p1) // 2000)
    keep_m = np.array([x not in bad for x in h[:, 0]])
    keep_r = np.array([x not in bad for x in ref[:, 0]])
    assert np.array_equal(h[keep_m].astype(np.int64), ref[keep_r])


def test_twogap_patterns_and_hits(micro_oracle, golden):
    g = golden[0]
    o = micro_oracle
    p2 = o.twogap_patterns()
    r2, rp2 = g["twoGapSearch"], g["twogapPattern"]
    assert len(p2) == len(r2)
    assert np.array_equal(p2[:, 1], rp2["pattern"][r2["position"], 0])
    h, rh = o.twogap_hits(), g["twoGapSA"]
    ref = np.stack([rh[k].astype(np.int64) for k in ("position", "str_position", "length", "length2")], 1)     # the reference's own order
    mine_cnt = np.bincount(h[:, 0], minlength=len(p2))
    ref_cnt = np.bincount(ref[:, 0], minlength=len(p2))
    bad = set(np.nonzero(mine_cnt != ref_cnt)[0].tolist())
    assert len(bad) <= max(4, len(p2) // 500)
    mine = h[np.array([x not in bad for x in h[:, 0]])].astype(np.int64)
    ref = ref[np.array([x not in bad for x in ref[:, 0]])]
    # (1) the hit SET is the reference's, and so is the order of the (pattern, position) groups
    canon = lambda a: a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]
    assert np.array_equal(canon(mine), canon(ref))
    assert np.array_equal(mine[:, :2], ref[:, :2])
    # (2) inside a group (hits that tie on the reference's sort key) the reference keeps its atomicAdd arrival order, which is
    # the warp scheduler's; the oracle's (width, length) order is the lock-step one and agrees with it on most groups -- the
    # (length, width) order used before round 2 agrees on far fewer (measured here on the reference's own dump)
    starts = np.nonzero(np.r_[True, (np.diff(ref[:, 0]) != 0) | (np.diff(ref[:, 1]) != 0)])[0]
    ends = np.r_[starts[1:], len(ref)]
    groups = [(a, b) for a, b in zip(starts, ends) if b - a > 1]
    same = sum(np.array_equal(mine[a:b], ref[a:b]) for a, b in groups)
    old = sum(np.array_equal(canon(ref[a:b]), ref[a:b]) for a, b in groups)
    assert len(groups) > 500 and same >= 0.9 * len(groups) and same > old * 1.5, (len(groups), same, old)


def test_extraction_records(micro_oracle, golden):
    g = golden[0]
    o = micro_oracle
    c = o.counts()
    sep1, sep2a, sep2b, glob = (int(x) for x in g["separators"])
    assert glob == c.G
    r0, ref0 = o.records(0), g["out_res"]
    assert ms(r0[:, :3]) == ms(np.stack([ref0["blocknumber"], ref0["tar_start"], ref0["tar_end"].astype(np.int32)], 1))
    r1, ref1 = o.records(1), g["oneGapRule"]
    ref1a = np.stack([ref1["id"]] + [ref1[k].astype(np.int32) for k in ("start", "end", "gap1", "gap1_1")], 1)
    assert ms(r1[r1[:, 0] < 2 * c.G][:, :5]) == ms(ref1a[:sep1])                     # Xab, abX: sampled in SA order, deterministic
    r2, ref2 = o.records(2), g["twoGapRule"]
    ref2a = np.stack([ref2["id"]] + [ref2[k].astype(np.int32) for k in ("start", "end", "gap1", "gap1_1", "gap2", "gap2_1")], 1)
    assert ms(r2[r2[:, 0] < c.G]) == ms(ref2a[:sep2a])                                # XabX
    # aXb-seeded records: sampled from hit lists whose tie order is arbitrary in the reference; patterns with
    # <= 65 hits are not sampled and must agree exactly
    p1 = o.onegap_patterns()
    cnt = np.where(p1[:, 8] >= 0, p1[:, 9] - p1[:, 8] + 1, 0)
    h1, pidx = o.onegap_hits(), o.precomp_index()
    for d in np.nonzero(cnt == 1)[0]:
        if h1[p1[d, 8], 2] == 0:
            pi = h1[p1[d, 8], 1]
            cnt[d] = pidx[pi, 1] - pidx[pi, 0] + 1
    ref_cnt = np.where(g["oneGapSearch"]["start"] >= 0, g["oneGapSearch"]["end"] - g["oneGapSearch"]["start"] + 1, 0)
    small = set(np.nonzero((cnt <= 65) & ((cnt == ref_cnt) | (ref_cnt == 1)))[0].tolist())
    ref1a[sep1:, 0] += 2 * c.G
    a = r1[r1[:, 0] >= 2 * c.G]
    b = ref1a[sep1:]
    fa = a[np.array([(x - 2 * c.G) in small for x in a[:, 0]], dtype=bool)]
    fb = b[np.array([(x - 2 * c.G) in small for x in b[:, 0]], dtype=bool)]
    assert len(fa) > 0 and ms(fa[:, :5]) == ms(fb)


def test_grammars_against_reference_run(micro, micro_files, micro_oracle, golden, tmp_path):
    """End to end through the text loaders and the writer: >= 97 % of the reference's rule lines identical
    (north-star bar: 90 %); the rest is the reference's sampling-tie nondeterminism."""
    from _oracle import Oracle
    o = Oracle.from_files(micro_files["f"], micro_files["e"], micro_files["a"], micro_files["lex"])
    o.build_sa()
    o.run_query_file(micro_files["q"])
    out = tmp_path / "orc"
    out.mkdir()
    o.write_grammars(str(out))
    ref = tmp_path / "ref"
    ref.mkdir()
    for q, lines in golden[1].items():
        (ref / ("grammar.%d.s" % q)).write_text("".join(lines))
    r = gc.compare_dirs(str(out), str(ref))
    assert r["files"] == len(golden[1]) and r["missing_files"] == 0
    assert r["frac_equal"] >= 0.97, r
