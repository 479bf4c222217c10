"""Shared parity checks of the GPU tests: the CUDA path (through the C ABI) against the CPU oracle, stage by stage.
TEST INFRASTRUCTURE ONLY (imports tests/_oracle.py)."""
import collections

import numpy as np


def ms(a):
    return collections.Counter(map(tuple, np.asarray(a).tolist()))


def expand_marker_hits(o):
    """The oracle keeps the reference's single marker record for frequent-pair patterns; expand it to the
    precomputed pair list it stands for."""
    oh, pidx, plist = o.onegap_hits(), o.precomp_index(), o.precomp_list()
    marker = oh[:, 2] == 0
    rows = [oh[~marker]]
    for h in oh[marker]:
        a, b = pidx[h[1]]
        k = np.arange(a, b + 1)
        rows.append(np.stack([np.full(len(k), h[0], dtype=np.int32), plist[k, 0], plist[k, 1]], 1))
    allr = np.concatenate(rows).astype(np.int32) if rows else np.zeros((0, 3), np.int32)
    return allr[np.lexsort((allr[:, 2], allr[:, 1], allr[:, 0]))].reshape(-1, 3)


def remap_ids(rows, res, o, kind):
    """Phrase ids differ (sorted (up,len) order here, first-appearance order in the reference): map the oracle's
    converted ids into ours through (up, len)."""
    ob = o.blocks()
    mine = {(int(p[0]), int(p[2])): g for g, p in enumerate(res.phrases)}
    g_map = np.array([mine[(int(b[0]), int(b[2]))] for b in ob], dtype=np.int64)
    G = res.G
    rows = rows.copy()
    ids = rows[:, 0].astype(np.int64)
    if kind == 0:
        rows[:, 0] = g_map[ids]
    elif kind == 1:
        sel = ids < 2 * G
        rows[sel, 0] = g_map[ids[sel] % G] + (ids[sel] // G) * G
    else:
        sel = ids < G
        rows[sel, 0] = g_map[ids[sel]]
    return rows


def assert_full_parity(ex, res, lay, o):
    """Every stage of one batch, bit-exact against the oracle `o` (already run on the same batch): suffix array, longest
    matches and n-gram intervals, distinct phrases, one- and two-gap pattern tables (same ids, same order), hit lists,
    featureMissingCount, extraction records (multisets per array) and distinct rules (id, paircount, f, fsample).  The two
    float features are checked with their 1e-5 tolerance on the grammar lines by the callers."""
    oc = o.counts()
    s = lay["str"]
    assert np.array_equal(ex.suffix_array(), o.sa()), "suffix array"
    T = res.T
    assert np.array_equal(ex.debug_fetch("longest", T), np.minimum(o.longest(), 5)), "longest match"
    assert np.array_equal(ex.debug_fetch("intervals", T * 10).reshape(T, 5, 2), o.intervals(5)), "n-gram intervals"
    assert (res.G, res.D1, res.D2, res.info["enu1"], res.info["enu2"]) == (oc.G, oc.D1, oc.D2, oc.enu1, oc.enu2)
    assert ms(res.phrases) == ms(o.blocks()), "distinct phrases"
    # one-gap patterns: same ids in the same order, spelled token by token
    p1 = res.pat1
    op1 = o.onegap_patterns()[:, :5]
    a, ls, b, le = (p1[:, k].astype(np.int64) for k in range(4))
    mine = np.full((res.D1, 5), -2, dtype=np.int32)
    for j in range(3):
        sel = ls > j
        mine[sel, j] = s[a[sel] + j]
    rows = np.arange(res.D1)
    mine[rows, ls] = -1
    for j in range(3):
        sel = le > j
        mine[rows[sel], ls[sel] + 1 + j] = s[b[sel] + j]
    assert np.array_equal(mine, op1), "one-gap pattern table"
    assert np.array_equal(res.pat2[:, :2], o.twogap_patterns()[:, :2]), "two-gap pattern table"
    assert np.array_equal(ex.debug_fetch("hits1", int(res.info["hits1"]) * 3, 3), expand_marker_hits(o)), "one-gap hit list"
    assert np.array_equal(ex.debug_fetch("hits2", int(res.info["hits2"]) * 4, 4), o.twogap_hits()), "two-gap hit list"
    miss = o.feature_missing()
    pf = ex.debug_fetch("pat1_full", res.D1 * 8, 8)             # device records: + hit_start, hit_count, marker_pair, fs_extra
    assert np.array_equal(pf[:, :4], p1)
    mk = (pf[:, 6] >= 0) & (pf[:, 5] > 0)
    assert np.array_equal(pf[mk, 7], miss[pf[mk, 6]]), "featureMissingCount"
    for k, (name, cnt) in enumerate((("rec_ab", "n_ab"), ("rec_1", "n_1gap"), ("rec_2", "n_2gap"))):
        mine_r = ex.debug_fetch(name, int(res.info[cnt]) * 7, 7)
        orc_r = remap_ids(o.records(k), res, o, k)
        assert ms(mine_r) == ms(orc_r), name
    for k in range(3):
        m, ref = res.rules[k], o.rules(k)
        rid = remap_ids(ref["id"].reshape(-1, 1).astype(np.int64), res, o, k)[:, 0]
        key_m = collections.Counter(zip(m["id"].tolist(), m["pc"].tolist(), m["f"].tolist(), m["fs"].tolist()))
        key_r = collections.Counter(zip(rid.tolist(), ref["pc"].tolist(), ref["f"].tolist(), ref["fs"].tolist()))
        assert key_m == key_r, "distinct rules, kind %d" % k
