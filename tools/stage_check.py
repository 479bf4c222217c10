"""Stage-by-stage comparison of the CUDA path with the CPU oracle on one synthetic corpus (GPU box).
Development / debugging aid; the pytest -m gpu suite runs the same comparisons as assertions."""
import collections
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cgx_b200 import synth  # noqa: E402
from cgx_b200.extractor import GrammarExtractor  # noqa: E402
from _oracle import Oracle  # noqa: E402


def ms(a):
    return collections.Counter(map(tuple, np.asarray(a).tolist()))


def main():
    ns, nq, v, nph = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (2000, 20, 500, 800)))
    c = synth.generate(ns, nq, v_src=v, v_tgt=v, n_phrases=nph)
    lay = synth.text_layout(c)
    print("corpus n=%d m=%d Q=%d T=%d" % (lay["n"], lay["m"], c.n_qry, len(lay["qry_tok"])))
    t0 = time.time()
    o = Oracle.from_layout(lay)
    o.build_sa()
    oc = o.run(lay["qry_tok"], lay["qry_off"])
    print("oracle %.2fs" % (time.time() - t0))
    ex = GrammarExtractor(0)
    info = ex.build_index(lay)
    print("index", info)
    ok = True

    def check(name, cond):
        nonlocal ok
        print(("OK   " if cond else "FAIL ") + name)
        ok &= bool(cond)

    check("suffix array", np.array_equal(ex.suffix_array(), o.sa()))
    check("frequent tokens", np.array_equal(ex.frequent_tokens(), o.frequent()))
    s = lay["str"]
    n = lay["n"]
    for m in (1, 2, 3):
        inv = ex.occurrence_list(m)
        key = np.zeros(n, dtype=np.int64)
        for j in range(m):
            key = key * (int(s.max()) + 1) + s[np.arange(n) + j]
        order = np.lexsort((np.arange(n), key))
        check("occurrence list m=%d" % m, np.array_equal(inv, order))
    res = ex.extract(lay["qry_tok"], lay["qry_off"])
    print("batch", res.info)
    T = oc.T
    lg = ex.debug_fetch("longest", T)
    check("longest (capped 5)", np.array_equal(lg, np.minimum(o.longest(), 5)))
    iv = ex.debug_fetch("intervals", T * 10).reshape(T, 5, 2)
    check("intervals", np.array_equal(iv, o.intervals(5)))
    check("G", res.G == oc.G)
    ob = o.blocks()
    check("phrase set", ms(res.phrases) == ms(ob))
    check("enu1/D1", res.info["enu1"] == oc.enu1 and res.D1 == oc.D1)
    op = o.onegap_patterns()
    mine = np.stack([s[res.pat1[:, 0] + k] if True else 0 for k in range(1)], 1) if res.D1 else None
    # pattern tokens
    def toks(p):
        a, ls, b, le = (int(x) for x in p[:4])
        t = list(s[a:a + ls]) + [-1] + list(s[b:b + le])
        return t + [-2] * (5 - len(t))
    if res.D1 == oc.D1:
        mp = np.array([toks(p) for p in res.pat1], dtype=np.int32)
        check("one-gap patterns (order + tokens)", np.array_equal(mp, op[:, :5]))
    oh = o.onegap_hits()
    # oracle marker patterns hold a single marker record; expand them to the pair list for comparison
    pidx, plist = o.precomp_index(), o.precomp_list()
    exp = []
    for h in oh:
        if h[2] == 0:
            a, b = pidx[h[1]]
            for k in range(a, b + 1):
                exp.append((h[0], plist[k, 0], plist[k, 1]))
        else:
            exp.append(tuple(h))
    exp = np.array(sorted(exp), dtype=np.int32).reshape(-1, 3)
    h1 = ex.debug_fetch("hits1", int(res.info["hits1"]) * 3, 3)
    check("one-gap hits (%d)" % len(h1), np.array_equal(h1, exp))
    if not np.array_equal(h1, exp):
        print("   mine", len(h1), "oracle", len(exp), "only mine", sum((ms(h1) - ms(exp)).values()), "only oracle", sum((ms(exp) - ms(h1)).values()))
    # featureMissingCount of marker patterns
    miss = o.feature_missing()
    okm = True
    pf = ex.debug_fetch("pat1_full", res.D1 * 8, 8)
    for d in range(res.D1):
        if pf[d, 6] >= 0 and pf[d, 5] > 0:
            okm &= int(pf[d, 7]) == int(miss[pf[d, 6]])
    check("featureMissingCount of frequent-pair patterns", okm)
    check("enu2/D2", res.info["enu2"] == oc.enu2 and res.D2 == oc.D2)
    if res.D2 == oc.D2:
        check("two-gap patterns", np.array_equal(res.pat2[:, :2], o.twogap_patterns()[:, :2]))
    h2 = ex.debug_fetch("hits2", int(res.info["hits2"]) * 4, 4)
    check("two-gap hits (%d)" % len(h2), np.array_equal(h2, o.twogap_hits()))
    for k, name in enumerate(("rec_ab", "rec_1", "rec_2")):
        mine = ex.debug_fetch(name, int(res.info[("n_ab", "n_1gap", "n_2gap")[k]]) * 7, 7)
        orc = o.records(k)
        if k == 0:
            # block ids differ (sorted vs first-appearance): compare through (up,len)
            mg = res.phrases[mine[:, 0]][:, [0, 2]]
            og = ob[orc[:, 0]][:, [0, 2]]
            a = ms(np.concatenate([mg, mine[:, 1:3]], 1))
            b = ms(np.concatenate([og, orc[:, 1:3]], 1))
        else:
            a, b = None, None
        if k == 0:
            check("records ab (%d)" % len(mine), a == b)
        else:
            print("     %s: mine %d oracle %d" % (name, len(mine), len(orc)))
    for k in range(3):
        print("     rules kind %d: mine %d oracle %d" % (k, len(res.rules[k]), len(o.rules(k))))
    # final rule sets per query, as multisets of printed lines (floats rounded)
    outdir = "/tmp/cgx_stage_check"
    os.makedirs(outdir + "/mine", exist_ok=True)
    for q in range(res.Q):
        with open("%s/mine/grammar.%d.s" % (outdir, q), "w") as fh:
            for line in res.grammar_lines(q, lay):
                fh.write(line + "\n")
    print("ALL OK" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
