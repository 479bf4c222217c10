"""GPU box: work distribution of the gappy joins at a given scale (development aid).
    python tools/join_stats.py <sentence pairs> <queries> [vocab]"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgx_b200 import synth
from cgx_b200.extractor import GrammarExtractor

ns, nq = int(sys.argv[1]), int(sys.argv[2])
v = int(sys.argv[3]) if len(sys.argv) > 3 else 50000
c = synth.generate(ns, nq, v_src=v, v_tgt=v)
lay = synth.text_layout(c)
ex = GrammarExtractor(0)
ex.build_index(lay)
res = ex.extract(lay["qry_tok"], lay["qry_off"])
D1, D2 = res.D1, res.D2
pd = ex.debug_fetch("pat1_dev", D1 * 4, 4)
p1 = ex.debug_fetch("pat1_full", D1 * 8, 8)
nA = pd[:, 1] - pd[:, 0] + 1
nB = pd[:, 3] - pd[:, 2] + 1
W = np.minimum(nA, nB).astype(np.int64)
hc = p1[:, 5].astype(np.int64)
ls, le = p1[:, 1], p1[:, 3]
print("D1 %d  W %.3e  hits %.3e  sum(max) %.3e" % (D1, W.sum(), hc.sum(), np.maximum(nA, nB).astype(np.int64).sum()))
print("by (ls,le): count, W, hits")
for a in (1, 2, 3):
    for b in (1, 2, 3):
        m = (ls == a) & (le == b)
        if m.any():
            print("  (%d,%d) %9d  W %.3e  hits %.3e" % (a, b, m.sum(), W[m].sum(), hc[m].sum()))
print("by log2(W): count, W, hits, driveA share")
lg = np.floor(np.log2(np.maximum(W, 1))).astype(int)
for k in range(lg.max() + 1):
    m = lg == k
    if m.any():
        print("  2^%-2d %9d  W %.3e  hits %.3e  other-list mean %.3e" % (k, m.sum(), W[m].sum(), hc[m].sum(), np.maximum(nA, nB)[m].mean()))
# heavy single-token pairs by token frequency rank
cnt = np.bincount(lay["str"][: lay["n"]])
order = np.argsort(-cnt, kind="stable")
rank = np.empty_like(order); rank[order] = np.arange(len(order))
s = lay["str"]
ra = rank[s[p1[:, 0]]]; rb = rank[s[p1[:, 2]]]
single = (ls == 1) & (le == 1)
for F in (128, 256, 512, 1024, 2048, 4096, 8192):
    m = single & (ra < F) & (rb < F)
    print("single-token pairs with both ranks < %5d: %9d patterns  W %.3e  hits %.3e ; rest W %.3e (max W %d)" % (F, m.sum(), W[m].sum(), hc[m].sum(), W[~m].sum(), W[~m].max()))
# two-gap
p2 = ex.debug_fetch("pat2_full", res.D2 * 4, 4)     # device records: {pat1, ctok, hit_start, hit_count}
nH = p1[p2[:, 0], 5].astype(np.int64)
nC = cnt[p2[:, 1]].astype(np.int64)
W2 = np.minimum(nH, nC)
print("D2 %d  W2 %.3e  hits2 %.3e  sum nH %.3e  driveH share of W %.3f" % (D2, W2.sum(), p2[:, 3].astype(np.int64).sum(), nH.sum(), W2[nH <= nC].sum() / max(1, W2.sum())))
lg = np.floor(np.log2(np.maximum(W2, 1))).astype(int)
for k in range(lg.max() + 1):
    m = lg == k
    if m.any():
        print("  2^%-2d %9d  W2 %.3e  hits %.3e" % (k, m.sum(), W2[m].sum(), p2[m, 3].astype(np.int64).sum()))
rc = rank[p2[:, 1]]
par = p2[:, 0]
for F in (256, 1024, 4096):
    m = (ra[par] < F) & (rb[par] < F) & (rc < F)
    print("two-gap with all three ranks < %5d: %9d patterns W2 %.3e hits %.3e ; rest W2 %.3e" % (F, m.sum(), W2[m].sum(), p2[m, 3].astype(np.int64).sum(), W2[~m].sum()))
ex.close()
