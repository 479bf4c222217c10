"""GPU box: BASELINE.json configs[3] on ONE B200 -- a ~1 G-token corpus resident in HBM, suffix array + auxiliary index built,
query batches extracted -- with size-independent checks (the oracle cannot run at this size).

    python tools/c4_probe.py [tokens=1040000000] [vocab=50000] [queries=2000] [batch=500]

The corpus is synthesised on the GPU (the Python generator of cgx_b200/synth.py would need an hour and ~100 GB of host memory):
Zipf tokens, 25-token sentences, the target side a token-wise image of the source with 10 % of the links dropped, alignment
fields in the reference's layout (ExtractPair.cu:2639-2739), a lexical table over the 200 k most frequent (f, e) pairs.  Queries
are sentences of the corpus itself.  Checks: the suffix array is a permutation with sampled neighbours in order; per-query
grammar lines do not depend on the batch composition (one batch of 2 x k queries == two batches of k); every rule's target span
lies inside one target sentence.  At 8 GPUs configs[3] shards the QUERIES: every GPU holds this same index."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cgx_b200.extractor import GrammarExtractor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_040_000_000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
NQ = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
BATCH = int(sys.argv[4]) if len(sys.argv) > 4 else 200
SENT = 26                                    # 25 tokens + EOS
n = (n // SENT) * SENT + 2                   # whole sentences + the reference's trailer "1, V+2" (Start.cu:321-330)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1234)
t0 = time.time()
s = torch.empty(n + 3, dtype=torch.int32, device=dev)
step = 1 << 27
for a in range(0, n, step):
    b = min(n, a + step)
    u = torch.rand(b - a, device=dev, generator=g)
    s[a:b] = (torch.exp(u * float(np.log(V))).to(torch.int32).clamp_(1, V) + 1)
    del u
idx = torch.arange(n, device=dev)
P = (idx % SENT).to(torch.int32)
eos = P == SENT - 1
s[:n][eos] = 1
s[n - 2] = 1
s[n - 1] = V + 2
s[n:] = 0
# target side: same sentence layout, token image e = (f * 7919) % V + 2; 10 % of the links dropped
tgt = torch.where(s[:n] >= 2, (s[:n].long() * 7919 % V + 2).to(torch.int32), s[:n])
tgt[n - 1] = V + 2
tgt = torch.cat([tgt, torch.zeros(3, dtype=torch.int32, device=dev)])
aligned = (torch.rand(n, device=dev, generator=g) >= 0.10) & ~eos
aligned[n - 2:] = False
L = torch.where(aligned, P, torch.full_like(P, 255))
rlp = ((L.long() << 24) | (L.long() << 16) | (P.long() << 8)).to(torch.int64)
rlp[eos] = (idx[eos] + 1)                    # the word at an EOS: target offset of the next sentence
rlp[n - 2:] = 0
rlp = rlp.to(torch.uint32) if hasattr(torch, "uint32") else rlp
lt = L.to(torch.uint8)
lay = dict(str=s.cpu().numpy(), n=n, tgt=tgt.cpu().numpy(), m=n, RLP=rlp.cpu().numpy().astype(np.uint32), L_tar=lt.cpu().numpy(), R_tar=lt.cpu().numpy())
# lexical table: (f, e = image(f)) for the most frequent tokens, plus NULL rows
f = np.arange(2, min(V, 200_000) + 2, dtype=np.int32)
e = (f.astype(np.int64) * 7919 % V + 2).astype(np.int32)
lay.update(lex_f=np.concatenate([f, f, np.full_like(f, -1)]), lex_e=np.concatenate([e, np.full_like(e, -1), e]),
           lex_v1=np.concatenate([np.full(len(f), 0.5, np.float32), np.full(len(f), 0.01, np.float32), np.full(len(f), 0.02, np.float32)]),
           lex_v2=np.concatenate([np.full(len(f), 0.4, np.float32), np.full(len(f), 0.03, np.float32), np.full(len(f), 0.04, np.float32)]))
del idx, P, eos, aligned, L, rlp, lt, tgt, s
torch.cuda.empty_cache()                     # the corpus lives on the host from here; HBM belongs to the index and the batches
print("corpus: %d tokens synthesised in %.1f s" % (n, time.time() - t0), flush=True)

ex = GrammarExtractor(0)
t0 = time.time()
info = ex.build_index(lay)
print("index: SA %.1f ms (%d rounds, %d-bit keys), auxiliary %.1f ms, %.1f GB resident, wall %.1f s" % (info["sa_build_ms"], info["sa_rounds"], info["sa_key_bits"],
      info["aux_build_ms"], info["index_bytes"] / 1e9, time.time() - t0), flush=True)
sa = ex.suffix_array()
assert np.array_equal(np.bincount(sa.astype(np.int64) >> 8, minlength=(n >> 8) + 1)[: n >> 8], np.full(n >> 8, 256)), "suffix array is not a permutation"
rng = np.random.default_rng(7)
k = rng.integers(0, n - 1, 200_000)
hs = lay["str"]
bad = 0
for a, b in zip(sa[k][:20000].tolist(), sa[k + 1][:20000].tolist()):
    x, y = hs[a:a + 40], hs[b:b + 40]
    m = min(len(x), len(y))
    d = np.nonzero(x[:m] != y[:m])[0]
    if len(d) and x[d[0]] > y[d[0]]:
        bad += 1
assert bad == 0, "%d sampled suffix-array neighbours out of order" % bad
del sa
# queries: sentences of the corpus
sent = rng.integers(0, n // SENT - 1, NQ)
qtok = np.concatenate([hs[i * SENT:i * SENT + SENT - 1] for i in sent]).astype(np.int32)
qoff = (np.arange(NQ + 1) * (SENT - 1)).astype(np.int32)
t0 = time.time()
infos = ex.extract_stream(qtok, qoff, batch_queries=BATCH)
wall = time.time() - t0
dev_ms = sum(i["ms_total"] for i in infos)
print("extraction: %d queries in %d batches, device %.1f ms (%.0f q/s), wall %.1f s (first batches grow the buffers)" % (NQ, len(infos), dev_ms, NQ / (dev_ms / 1e3), wall), flush=True)
for i in infos[:3] + infos[-2:]:
    print("  batch %d..%d: hits %d / %d, rules %s, device %.1f ms (join %.1f, extract %.1f, aggregate %.1f)" % (i["q0"], i["q1"], i["hits1"], i["hits2"], i["rules"], i["ms_total"],
          i["ms_join"], i["ms_extract"], i["ms_aggregate"]), flush=True)
# batch transparency on 2 x 40 queries, and target spans inside one sentence
fake = dict(lay, src_names=np.arange(V + 8), tgt_names=np.arange(V + 8))
kq = 40
one = ex.extract(qtok[: qoff[2 * kq]], qoff[: 2 * kq + 1])
lines_one = [sorted(one.grammar_lines(q, fake)) for q in range(2 * kq)]
for kind in range(3):
    r = one.rules[kind]
    if len(r):
        a = r["tgt_start"].astype(np.int64)
        assert np.array_equal(a // SENT, (a + r["end"]) // SENT), "a rule's target span crosses a sentence boundary"
a_ = ex.extract(qtok[: qoff[kq]], qoff[: kq + 1])
b_ = ex.extract(qtok[qoff[kq]: qoff[2 * kq]], qoff[kq: 2 * kq + 1] - qoff[kq])
lines_two = [sorted(a_.grammar_lines(q, fake)) for q in range(kq)] + [sorted(b_.grammar_lines(q, fake)) for q in range(kq)]
assert lines_one == lines_two, "per-query output depends on the batch composition"
print("checks ok: permutation, 20000 sampled neighbours in order, %d grammar lines of %d queries independent of the batching, spans inside sentences" % (sum(len(x) for x in lines_one), 2 * kq))
out = {"tokens": n, "vocabulary": V, "index": info, "queries": NQ, "batches": len(infos), "device_ms": dev_ms, "queries_per_s_device": NQ / (dev_ms / 1e3),
       "steady_batches": [{k: i[k] for k in ("q0", "q1", "hits1", "hits2", "ms_total", "ms_join", "ms_extract", "ms_aggregate")} for i in infos]}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "c4_probe.json"), "w"), indent=1)
ex.close()
