"""Deterministic synthetic word-aligned parallel corpora.

The reference ships no data (its ``toy/`` directory is absent from the snapshot), so every
fixture and every benchmark input comes from here.  A corpus is a list of source sentences built
by concatenating phrases drawn (Zipfian) from a phrase inventory; each source phrase has a target
phrase and an internal word alignment; target phrases are locally reordered; a fraction of the
links is dropped so that unaligned words exist on both sides (this exercises the tight-phrase
logic of the extractor).  Everything is numpy-vectorised so that the 1 M sentence-pair config is
generated in seconds.

Two views of the same data are offered:
  * raw arrays (``SynthCorpus``) -> ``text_layout`` gives the int layouts the reference's loaders
    produce from text (``Start.cu:240-380``, ``ExtractPair.cu:2639-2739``);
  * text files (``write_text``) in the six-argument ``strmatchcuda`` format, for the CLI / the
    reference binary.
Token ids follow the reference rule: id = 2 + rank of first appearance (``Start.cu:288``).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

__all__ = ["SynthCorpus", "generate", "generate_queries", "query_ids", "text_layout", "write_text", "first_appearance_ids"]


@dataclass
class SynthCorpus:
    # word "names" are integer ranks in a Zipfian vocabulary; the strings are f"s{rank}" / f"t{rank}"
    src_words: np.ndarray   # int32 [Ns]  concatenated sentences (no EOS)
    src_off: np.ndarray     # int64 [S+1]
    tgt_words: np.ndarray   # int32 [Nt]
    tgt_off: np.ndarray     # int64 [S+1]
    link_sent: np.ndarray   # int32 [L]   sentence of each alignment link
    link_s: np.ndarray      # int32 [L]   source position inside the sentence
    link_t: np.ndarray      # int32 [L]   target position inside the sentence
    qry_words: np.ndarray   # int32 [T]   query sentences (same word-name space as src_words; -1 = OOV)
    qry_off: np.ndarray     # int64 [Q+1]
    v_src: int
    v_tgt: int
    meta: dict = field(default_factory=dict)

    @property
    def n_sent(self) -> int:
        return len(self.src_off) - 1

    @property
    def n_qry(self) -> int:
        return len(self.qry_off) - 1


def _zipf_sampler(rng: np.random.Generator, v: int, s: float = 1.0):
    w = 1.0 / np.power(np.arange(1, v + 1, dtype=np.float64), s)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]

    def draw(size):
        return np.minimum(np.searchsorted(cdf, rng.random(size), side="left"), v - 1).astype(np.int32)

    return draw


def _ragged_arange(lens: np.ndarray) -> np.ndarray:
    """concatenate(arange(l) for l in lens), vectorised."""
    lens = np.asarray(lens, dtype=np.int64)
    tot = int(lens.sum())
    starts = np.cumsum(lens) - lens
    return np.arange(tot, dtype=np.int64) - np.repeat(starts, lens)


def _sentences(rng, n_sent, inv, mean_len, sd_len, min_len, max_len, swap_p):
    """Build n_sent sentence pairs from the phrase inventory `inv`."""
    draw_phrase = inv["draw"]
    p_slen, p_tlen = inv["slen"], inv["tlen"]
    want = np.clip(np.rint(rng.normal(mean_len, sd_len, n_sent)), min_len, max_len).astype(np.int64)
    bounds = np.cumsum(want)
    total = int(bounds[-1])
    # draw enough phrase instances to cover `total` source tokens
    m_est = int(total / max(1.0, float(np.mean(p_slen[draw_phrase(4096)]))) * 1.1) + 64
    ph = draw_phrase(m_est)
    cs = np.cumsum(p_slen[ph].astype(np.int64))
    while cs[-1] < total:
        more = draw_phrase(m_est // 4 + 64)
        ph = np.concatenate([ph, more])
        cs = np.cumsum(p_slen[ph].astype(np.int64))
    m = int(np.searchsorted(cs, total, side="left")) + 1
    ph, cs = ph[:m], cs[:m]
    starts = cs - p_slen[ph]
    sent_of = np.minimum(np.searchsorted(bounds, starts, side="right"), n_sent - 1).astype(np.int64)
    # guarantee every sentence owns at least one phrase: drop empty sentences (rare, tiny `want`)
    uniq, first_idx, counts = np.unique(sent_of, return_index=True, return_counts=True)
    remap = -np.ones(n_sent, dtype=np.int64)
    remap[uniq] = np.arange(len(uniq))
    sent_of = remap[sent_of]
    n_sent = len(uniq)
    slen = p_slen[ph].astype(np.int64)
    tlen = p_tlen[ph].astype(np.int64)
    # position of the phrase inside its sentence (phrase index)
    pidx = np.arange(m) - np.repeat(first_idx, counts)
    # local reordering on the target side: swap non-overlapping adjacent phrase pairs
    order = pidx.copy()
    cand = (rng.random(m) < swap_p)
    cand[:-1] &= sent_of[:-1] == sent_of[1:]
    cand[-1] = False
    cand[1:] &= ~cand[:-1]          # no overlap (greedy left to right on a Bernoulli mask)
    idx = np.nonzero(cand)[0]
    order[idx] += 1
    order[idx + 1] -= 1
    # source offsets inside the sentence
    sent_src_start = np.concatenate([[0], np.cumsum(np.bincount(sent_of, weights=slen, minlength=n_sent).astype(np.int64))])
    src_off_in = starts - starts[np.repeat(first_idx, counts)]
    # target offsets: sort phrases of each sentence by target order, cumulative tlen
    key = sent_of * (m + 1) + order
    perm = np.argsort(key, kind="stable")
    tl_sorted = tlen[perm]
    ctl = np.cumsum(tl_sorted) - tl_sorted
    sent_tgt_len = np.bincount(sent_of, weights=tlen, minlength=n_sent).astype(np.int64)
    sent_tgt_start = np.concatenate([[0], np.cumsum(sent_tgt_len)])
    tgt_off_in = np.empty(m, dtype=np.int64)
    tgt_off_in[perm] = ctl - sent_tgt_start[sent_of[perm]]
    # tokens
    src_tok_idx = np.repeat(inv["sstart"][ph], slen) + _ragged_arange(slen)
    src_words = inv["swords"][src_tok_idx]
    tgt_words = np.empty(int(sent_tgt_start[-1]), dtype=np.int32)
    tpos = np.repeat(sent_tgt_start[sent_of] + tgt_off_in, tlen) + _ragged_arange(tlen)
    tgt_words[tpos] = inv["twords"][np.repeat(inv["tstart"][ph], tlen) + _ragged_arange(tlen)]
    # links
    nl = inv["nlinks"][ph].astype(np.int64)
    lidx = np.repeat(inv["lstart"][ph], nl) + _ragged_arange(nl)
    link_sent = np.repeat(sent_of, nl).astype(np.int32)
    link_s = (np.repeat(src_off_in, nl) + inv["link_s"][lidx]).astype(np.int32)
    link_t = (np.repeat(tgt_off_in, nl) + inv["link_t"][lidx]).astype(np.int32)
    return dict(src_words=src_words.astype(np.int32), src_off=sent_src_start, tgt_words=tgt_words,
                tgt_off=sent_tgt_start, link_sent=link_sent, link_s=link_s, link_t=link_t, n_sent=n_sent)


def _inventory(rng, n_phrases, v_src, v_tgt, zipf_s):
    draw_s = _zipf_sampler(rng, v_src, zipf_s)
    draw_t = _zipf_sampler(rng, v_tgt, zipf_s)
    slen = rng.choice([1, 2, 3, 4], size=n_phrases, p=[0.45, 0.30, 0.15, 0.10]).astype(np.int32)
    tlen = np.clip(slen + rng.choice([-1, 0, 1], size=n_phrases, p=[0.2, 0.6, 0.2]), 1, 4).astype(np.int32)
    sstart = np.cumsum(slen, dtype=np.int64) - slen
    tstart = np.cumsum(tlen, dtype=np.int64) - tlen
    swords = draw_s(int(slen.sum()))
    twords = draw_t(int(tlen.sum()))
    # links: every source word i -> round(i*tlen/slen), every target word j -> round(j*slen/tlen)
    a = np.repeat(np.arange(n_phrases), slen)
    ia = _ragged_arange(slen)
    ja = np.minimum((ia * tlen[a] + slen[a] // 2) // slen[a], tlen[a] - 1)
    b = np.repeat(np.arange(n_phrases), tlen)
    jb = _ragged_arange(tlen)
    ib = np.minimum((jb * slen[b] + tlen[b] // 2) // tlen[b], slen[b] - 1)
    p_all = np.concatenate([a, b])
    s_all = np.concatenate([ia, ib])
    t_all = np.concatenate([ja, jb])
    code = (p_all * 8 + s_all) * 8 + t_all
    code = np.unique(code)
    p_u = code // 64
    link_s = (code // 8) % 8
    link_t = code % 8
    nlinks = np.bincount(p_u, minlength=n_phrases).astype(np.int32)
    lstart = np.cumsum(nlinks, dtype=np.int64) - nlinks
    pz = _zipf_sampler(rng, n_phrases, zipf_s)
    return dict(draw=pz, slen=slen, tlen=tlen, sstart=sstart, tstart=tstart, swords=swords, twords=twords,
                nlinks=nlinks, lstart=lstart, link_s=link_s.astype(np.int32), link_t=link_t.astype(np.int32))


def generate(n_sent: int, n_qry: int, v_src: int = 50000, v_tgt: int = 50000, *, seed: int = 1234,
             qry_seed: int = 4321, n_phrases: int | None = None, zipf_s: float = 1.0, mean_len: float = 25.0,
             sd_len: float = 8.0, min_len: int = 3, max_len: int = 80, swap_p: float = 0.2,
             drop_p: float = 0.10, oov_p: float = 0.0, qry_mean_len: float | None = None) -> SynthCorpus:
    rng = np.random.default_rng(seed)
    if n_phrases is None:
        n_phrases = max(2000, min(2_000_000, n_sent // 2))
    inv = _inventory(rng, n_phrases, v_src, v_tgt, zipf_s)
    c = _sentences(rng, n_sent, inv, mean_len, sd_len, min_len, max_len, swap_p)
    keep = rng.random(len(c["link_s"])) >= drop_p
    for k in ("link_sent", "link_s", "link_t"):
        c[k] = c[k][keep]
    qrng = np.random.default_rng(qry_seed)
    inv_q = dict(inv)
    inv_q["draw"] = _zipf_sampler(qrng, n_phrases, zipf_s)
    q = _sentences(qrng, n_qry, inv_q, qry_mean_len or mean_len, sd_len, min_len, max_len, 0.0)
    qw = q["src_words"].copy()
    if oov_p > 0:
        qw[qrng.random(len(qw)) < oov_p] = -1
    return SynthCorpus(src_words=c["src_words"], src_off=c["src_off"], tgt_words=c["tgt_words"], tgt_off=c["tgt_off"],
                       link_sent=c["link_sent"], link_s=c["link_s"], link_t=c["link_t"], qry_words=qw,
                       qry_off=q["src_off"], v_src=v_src, v_tgt=v_tgt,
                       meta=dict(seed=seed, qry_seed=qry_seed, n_phrases=n_phrases, zipf_s=zipf_s, drop_p=drop_p,
                                 swap_p=swap_p, oov_p=oov_p))


def generate_queries(n_sent: int, n_qry: int, v_src: int = 50000, v_tgt: int = 50000, *, seed: int = 1234, qry_seed: int = 4321,
                     n_phrases: int | None = None, zipf_s: float = 1.0, mean_len: float = 25.0, sd_len: float = 8.0, min_len: int = 3,
                     max_len: int = 80, oov_p: float = 0.0, qry_mean_len: float | None = None):
    """Another query set for the corpus generate(n_sent, ..., seed=seed) makes, without re-generating the corpus: the phrase
    inventory is the first thing drawn from `seed`, so it is rebuilt exactly; the queries come from `qry_seed` alone.
    generate(..., qry_seed=s).qry_words == generate_queries(..., qry_seed=s)[0].  Returns (qry_words, qry_off)."""
    rng = np.random.default_rng(seed)
    if n_phrases is None:
        n_phrases = max(2000, min(2_000_000, n_sent // 2))
    inv = _inventory(rng, n_phrases, v_src, v_tgt, zipf_s)
    qrng = np.random.default_rng(qry_seed)
    inv["draw"] = _zipf_sampler(qrng, n_phrases, zipf_s)
    q = _sentences(qrng, n_qry, inv, qry_mean_len or mean_len, sd_len, min_len, max_len, 0.0)
    qw = q["src_words"].copy()
    if oov_p > 0:
        qw[qrng.random(len(qw)) < oov_p] = -1
    return qw, q["src_off"]


def query_ids(src_names: np.ndarray, qry_words: np.ndarray) -> np.ndarray:
    """Query word names -> ids of the source vocabulary (id = 2 + first-appearance rank, Start.cu:288), -1 = OOV (Start.cu:97)."""
    q_ids = np.full(len(qry_words), -1, dtype=np.int32)
    srt = np.argsort(src_names, kind="stable")
    sn_sorted = src_names[srt]
    k = np.minimum(np.searchsorted(sn_sorted, qry_words), len(sn_sorted) - 1)
    hit = (sn_sorted[k] == qry_words) & (qry_words >= 0)
    q_ids[hit] = (srt[k[hit]] + 2).astype(np.int32)
    return q_ids


def first_appearance_ids(words: np.ndarray):
    """word-name -> id = 2 + first-appearance rank (Start.cu:182,288).  Returns (ids, names_by_id)."""
    words = np.asarray(words)
    n, v = len(words), (int(words.max()) + 1 if len(words) else 0)
    # first occurrence of every word name in O(N): scatter positions in reverse order, the last write (= smallest position) wins
    first = np.full(v, n, dtype=np.int64)
    first[words[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)
    uniq = np.nonzero(first < n)[0]
    order = np.argsort(first[uniq], kind="stable")
    names_by_rank = uniq[order]
    rank_of = np.zeros(v, dtype=np.int32)
    rank_of[names_by_rank] = np.arange(len(uniq), dtype=np.int32)
    ids = rank_of[words] + 2
    return ids.astype(np.int32), names_by_rank.astype(np.int64)


def _layout_side(words, off, ids):
    """sentence tokens + EOS(1) after each sentence, then 1, last, 0,0,0  (Start.cu:306-327,354)."""
    n_sent = len(off) - 1
    lens = np.diff(off)
    n = int(off[-1]) + n_sent + 2
    buf = np.zeros(n + 3, dtype=np.int32)
    sent_start = off[:-1] + np.arange(n_sent)           # start of sentence k in the layout
    pos = np.repeat(sent_start, lens) + _ragged_arange(lens)
    buf[pos] = ids
    buf[sent_start + lens] = 1
    buf[n - 2] = 1
    buf[n - 1] = int(ids.max()) + 1 if len(ids) else 2
    sentenceind = np.concatenate([sent_start, [n - 2]]).astype(np.int64)
    return buf, n, sentenceind, pos


def text_layout(c: SynthCorpus) -> dict:
    """The integer arrays the reference's loaders build from the text files.

    str/tgt : int32 [n+3]/[m+3]   Start.cu:240-380 / :142-238
    P       : uint8 [n]           position in sentence (Start.cu:300)
    L_tar/R_tar : uint8 [m]       min/max aligned source index per target token, 255 = unaligned
    RLP     : uint32 [n]          (L<<24)|(R<<16)|(P<<8); at source EOS k: target offset of sentence k+1
                                  (ExtractPair.cu:2698-2728)
    lex_*   : lexical table (f id, e id, v1, v2) with NULL = -1 rows (ExtractPair.cu:2463-2512)
    qry_tok/qry_off : query ids in the source vocabulary, -1 = OOV (Start.cu:97)
    """
    s_ids, s_names = first_appearance_ids(c.src_words)
    t_ids, t_names = first_appearance_ids(c.tgt_words)
    s_buf, n, s_sent, s_pos = _layout_side(c.src_words, c.src_off, s_ids)
    t_buf, m, t_sent, t_pos = _layout_side(c.tgt_words, c.tgt_off, t_ids)
    lens = np.diff(c.src_off)
    P = np.zeros(n, dtype=np.uint8)
    P[s_pos] = (_ragged_arange(lens) & 0xFF).astype(np.uint8)
    L_src = np.full(n, 255, dtype=np.int64)
    R_src = np.full(n, -1, dtype=np.int64)
    L_tar = np.full(m, 255, dtype=np.int64)
    R_tar = np.full(m, -1, dtype=np.int64)
    si = s_sent[c.link_sent] + c.link_s
    ti = t_sent[c.link_sent] + c.link_t
    _min_max_by_key(si, c.link_t, L_src, R_src)
    _min_max_by_key(ti, c.link_s, L_tar, R_tar)
    R_src[R_src < 0] = 255
    R_tar[R_tar < 0] = 255
    RLP = ((L_src.astype(np.uint32) << 24) | (R_src.astype(np.uint32) << 16) | (P.astype(np.uint32) << 8))
    eos = s_sent[1:] - 1                      # EOS of sentence k-1 sits at sentenceind[k]-1
    RLP[eos] = t_sent[1:].astype(np.uint32)
    RLP[n - 1] = 0                            # never initialised by the reference (loop stops at toklen-1)
    # lexical table: relative frequencies over the alignment links + NULL rows
    f = s_buf[si].astype(np.int64)
    e = t_buf[ti].astype(np.int64)
    pair = f * (int(t_ids.max()) + 3) + e
    up, cnt = np.unique(pair, return_counts=True)
    pf = up // (int(t_ids.max()) + 3)
    pe = up % (int(t_ids.max()) + 3)
    cf = np.bincount(f, minlength=int(s_ids.max()) + 1).astype(np.float64)
    ce = np.bincount(e, minlength=int(t_ids.max()) + 1).astype(np.float64)
    v1 = cnt / cf[pf]
    v2 = cnt / ce[pe]
    all_f = np.arange(2, int(s_ids.max()) + 1, dtype=np.int64)
    all_e = np.arange(2, int(t_ids.max()) + 1, dtype=np.int64)
    # NULL rows: a small, word-dependent probability so that max() decisions are exercised
    nf1 = 0.001 + 0.05 / (1.0 + (all_f % 17))
    nf2 = 0.002 + 0.04 / (1.0 + (all_f % 13))
    ne1 = 0.0015 + 0.03 / (1.0 + (all_e % 11))
    ne2 = 0.0005 + 0.06 / (1.0 + (all_e % 19))
    lex_f = np.concatenate([[-1], all_f, np.full(len(all_e), -1), pf]).astype(np.int32)
    lex_e = np.concatenate([[-1], np.full(len(all_f), -1), all_e, pe]).astype(np.int32)
    lex_v1 = np.concatenate([[1.0], nf1, ne1, v1])
    lex_v2 = np.concatenate([[1.0], nf2, ne2, v2])
    # values are what a "%.6g" text round trip gives (so text and array paths agree bit-for-bit)
    lex_v1 = np.array([float("%.6g" % x) for x in lex_v1], dtype=np.float32) if len(lex_v1) < 200000 else _round6(lex_v1)
    lex_v2 = np.array([float("%.6g" % x) for x in lex_v2], dtype=np.float32) if len(lex_v2) < 200000 else _round6(lex_v2)
    q_ids = query_ids(s_names, c.qry_words)
    return dict(str=s_buf, n=n, tgt=t_buf, m=m, P=P, L_tar=L_tar.astype(np.uint8), R_tar=R_tar.astype(np.uint8),
                RLP=RLP.astype(np.uint32), src_sentenceind=s_sent, tgt_sentenceind=t_sent,
                src_names=s_names, tgt_names=t_names, src_last=int(s_buf[n - 1]), tgt_last=int(t_buf[m - 1]),
                lex_f=lex_f, lex_e=lex_e, lex_v1=lex_v1, lex_v2=lex_v2,
                qry_tok=q_ids, qry_off=c.qry_off.astype(np.int32))


def _min_max_by_key(key, val, mn_out, mx_out):
    """mn_out[k] = min(val[key == k]), mx_out[k] = max(...) for the keys that occur (np.minimum.at / np.maximum.at in one
    sort instead of two scattered read-modify-write passes).  val < 256."""
    if len(key) == 0:
        return
    packed = key.astype(np.int64) * 256 + val.astype(np.int64)
    if np.any(packed[1:] < packed[:-1]):
        packed = np.sort(packed)
    k = packed >> 8
    first = np.concatenate([[True], k[1:] != k[:-1]])
    last = np.concatenate([first[1:], [True]])
    mn_out[k[first]] = packed[first] & 255
    mx_out[k[last]] = packed[last] & 255


def _round6(x: np.ndarray) -> np.ndarray:
    """vectorised equivalent of float('%.6g' % x) for positive x (6 significant digits)."""
    x = np.asarray(x, dtype=np.float64)
    e = np.floor(np.log10(x))
    s = np.power(10.0, 5 - e)
    r = np.rint(x * s) / s
    return r.astype(np.float32)


def write_text(c: SynthCorpus, outdir: str, prefix: str = "corpus") -> dict:
    """Write source / query / target / alignment / lex files in the strmatchcuda input format."""
    os.makedirs(outdir, exist_ok=True)
    lay = text_layout(c)
    paths = {k: os.path.join(outdir, f"{prefix}.{k}") for k in ("f", "q", "e", "a", "lex")}

    def dump_lines(path, strs, off):
        """One line per [off[k], off[k+1]) run of strs, blank-separated; empty runs give empty lines.  Vectorised: every
        string gets its separator appended (a blank, or as many newlines as lines end after it) and the lot is joined once."""
        off = np.asarray(off, dtype=np.int64)
        n_lines, n = len(off) - 1, len(strs)
        with open(path, "w") as fh:
            if n == 0:
                fh.write("\n" * n_lines)
                return
            line_of = np.searchsorted(off, np.arange(n), side="right") - 1          # line of every string
            ends_here = np.bincount(line_of, minlength=n_lines)                     # strings per line
            last = off[1:][ends_here > 0] - 1                                       # last string of every non-empty line
            nonempty = np.nonzero(ends_here > 0)[0]
            nxt = np.append(nonempty[1:], n_lines)                                  # next non-empty line (or the end)
            sep = np.full(n, " ", dtype=object)
            sep[last] = ["\n" * int(k) for k in (nxt - nonempty)]
            fh.write("\n" * int(nonempty[0]))
            step = 1 << 22
            for a in range(0, n, step):
                fh.write("".join(map(str.__add__, strs[a:a + step].tolist(), sep[a:a + step].tolist())))

    def names(tag, words):                                     # "<tag><word>" through a table of the distinct words
        top = int(words.max()) + 1 if len(words) else 0
        return np.char.add(tag, np.arange(top).astype(str))[words]

    dump_lines(paths["f"], names("s", c.src_words), c.src_off)
    dump_lines(paths["e"], names("t", c.tgt_words), c.tgt_off)
    dump_lines(paths["q"], np.where(c.qry_words >= 0, names("s", np.maximum(c.qry_words, 0)), "OOV" + "x"), c.qry_off)
    order = np.argsort(c.link_sent, kind="stable")
    ls, lsrc, ltgt = c.link_sent[order], c.link_s[order], c.link_t[order]
    bounds = np.searchsorted(ls, np.arange(c.n_sent + 1))
    dump_lines(paths["a"], np.char.add(names("", lsrc), names("-", ltgt)), bounds)
    sname = lambda i: "NULL" if i < 0 else "s%d" % lay["src_names"][i - 2]
    tname = lambda i: "NULL" if i < 0 else "t%d" % lay["tgt_names"][i - 2]
    with open(paths["lex"], "w") as fh:
        for f, e, a, b in zip(lay["lex_f"], lay["lex_e"], lay["lex_v1"], lay["lex_v2"]):
            fh.write("%s %s %.6g %.6g\n" % (sname(int(f)), tname(int(e)), float(a), float(b)))
    return paths
