"""Python host-side mirror of the reference's start() sequence (Start.cu:488-629) over the C ABI:

    ex = GrammarExtractor(device=0)
    ex.build_index(layout)        # initRefSet/initRefTargetSet/initAlignment output -> GPU SA + index
    ex.load_lex(f, e, v1, v2)     # initWordPossibilityIntKey
    res = ex.extract(qry_tok, qry_off)   # suffixArraySearch + ExtractPairs_Large_Data_Gappy
    res.grammar_lines(q, src_names, tgt_names)   # print_query_GPU_Gappy, one query

Everything numeric happens in libcgx_b200.so on the GPU; this module only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import BatchInfo, IndexArrays, IndexInfo, Result, RULE_DTYPE, RULE_WIRE_DTYPE


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _np(ptr, n, dtype=np.int32):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


def decode_rules(wire, updown, idinfo):
    """cgx_rule_t wire records (16 B, include/cgx_b200.h) -> RULE_DTYPE rows with the converted id, f and fs filled in:
    the id of a rule is the [first, last] range it sits in; f and fs travel once per id in idinfo."""
    n = len(wire)
    out = np.zeros(n, dtype=RULE_DTYPE)
    if n == 0:
        return out
    span = wire["span"]
    out["tgt_start"], out["mlfe"], out["mlef"] = wire["tgt_start"], wire["mlfe"], wire["mlef"]
    out["end"] = span & 15
    for name, sh in (("gap1", 4), ("gap1_1", 8), ("gap2", 12), ("gap2_1", 16)):
        g = (span >> sh) & 15
        out[name] = np.where(g == 15, 255, g)
    out["pc"] = (span >> 20) & 511
    ids = np.nonzero(updown[:, 0] >= 0)[0]                 # ids with rules, ascending = rule order
    lo, hi = updown[ids, 0], updown[ids, 1]
    starts = np.zeros(n, dtype=np.int64)
    starts[lo] = 1
    seg = np.cumsum(starts) - 1
    assert np.array_equal(hi - lo + 1, np.bincount(seg, minlength=len(ids))), "updown ranges do not tile the rules"
    rid = ids[seg].astype(np.int32)
    out["id"] = rid
    out["f"] = idinfo[rid] & 0x1FF                        # CGX_ID_F / CGX_ID_FS of include/cgx_b200.h
    out["fs"] = (idinfo[rid] >> 9) & 0x1FF
    return out


class BatchResult:
    """Host copy of one batch's results (cgx_result_t)."""

    def __init__(self, r: Result, info: BatchInfo, qry_tok, qry_off):
        self.Q, self.T, self.G, self.D1, self.D2 = r.Q, r.T, r.G, r.D1, r.D2
        self.info = info.as_dict()
        self.qry_tok = np.asarray(qry_tok, dtype=np.int32)
        self.qry_off = np.asarray(qry_off, dtype=np.int32)
        self.phrase_id = _np(r.phrase_id, self.T * 5).reshape(self.T, 5)
        self.phrases = _np(r.phrases, self.G * 4).reshape(self.G, 4)
        self.pat1 = _np(r.pat1, self.D1 * 4).reshape(self.D1, 4)          # {a_pos, ls, b_pos, le}
        self.pat2 = _np(r.pat2, self.D2 * 2).reshape(self.D2, 2)          # {pat1, ctok}
        self.q1_off = _np(r.q1_off, self.Q + 1)
        self.q1_ids = _np(r.q1_ids, int(self.q1_off[-1]) if self.Q else 0)
        self.q2_off = _np(r.q2_off, self.Q + 1)
        self.q2_ids = _np(r.q2_ids, int(self.q2_off[-1]) if self.Q else 0)
        self.rules, self.updown, self.idinfo = [], [], []
        for k in range(3):
            n = r.n_rules[k]
            first = _np(r.first[k], r.n_ids[k])
            ii = _np(r.idinfo[k], r.n_ids[k], dtype=np.uint32) if r.n_ids[k] else np.zeros(0, dtype=np.uint32)
            # [first, last] rule of every id (-1, -1 when it has none): first[id] and CGX_ID_RULES(idinfo[id])
            ud = np.stack([first, np.where(first >= 0, first + ((ii >> 18) & 0x1FF).astype(np.int32) - 1, -1)], axis=1).astype(np.int32)
            self.updown.append(ud)
            self.idinfo.append(ii)
            if n:
                buf = (C.c_char * (n * RULE_WIRE_DTYPE.itemsize)).from_address(r.rules[k])
                self.rules.append(decode_rules(np.frombuffer(buf, dtype=RULE_WIRE_DTYPE), ud, ii))
            else:
                self.rules.append(np.zeros(0, dtype=RULE_DTYPE))

    # ---- features exactly as the reference's host code computes them (ExtractPair.c:653-655, :641) ----
    @staticmethod
    def features(rule):
        pc, fs = int(rule["pc"]), int(rule["fs"])
        aa = -np.log10(np.float32(pc) / np.float32(fs), dtype=np.float32)
        score = np.float32(math.log10(1 + fs))
        bb = np.float32(math.log10(1 + pc))
        return float(aa), float(score), float(bb)

    # ---- grammar text of one query (PrintResults.c:407-577) ----
    def source_tokens(self, kind, cid, layout_str):
        G, D1, D2 = self.G, self.D1, self.D2

        def phrase(g):
            s, l = int(self.phrases[g, 3]), int(self.phrases[g, 2])
            return [int(x) for x in layout_str[s:s + l]]

        def p1(d, gap):
            a, ls, b, le = (int(x) for x in self.pat1[d, :4])
            return [int(x) for x in layout_str[a:a + ls]] + [gap] + [int(x) for x in layout_str[b:b + le]]

        if kind == 0:
            return phrase(cid)
        if kind == 1:
            if cid < G:
                return [-1] + phrase(cid)
            if cid < 2 * G:
                return phrase(cid - G) + [-1]
            return p1(cid - 2 * G, -1)
        if cid < G:
            return [-1] + phrase(cid) + [-2]
        if cid < G + D2:
            d2 = cid - G
            return p1(int(self.pat2[d2, 0]), -1) + [-2, int(self.pat2[d2, 1])]
        if cid < G + D2 + D1:
            return [-1] + p1(cid - G - D2, -2)
        return p1(cid - G - D2 - D1, -1) + [-2]

    def target_symbols(self, rule, layout_tgt):
        ts, end = int(rule["tgt_start"]), int(rule["end"])
        g1, g1e, g2, g2e = int(rule["gap1"]), int(rule["gap1_1"]), int(rule["gap2"]), int(rule["gap2_1"])
        out, j = [], 0
        while j <= end:
            if g1 != 255 and g1 <= j <= g1e:
                out.append(-1)
                j = g1e + 1
            elif g2 != 255 and g2 <= j <= g2e:
                out.append(-2)
                j = g2e + 1
            else:
                out.append(int(layout_tgt[ts + j]))
                j += 1
        return out

    def query_groups(self, q):
        """(kind, converted id) groups of query q in the reference's print order."""
        G, D1, D2 = self.G, self.D1, self.D2
        seen, order = set(), []
        for t in range(int(self.qry_off[q]), int(self.qry_off[q + 1])):
            for m in range(5):
                g = int(self.phrase_id[t, m])
                if g >= 0 and g not in seen:
                    seen.add(g)
                    order.append(g)
        groups = []
        for g in order:
            groups += [(1, g + G), (1, g), (2, g), (0, g)]
        for d in self.q1_ids[self.q1_off[q]:self.q1_off[q + 1]]:
            d = int(d)
            groups += [(1, 2 * G + d), (2, G + D2 + d), (2, G + D2 + D1 + d)]
        for d in self.q2_ids[self.q2_off[q]:self.q2_off[q + 1]]:
            groups.append((2, G + int(d)))
        return groups

    def grammar_lines(self, q, layout, src_name=lambda i: "s%d" % i, tgt_name=lambda i: "t%d" % i):
        sn, tn = layout["src_names"], layout["tgt_names"]

        def sw(t):
            return "[X,1]" if t == -1 else "[X,2]" if t == -2 else src_name(int(sn[t - 2]))

        def tw(t):
            return "[X,1]" if t == -1 else "[X,2]" if t == -2 else tgt_name(int(tn[t - 2]))

        lines = []
        for kind, cid in self.query_groups(q):
            lo, hi = (int(x) for x in self.updown[kind][cid])
            if lo < 0:
                continue
            src = " ".join(sw(t) for t in self.source_tokens(kind, cid, layout["str"]))
            for r in self.rules[kind][lo:hi + 1]:
                aa, score, bb = self.features(r)
                tgt = " ".join(tw(t) for t in self.target_symbols(r, layout["tgt"]))
                lines.append("[X] ||| %s ||| %s ||| EgivenFCoherent=%f SampleCountF=%f CountEF=%f MaxLexFgivenE=%f MaxLexEgivenF=%f "
                             "IsSingletonF=%d IsSingletonFE=%d" % (src, tgt, aa, score, bb, float(r["mlfe"]), float(r["mlef"]),
                                                                  int(r["f"]) == 1, int(r["pc"]) == 1))
        return lines


class GrammarExtractor:
    def __init__(self, device: int = 0):
        self.L = _lib.load()
        h = C.c_void_p()
        rc = self.L.cgx_create(device, C.byref(h))
        if rc:
            raise RuntimeError("cgx_create failed: " + (self.L.cgx_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def _check(self, rc, what):
        if rc:
            raise RuntimeError(f"{what} failed: " + (self.L.cgx_last_error(self.h) or b"").decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.cgx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- index ----
    def build_index(self, layout):
        s = np.ascontiguousarray(layout["str"], dtype=np.int32)
        t = np.ascontiguousarray(layout["tgt"], dtype=np.int32)
        if layout.get("wide"):        # 16-bit alignment fields: a sentence of 255 tokens or more (include/cgx_b200.h cgx_index_build_wide)
            rlp = np.ascontiguousarray(layout["RLP"], dtype=np.uint64)
            lt = np.ascontiguousarray(layout["L_tar"], dtype=np.uint16)
            rt = np.ascontiguousarray(layout["R_tar"], dtype=np.uint16)
            self._check(self.L.cgx_index_build_wide(self.h, _p(s, C.c_int32), int(layout["n"]), _p(t, C.c_int32), int(layout["m"]), _p(rlp, C.c_uint64),
                                                    _p(lt, C.c_uint16), _p(rt, C.c_uint16)), "cgx_index_build_wide")
        else:
            rlp = np.ascontiguousarray(layout["RLP"], dtype=np.uint32)
            lt = np.ascontiguousarray(layout["L_tar"], dtype=np.uint8)
            rt = np.ascontiguousarray(layout["R_tar"], dtype=np.uint8)
            self._check(self.L.cgx_index_build(self.h, _p(s, C.c_int32), int(layout["n"]), _p(t, C.c_int32), int(layout["m"]), _p(rlp, C.c_uint32),
                                               _p(lt, C.c_uint8), _p(rt, C.c_uint8)), "cgx_index_build")
        if "lex_f" in layout:
            self.load_lex(layout["lex_f"], layout["lex_e"], layout["lex_v1"], layout["lex_v2"])
        return self.index_info()

    def load_lex(self, f, e, v1, v2):
        f = np.ascontiguousarray(f, dtype=np.int32)
        e = np.ascontiguousarray(e, dtype=np.int32)
        v1 = np.ascontiguousarray(v1, dtype=np.float32)
        v2 = np.ascontiguousarray(v2, dtype=np.float32)
        self._check(self.L.cgx_lex_load(self.h, _p(f, C.c_int32), _p(e, C.c_int32), _p(v1, C.c_float), _p(v2, C.c_float), len(f)), "cgx_lex_load")

    def save_index(self, path):
        self._check(self.L.cgx_index_save(self.h, str(path).encode()), "cgx_index_save")

    def load_index(self, path):
        self._check(self.L.cgx_index_load(self.h, str(path).encode()), "cgx_index_load")
        return self.index_info()

    def index_info(self):
        info = IndexInfo()
        self.L.cgx_index_info(self.h, C.byref(info))
        return {k: getattr(info, k) for k, _ in info._fields_}

    def suffix_array(self):
        n = self.index_info()["n"]
        out = np.empty(n, dtype=np.int32)
        self._check(self.L.cgx_index_copy_sa(self.h, _p(out, C.c_int32)), "cgx_index_copy_sa")
        return out

    def occurrence_list(self, m):
        n = self.index_info()["n"]
        out = np.empty(n, dtype=np.int32)
        self._check(self.L.cgx_index_copy_inv(self.h, m, _p(out, C.c_int32)), "cgx_index_copy_inv")
        return out

    def frequent_tokens(self):
        out = np.empty(100, dtype=np.int32)
        self._check(self.L.cgx_index_copy_frequent(self.h, _p(out, C.c_int32)), "cgx_index_copy_frequent")
        return out

    def sa_build_dev(self, str_ptr: int, n: int, max_token: int, sa_ptr: int):
        rounds, ms = C.c_int32(), C.c_float()
        self._check(self.L.cgx_sa_build_dev(self.h, C.c_void_p(str_ptr), n, max_token, C.c_void_p(sa_ptr), C.byref(rounds), C.byref(ms)), "cgx_sa_build_dev")
        return rounds.value, ms.value

    def export_index(self) -> IndexArrays:
        a = IndexArrays()
        self._check(self.L.cgx_index_export(self.h, C.byref(a)), "cgx_index_export")
        return a

    def alloc_index(self, shape: IndexArrays) -> IndexArrays:
        a = IndexArrays()
        self._check(self.L.cgx_index_alloc(self.h, C.byref(shape), C.byref(a)), "cgx_index_alloc")
        return a

    def commit_index(self):
        self._check(self.L.cgx_index_commit(self.h), "cgx_index_commit")

    # ---- queries ----
    def extract(self, qry_tok, qry_off, fetch=True):
        qt = np.ascontiguousarray(qry_tok, dtype=np.int32)
        qo = np.ascontiguousarray(qry_off, dtype=np.int32)
        if len(qt) == 0:
            qt = np.zeros(1, dtype=np.int32)
        self._check(self.L.cgx_extract(self.h, _p(qt, C.c_int32), _p(qo, C.c_int32), len(qo) - 1), "cgx_extract")
        info = BatchInfo()
        self.L.cgx_batch_info(self.h, C.byref(info))
        if not fetch:
            return info.as_dict()
        r = Result()
        self.L.cgx_result(self.h, C.byref(r))
        return BatchResult(r, info, qry_tok, qry_off)

    def extract_begin(self, qry_tok, qry_off):
        """Pipelined batch (cgx_extract_begin): returns when the kernels are done, the result copy still travelling; it
        overlaps the next extract_begin.  Collect with result_at(age, ...)."""
        qt = np.ascontiguousarray(qry_tok, dtype=np.int32)
        qo = np.ascontiguousarray(qry_off, dtype=np.int32)
        if len(qt) == 0:
            qt = np.zeros(1, dtype=np.int32)
        self._check(self.L.cgx_extract_begin(self.h, _p(qt, C.c_int32), _p(qo, C.c_int32), len(qo) - 1), "cgx_extract_begin")
        info = BatchInfo()
        self.L.cgx_batch_info(self.h, C.byref(info))
        return info

    def result_at(self, age, qry_tok=None, qry_off=None, info=None, raw=False):
        """Results of the batch `age` begins ago (0 = latest), complete on the host when this returns.  raw=True
        returns the ctypes views (no copy), else a BatchResult (needs that batch's qry_tok / qry_off)."""
        r = Result()
        self._check(self.L.cgx_result_at(self.h, age, C.byref(r)), "cgx_result_at")
        if raw:
            return r
        return BatchResult(r, info if info is not None else BatchInfo(), qry_tok, qry_off)

    def extract_stream(self, qry_tok, qry_off, batch_queries=10000, on_batch=None):
        """A stream of query batches through the pipelined API, the way bin/strmatchcuda drives it: batches of
        `batch_queries`, a batch refused as too large (CGX_E_BATCH_TOO_LARGE) is cut in two.  on_batch(q0, q1, raw
        cgx_result_t views) is called for every finished batch, in order; returns the list of batch infos."""
        qt = np.ascontiguousarray(qry_tok, dtype=np.int32)
        qo = np.ascontiguousarray(qry_off, dtype=np.int32)
        Q = len(qo) - 1
        todo = [(a, min(Q, a + batch_queries)) for a in range(0, Q, batch_queries)][::-1]
        infos, prev = [], None
        self.stream_refusals = 0                                      # batches refused (and split) by this stream
        while todo:
            q0, q1 = todo.pop()
            fit = int(self.L.cgx_batch_advice(self.h, q1 - q0))       # from the hits per query of the last batch: at most one refusal per stream
            if fit < q1 - q0:
                todo.append((q0 + fit, q1))
                q1 = q0 + fit
            t = qt[qo[q0]:qo[q1]] if qo[q1] > qo[q0] else np.zeros(1, dtype=np.int32)
            o = np.ascontiguousarray(qo[q0:q1 + 1] - qo[q0])
            rc = self.L.cgx_extract_begin(self.h, _p(np.ascontiguousarray(t), C.c_int32), _p(o, C.c_int32), q1 - q0)
            if rc == 3 and q1 - q0 > 1:
                self.stream_refusals += 1
                mid = (q0 + q1) // 2
                todo += [(mid, q1), (q0, mid)]
                continue
            self._check(rc, "cgx_extract_begin")
            info = BatchInfo()
            self.L.cgx_batch_info(self.h, C.byref(info))
            d = info.as_dict()
            d["q0"], d["q1"] = q0, q1
            infos.append(d)
            if prev is not None and on_batch is not None:
                on_batch(prev[0], prev[1], self.result_at(1, raw=True))
            prev = (q0, q1)
        if prev is not None:
            r = self.result_at(0, raw=True)
            if on_batch is not None:
                on_batch(prev[0], prev[1], r)
        return infos

    def extract_dev(self, tok_ptr: int, off_ptr: int, t2q_ptr: int, Q: int, T: int):
        """Queries resident in HBM, results left in HBM (device-resident throughput)."""
        self._check(self.L.cgx_extract_dev(self.h, C.c_void_p(tok_ptr), C.c_void_p(off_ptr), C.c_void_p(t2q_ptr), Q, T), "cgx_extract_dev")
        info = BatchInfo()
        self.L.cgx_batch_info(self.h, C.byref(info))
        return info.as_dict()

    def profile(self, on=True):
        self.L.cgx_profile_enable(self.h, 1 if on else 0)

    def profile_report(self):
        import json
        return json.loads((self.L.cgx_profile_report(self.h) or b"{}").decode())

    def debug_fetch(self, what, cap, cols=None):
        out = np.empty(max(1, cap), dtype=np.int32)
        n = self.L.cgx_debug_fetch(self.h, what.encode(), _p(out, C.c_int32), cap)
        if n < 0:
            raise RuntimeError(f"cgx_debug_fetch({what}) -> {n}")
        out = out[:n].copy()
        return out.reshape(-1, cols) if cols else out
