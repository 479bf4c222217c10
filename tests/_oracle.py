"""ctypes binding of the CPU oracle (oracle/cgx_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this; the product (cgx_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libcgx_oracle.so")
WIDE_LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libcgx_oracle_wide.so")      # the same algorithm on 16-bit alignment fields
CLI_PATH = os.path.join(ORACLE_DIR, "_build", "cgx_oracle_cli")
REF_SA_PATH = os.path.join(ORACLE_DIR, "_ref", "libref_sa.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "strmatchcuda")
REF_DUMP_BIN = os.path.join(ORACLE_DIR, "_ref", "strmatchcuda_dump")


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return LIB_PATH


class Counts(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n", "m", "Q", "T", "G", "enu1", "D1", "hits1", "enu2", "D2", "hits2", "precomp_count", "n_ab",
        "n_1gap_contig", "n_2gap_contig", "n_axbxc", "n_axb", "n_2gap_from1", "lex_1gap", "lex_2gap", "lex_ab")]


RULE_DTYPE = np.dtype([("id", "<i4"), ("rec", "<i4", (6,)), ("f", "<i4"), ("fs", "<i4"), ("pc", "<i4"),
                       ("aa", "<f4"), ("score", "<f4"), ("bb", "<f4"), ("mlfe", "<f4"), ("mlef", "<f4")])

_libs = {}


def lib(wide=False):
    """The oracle library: the reference's 8-bit alignment fields, or (wide) the -DORC_WIDE build of the same source with 16-bit
    fields for corpora with sentences of 255 tokens and more."""
    if wide not in _libs:
        path = WIDE_LIB_PATH if wide else LIB_PATH
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        vp, i32p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float)
        rlp_p, lr_p = (C.POINTER(C.c_uint64), C.POINTER(C.c_uint16)) if wide else (C.POINTER(C.c_uint32), C.POINTER(C.c_uint8))
        L.orc_is_wide.restype = C.c_int
        assert bool(L.orc_is_wide()) == bool(wide)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [i32p, C.c_int32, i32p, C.c_int32, rlp_p, lr_p, lr_p, i32p, i32p, f32p, f32p, C.c_int32]
        L.orc_create_from_files.restype = vp
        L.orc_create_from_files.argtypes = [C.c_char_p] * 4
        L.orc_destroy.argtypes = [vp]
        L.orc_set_sa.argtypes = [vp, i32p]
        L.orc_build_sa.argtypes = [vp]
        L.orc_run.argtypes = [vp, i32p, i32p, C.c_int32]
        L.orc_run_query_file.argtypes = [vp, C.c_char_p]
        L.orc_write_grammars.argtypes = [vp, C.c_char_p]
        L.orc_get_counts.argtypes = [vp, C.POINTER(Counts)]
        for name in ("orc_sa", "orc_longest", "orc_blocks", "orc_onegap_patterns", "orc_onegap_hits", "orc_twogap_patterns",
                     "orc_twogap_hits", "orc_frequent", "orc_feature_missing", "orc_precomp_index", "orc_precomp_list"):
            getattr(L, name).restype = i32p
            getattr(L, name).argtypes = [vp]
        L.orc_interval.argtypes = [vp, C.c_int32, C.c_int32, i32p, i32p]
        L.orc_records.restype = C.c_int32
        L.orc_records.argtypes = [vp, C.c_int, C.POINTER(i32p)]
        L.orc_query_list.restype = C.c_int32
        L.orc_query_list.argtypes = [vp, C.c_int, C.c_int32, C.POINTER(i32p)]
        L.orc_rules.restype = C.c_int32
        L.orc_rules.argtypes = [vp, C.c_int, C.POINTER(vp)]
        _libs[wide] = L
    return _libs[wide]


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _arr(ptr, n, cols=None):
    if n == 0:
        return np.zeros((0, cols) if cols else (0,), dtype=np.int32)
    a = np.ctypeslib.as_array(ptr, shape=(n * (cols or 1),)).copy()
    return a.reshape(n, cols) if cols else a


class Oracle:
    def __init__(self, handle, wide=False):
        if not handle:
            raise RuntimeError("oracle construction failed")
        self.h = handle
        self.L = lib(wide)
        self.wide = wide

    @classmethod
    def from_layout(cls, lay, wide=None):
        """wide=None: as the layout says; True on an 8-bit layout widens it (the wide build must then agree with the narrow one)."""
        lay_wide = bool(lay.get("wide", False))
        wide = lay_wide if wide is None else bool(wide)
        L = lib(wide)
        s = np.ascontiguousarray(lay["str"], dtype=np.int32)
        t = np.ascontiguousarray(lay["tgt"], dtype=np.int32)
        rlp, lt, rt = np.asarray(lay["RLP"]), np.asarray(lay["L_tar"]), np.asarray(lay["R_tar"])
        if wide and not lay_wide:
            x = rlp.astype(np.uint64)
            Lf, Rf, Pf = (x >> 24) & 0xFF, (x >> 16) & 0xFF, (x >> 8) & 0xFF
            Lf[Lf == 255] = 65535
            Rf[Rf == 255] = 65535
            eos = s[: len(x)] < 2                       # the word at an EOS is the target sentence offset
            rlp = np.where(eos, x, (Lf << 48) | (Rf << 32) | (Pf << 16))
            lt = np.where(lt == 255, 65535, lt.astype(np.uint16))
            rt = np.where(rt == 255, 65535, rt.astype(np.uint16))
        elif lay_wide and not wide:
            raise ValueError("a 16-bit layout needs the wide oracle")
        rlp = np.ascontiguousarray(rlp, dtype=np.uint64 if wide else np.uint32)
        lt = np.ascontiguousarray(lt, dtype=np.uint16 if wide else np.uint8)
        rt = np.ascontiguousarray(rt, dtype=np.uint16 if wide else np.uint8)
        ct_rlp, ct_lr = (C.c_uint64, C.c_uint16) if wide else (C.c_uint32, C.c_uint8)
        lf = np.ascontiguousarray(lay["lex_f"], dtype=np.int32)
        le = np.ascontiguousarray(lay["lex_e"], dtype=np.int32)
        v1 = np.ascontiguousarray(lay["lex_v1"], dtype=np.float32)
        v2 = np.ascontiguousarray(lay["lex_v2"], dtype=np.float32)
        h = L.orc_create(_p(s, C.c_int32), int(lay["n"]), _p(t, C.c_int32), int(lay["m"]), _p(rlp, ct_rlp), _p(lt, ct_lr),
                         _p(rt, ct_lr), _p(lf, C.c_int32), _p(le, C.c_int32), _p(v1, C.c_float), _p(v2, C.c_float), len(lf))
        return cls(h, wide)

    @classmethod
    def from_files(cls, src, tgt, align, lex, wide=False):
        return cls(lib(wide).orc_create_from_files(src.encode(), tgt.encode(), align.encode(), lex.encode()), wide)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build_sa(self):
        self.L.orc_build_sa(self.h)

    def set_sa(self, sa):
        sa = np.ascontiguousarray(sa, dtype=np.int32)
        self.L.orc_set_sa(self.h, _p(sa, C.c_int32))

    def run(self, qry_tok, qry_off):
        qt = np.ascontiguousarray(qry_tok, dtype=np.int32)
        if len(qt) == 0:
            qt = np.zeros(1, dtype=np.int32)
        qo = np.ascontiguousarray(qry_off, dtype=np.int32)
        rc = self.L.orc_run(self.h, _p(qt, C.c_int32), _p(qo, C.c_int32), len(qo) - 1)
        if rc:
            raise RuntimeError(f"oracle run failed rc={rc}")
        return self.counts()

    def run_query_file(self, path):
        rc = self.L.orc_run_query_file(self.h, path.encode())
        if rc:
            raise RuntimeError(f"oracle run failed rc={rc}")
        return self.counts()

    def write_grammars(self, outdir):
        rc = self.L.orc_write_grammars(self.h, outdir.encode())
        if rc:
            raise RuntimeError(f"oracle write failed rc={rc}")

    def counts(self):
        c = Counts()
        self.L.orc_get_counts(self.h, C.byref(c))
        return c

    def sa(self):
        return _arr(self.L.orc_sa(self.h), self.counts().n if self.counts().n else self._n())

    def _n(self):
        raise RuntimeError("run() first")

    def longest(self):
        return _arr(self.L.orc_longest(self.h), self.counts().T)

    def intervals(self, cap=5):
        """dense [T, cap, 2] array of SA intervals, -1 where m > longest."""
        T = self.counts().T
        out = -np.ones((T, cap, 2), dtype=np.int32)
        lg = self.longest()
        up, dn = C.c_int32(), C.c_int32()
        for t in range(T):
            for m in range(1, min(cap, int(lg[t])) + 1):
                self.L.orc_interval(self.h, t, m, C.byref(up), C.byref(dn))
                out[t, m - 1] = (up.value, dn.value)
        return out

    def blocks(self):
        return _arr(self.L.orc_blocks(self.h), self.counts().G, 4)

    def onegap_patterns(self):
        return _arr(self.L.orc_onegap_patterns(self.h), self.counts().D1, 10)

    def onegap_hits(self):
        return _arr(self.L.orc_onegap_hits(self.h), self.counts().hits1, 3)

    def twogap_patterns(self):
        return _arr(self.L.orc_twogap_patterns(self.h), self.counts().D2, 4)

    def twogap_hits(self):
        return _arr(self.L.orc_twogap_hits(self.h), self.counts().hits2, 4)

    def frequent(self):
        return _arr(self.L.orc_frequent(self.h), 100)

    def feature_missing(self):
        return _arr(self.L.orc_feature_missing(self.h), 10000)

    def precomp_index(self):
        return _arr(self.L.orc_precomp_index(self.h), 10000, 2)

    def precomp_list(self):
        return _arr(self.L.orc_precomp_list(self.h), self.counts().precomp_count, 2)

    def query_list(self, which, qi):
        """ids the writer walks for query qi: which = 0 contiguous phrases (first-appearance order), 1 one-gap, 2 two-gap patterns"""
        p = C.POINTER(C.c_int32)()
        n = self.L.orc_query_list(self.h, which, qi, C.byref(p))
        return _arr(p, n)

    def records(self, kind):
        p = C.POINTER(C.c_int32)()
        n = self.L.orc_records(self.h, kind, C.byref(p))
        return _arr(p, n, 7)

    def rules(self, kind):
        p = C.c_void_p()
        n = self.L.orc_rules(self.h, kind, C.byref(p))
        if n == 0:
            return np.zeros(0, dtype=RULE_DTYPE)
        buf = (C.c_char * (n * RULE_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=RULE_DTYPE).copy()


# ---- reference dump parsing (oracle/build_ref_dump.sh) -----------------------------------------
DUMP_DTYPES = {
    "sa": np.dtype("<i4"), "str": np.dtype("<i4"), "qrysbuf": np.dtype("<i4"), "connectoffset": np.dtype("<i4"),
    "result_two": np.dtype([("up", "<i4"), ("down", "<i4"), ("ffh", "<i4"), ("ffhL", "<i4"), ("ffhR", "<i4"), ("longestind", "<i4"), ("longestmatch", "<i4")]),
    "result_connect": np.dtype([("up", "<i4"), ("down", "<i4")]),
    "oneGapSA": np.dtype([("position", "<u4"), ("str_position", "<u4"), ("length", "u1")]),
    "oneGapSearch": np.dtype([("qrystart", "<i4"), ("ls", "u1"), ("le", "u1"), ("gap", "u1"), ("position", "<u4"), ("start", "<i4"), ("end", "<i4")]),
    "onegapPattern": np.dtype([("pattern", "<i4", (5,)), ("number", "u1")]),
    "onegap": np.dtype([("qrystart", "<i4"), ("ls", "u1"), ("le", "u1"), ("gap", "u1")]),
    "twoGapSA": np.dtype([("position", "<u4"), ("str_position", "<u4"), ("length", "u1"), ("length2", "u1")]),
    "twoGapSearch": np.dtype([("blockid", "<u4"), ("gap2", "<u4"), ("le", "u1"), ("position", "<u4"), ("start", "<i4"), ("end", "<i4")]),
    "twogapPattern": np.dtype([("pattern", "<i4", (1,)), ("number", "u1"), ("blockid", "<u4")]),
    "precomp_index": np.dtype([("start", "<u4"), ("end", "<u4")]),
    "precomp_onegap": np.dtype([("start", "<u4"), ("length", "u1")]),
    "featureMissingCount": np.dtype("<i4"), "frequentList": np.dtype("<i4"),
    "out_res": np.dtype([("tar_start", "<i4"), ("blocknumber", "<i4"), ("tar_end", "u1")]),
    "oneGapRule": np.dtype([("id", "<i4"), ("start", "<u4"), ("end", "u1"), ("gap1", "u1"), ("gap1_1", "u1")]),
    "twoGapRule": np.dtype([("id", "<i4"), ("start", "<u4"), ("end", "u1"), ("gap1", "u1"), ("gap1_1", "u1"), ("gap2", "u1"), ("gap2_1", "u1")]),
    "separators": np.dtype("<i4"), "blocks": np.dtype([("start", "<i4"), ("end", "<i4"), ("matchlen", "<i4"), ("string_start", "<i4")]),
    "RLP": np.dtype("<u4"), "L_tar": np.dtype("u1"), "R_tar": np.dtype("u1"), "tgt": np.dtype("<i4"),
}


def load_dump(dump_dir):
    out = {}
    for name, dt in DUMP_DTYPES.items():
        p = os.path.join(dump_dir, name + ".bin")
        if os.path.exists(p):
            out[name] = np.fromfile(p, dtype=dt)
    return out
