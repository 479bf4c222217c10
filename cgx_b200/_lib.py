"""ctypes binding of libcgx_b200.so (include/cgx_b200.h).  The library is built in-tree by
``cgx_b200/csrc/Makefile`` (``__graft_entry__.build()``); there is no pure-Python / CPU fallback --
importing the extractor without the built CUDA library raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "lib", "libcgx_b200.so")
HOST_LIB_PATH = os.path.join(PKG_DIR, "lib", "libcgx_host.so")


class IndexInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("sa_rounds", C.c_int32), ("sa_key_bits", C.c_int32), ("sa_launches", C.c_int32),
                ("sa_build_ms", C.c_float), ("aux_build_ms", C.c_float), ("index_bytes", C.c_int64)]


class IndexArrays(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("lex_count", C.c_int64), ("max_token", C.c_int32), ("freq_list", C.c_int32 * 100), ("wide", C.c_int32)] + \
               [(k, C.c_void_p) for k in ("str", "sa", "inv1", "inv2", "inv3", "bkt1", "bkt2", "bkt3", "tok_start", "RLP", "L_tar", "R_tar", "tgt", "freq_flag", "gapw",
                                         "lex_key", "lex_v1", "lex_v2")]

    ARRAYS = (("str", 4, "n3"), ("sa", 4, "n"), ("inv1", 4, "n"), ("inv2", 4, "n"), ("inv3", 4, "n"), ("bkt1", 4, "n"), ("bkt2", 4, "n"), ("bkt3", 4, "n"), ("tok_start", 4, "nt"), ("RLP", 4, "n"),
              ("L_tar", 1, "m"), ("R_tar", 1, "m"), ("tgt", 4, "m3"), ("freq_flag", 1, "nt"), ("gapw", 4, "n"), ("lex_key", 8, "lex1"), ("lex_v1", 4, "lex1"),
              ("lex_v2", 4, "lex1"))

    def nbytes(self, name):
        for k, size, dim in self.ARRAYS:
            if k == name:
                cnt = {"n": self.n, "n3": self.n + 3, "m": self.m, "m3": self.m + 3, "nt": self.max_token + 2, "lex1": self.lex_count + 1}[dim]
                if self.wide and name in ("RLP", "L_tar", "R_tar"):      # 16-bit alignment fields: twice the bytes
                    size *= 2
                return int(cnt) * size
        raise KeyError(name)


class BatchInfo(C.Structure):
    _fields_ = [("Q", C.c_int32), ("T", C.c_int32), ("G", C.c_int32), ("enu1", C.c_int32), ("D1", C.c_int32), ("hits1", C.c_int64),
                ("enu2", C.c_int32), ("D2", C.c_int32), ("hits2", C.c_int64), ("samples", C.c_int64), ("n_ab", C.c_int64), ("n_1gap", C.c_int64),
                ("n_2gap", C.c_int64), ("rules", C.c_int32 * 3), ("launches", C.c_int32), ("ms_total", C.c_float), ("ms_lookup", C.c_float),
                ("ms_enum", C.c_float), ("ms_join", C.c_float), ("ms_extract", C.c_float), ("ms_aggregate", C.c_float)]

    def as_dict(self):
        d = {}
        for k, _ in self._fields_:
            v = getattr(self, k)
            d[k] = list(v) if hasattr(v, "__len__") else v
        return d


RULE_DTYPE = np.dtype([("id", "<i4"), ("tgt_start", "<i4"), ("end", "u1"), ("gap1", "u1"), ("gap1_1", "u1"), ("gap2", "u1"), ("gap2_1", "u1"),
                       ("pad", "u1"), ("f", "<u2"), ("fs", "<u2"), ("pc", "<u2"), ("mlfe", "<f4"), ("mlef", "<f4")])
assert RULE_DTYPE.itemsize == 28
# wire format (cgx_rule_t, 16 bytes); BatchResult decodes it into RULE_DTYPE with id / f / fs from first and idinfo
RULE_WIRE_DTYPE = np.dtype([("tgt_start", "<i4"), ("span", "<u4"), ("mlfe", "<f4"), ("mlef", "<f4")])
assert RULE_WIRE_DTYPE.itemsize == 16


class Result(C.Structure):
    _fields_ = [("Q", C.c_int32), ("T", C.c_int32), ("G", C.c_int32), ("D1", C.c_int32), ("D2", C.c_int32),
                ("phrase_id", C.POINTER(C.c_int32)), ("phrases", C.POINTER(C.c_int32)), ("pat1", C.POINTER(C.c_int32)), ("pat2", C.POINTER(C.c_int32)),
                ("q1_off", C.POINTER(C.c_int32)), ("q1_ids", C.POINTER(C.c_int32)), ("q2_off", C.POINTER(C.c_int32)), ("q2_ids", C.POINTER(C.c_int32)),
                ("rules", C.c_void_p * 3), ("n_rules", C.c_int32 * 3), ("first", C.POINTER(C.c_int32) * 3), ("n_ids", C.c_int32 * 3),
                ("idinfo", C.POINTER(C.c_uint32) * 3)]


_lib = None


def load():
    """Load libcgx_b200.so; fail loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(cgx_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32p, u32p, u8p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    L.cgx_version.restype = C.c_int
    L.cgx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.cgx_destroy.argtypes = [vp]
    L.cgx_destroy.restype = None
    L.cgx_last_error.argtypes = [vp]
    L.cgx_last_error.restype = C.c_char_p
    L.cgx_index_build.argtypes = [vp, i32p, C.c_int64, i32p, C.c_int64, u32p, u8p, u8p]
    L.cgx_index_build_wide.argtypes = [vp, i32p, C.c_int64, i32p, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
    L.cgx_lex_load.argtypes = [vp, i32p, i32p, f32p, f32p, C.c_int64]
    L.cgx_index_info.argtypes = [vp, C.POINTER(IndexInfo)]
    L.cgx_sa_build_dev.argtypes = [vp, vp, C.c_int64, C.c_int32, vp, i32p, f32p]
    L.cgx_index_export.argtypes = [vp, C.POINTER(IndexArrays)]
    L.cgx_index_alloc.argtypes = [vp, C.POINTER(IndexArrays), C.POINTER(IndexArrays)]
    L.cgx_index_commit.argtypes = [vp]
    L.cgx_index_save.argtypes = [vp, C.c_char_p]
    L.cgx_index_load.argtypes = [vp, C.c_char_p]
    L.cgx_index_copy_sa.argtypes = [vp, i32p]
    L.cgx_index_copy_inv.argtypes = [vp, C.c_int, i32p]
    L.cgx_index_copy_frequent.argtypes = [vp, i32p]
    L.cgx_extract.argtypes = [vp, i32p, i32p, C.c_int32]
    L.cgx_extract_begin.argtypes = [vp, i32p, i32p, C.c_int32]
    L.cgx_result_at.argtypes = [vp, C.c_int, C.POINTER(Result)]
    L.cgx_extract_dev.argtypes = [vp, vp, vp, vp, C.c_int32, C.c_int32]
    L.cgx_profile_enable.argtypes = [vp, C.c_int]
    L.cgx_profile_report.argtypes = [vp]
    L.cgx_profile_report.restype = C.c_char_p
    L.cgx_index_broadcast.argtypes = [C.POINTER(vp), C.c_int]
    L.cgx_batch_info.argtypes = [vp, C.POINTER(BatchInfo)]
    L.cgx_batch_advice.argtypes = [vp, C.c_int32]
    L.cgx_batch_advice.restype = C.c_int32
    L.cgx_result.argtypes = [vp, C.POINTER(Result)]
    L.cgx_debug_fetch.argtypes = [vp, C.c_char_p, i32p, C.c_int64]
    L.cgx_debug_fetch.restype = C.c_int64
    L.cgx_debug_sort_u64.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    _lib = L
    return L


EXPORTED_SYMBOLS = ("cgx_version", "cgx_create", "cgx_destroy", "cgx_last_error", "cgx_index_build", "cgx_index_build_wide", "cgx_index_matches", "cgx_lex_load", "cgx_index_info",
                    "cgx_sa_build_dev", "cgx_index_export", "cgx_index_alloc", "cgx_index_commit", "cgx_index_save", "cgx_index_load", "cgx_index_copy_sa", "cgx_index_copy_inv",
                    "cgx_index_copy_frequent", "cgx_extract", "cgx_extract_begin", "cgx_result_at", "cgx_extract_dev", "cgx_profile_enable", "cgx_profile_report",
                    "cgx_index_broadcast", "cgx_batch_info", "cgx_batch_advice", "cgx_result", "cgx_debug_fetch", "cgx_debug_sort_u64")
