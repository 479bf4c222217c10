"""ctypes view of the C host library (cgx_b200/host/*.c -> lib/libcgx_host.so): the text loaders and the
grammar writer that the `strmatchcuda` binary uses, callable from Python so that tests and bench.py go
through the very same host code path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Result


class Side(C.Structure):
    _fields_ = [("tok", C.POINTER(C.c_int32)), ("n", C.c_int64), ("P", C.POINTER(C.c_uint8)), ("sentenceind", C.POINTER(C.c_int32)),
                ("n_sent", C.c_int32), ("vocab", C.c_void_p), ("last", C.c_int32)]


class Align(C.Structure):
    _fields_ = [("RLP", C.POINTER(C.c_uint32)), ("L_tar", C.POINTER(C.c_uint8)), ("R_tar", C.POINTER(C.c_uint8)), ("wide", C.c_int),
                ("RLP64", C.POINTER(C.c_uint64)), ("L_tar16", C.POINTER(C.c_uint16)), ("R_tar16", C.POINTER(C.c_uint16))]


class Lex(C.Structure):
    _fields_ = [("f", C.POINTER(C.c_int32)), ("e", C.POINTER(C.c_int32)), ("v1", C.POINTER(C.c_float)), ("v2", C.POINTER(C.c_float)),
                ("count", C.c_int64)]


class Queries(C.Structure):
    _fields_ = [("tok", C.POINTER(C.c_int32)), ("off", C.POINTER(C.c_int32)), ("Q", C.c_int32), ("T", C.c_int32), ("max_len", C.c_int32)]


_h = None


def load():
    global _h
    if _h is not None:
        return _h
    _lib.load()   # libcgx_b200.so first (dependency)
    if not os.path.exists(_lib.HOST_LIB_PATH):
        raise RuntimeError(f"{_lib.HOST_LIB_PATH} not found: run __graft_entry__.build()")
    H = C.CDLL(_lib.HOST_LIB_PATH)
    H.cgxh_corpus_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(Side)]
    H.cgxh_side_free.argtypes = [C.POINTER(Side)]
    H.cgxh_vocab_name.argtypes = [C.c_void_p, C.c_int32]
    H.cgxh_vocab_name.restype = C.c_char_p
    H.cgxh_vocab_id.argtypes = [C.c_void_p, C.c_char_p]
    H.cgxh_vocab_id.restype = C.c_int32
    H.cgxh_vocab_size.argtypes = [C.c_void_p]
    H.cgxh_vocab_size.restype = C.c_int32
    H.cgxh_alignment_load.argtypes = [C.c_char_p, C.POINTER(Side), C.POINTER(Side), C.POINTER(Align)]
    H.cgxh_load_files.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(Side), C.POINTER(Side), C.POINTER(Align), C.POINTER(Lex)]
    H.cgxh_lex_load.argtypes = [C.c_char_p, C.POINTER(Side), C.POINTER(Side), C.POINTER(Lex)]
    H.cgxh_queries_load.argtypes = [C.c_char_p, C.POINTER(Side), C.POINTER(Queries)]
    H.cgxh_write_grammars.argtypes = [C.c_char_p, C.POINTER(Result), C.POINTER(C.c_int32), C.c_int32, C.POINTER(Side), C.POINTER(Side), C.c_int]
    _h = H
    return H


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


class HostCorpus:
    """The six strmatchcuda inputs loaded by the C host code (kept alive for the writer)."""

    def __init__(self, src, qry, tgt, align, lex):
        H = load()
        self.H = H
        self.src, self.tgt, self.al, self.lex, self.qry = Side(), Side(), Align(), Lex(), Queries()
        # the loaders the command line runs (run.c): source || target, then alignment || lexical file
        rc = H.cgxh_load_files(src.encode(), tgt.encode(), align.encode(), lex.encode(), C.byref(self.src), C.byref(self.tgt), C.byref(self.al),
                               C.byref(self.lex))
        if rc:
            raise RuntimeError("cannot load the corpus files (%s, %s, %s, %s): code %d" % (src, tgt, align, lex, rc))
        if H.cgxh_queries_load(qry.encode(), C.byref(self.src), C.byref(self.qry)):
            raise RuntimeError("cannot load " + qry)

    def layout(self):
        n, m = int(self.src.n), int(self.tgt.n)
        wide = bool(self.al.wide)
        if wide:      # a sentence of 255 tokens or more: 16-bit alignment fields (cgx_index_build_wide)
            align = dict(RLP=_arr(self.al.RLP64, n, np.uint64), L_tar=_arr(self.al.L_tar16, m, np.uint16), R_tar=_arr(self.al.R_tar16, m, np.uint16))
        else:
            align = dict(RLP=_arr(self.al.RLP, n, np.uint32), L_tar=_arr(self.al.L_tar, m, np.uint8), R_tar=_arr(self.al.R_tar, m, np.uint8))
        return dict(str=_arr(self.src.tok, n + 3, np.int32), n=n, tgt=_arr(self.tgt.tok, m + 3, np.int32), m=m, P=_arr(self.src.P, n, np.uint8),
                    wide=wide, **align,
                    src_sentenceind=_arr(self.src.sentenceind, self.src.n_sent + 1, np.int32),
                    tgt_sentenceind=_arr(self.tgt.sentenceind, self.tgt.n_sent + 1, np.int32),
                    lex_f=_arr(self.lex.f, int(self.lex.count), np.int32), lex_e=_arr(self.lex.e, int(self.lex.count), np.int32),
                    lex_v1=_arr(self.lex.v1, int(self.lex.count), np.float32), lex_v2=_arr(self.lex.v2, int(self.lex.count), np.float32),
                    qry_tok=_arr(self.qry.tok, int(self.qry.T), np.int32), qry_off=_arr(self.qry.off, int(self.qry.Q) + 1, np.int32),
                    src_last=int(self.src.last), tgt_last=int(self.tgt.last))

    def src_name(self, i):
        return (self.H.cgxh_vocab_name(self.src.vocab, int(i)) or b"").decode()

    def tgt_name(self, i):
        return (self.H.cgxh_vocab_name(self.tgt.vocab, int(i)) or b"").decode()

    def write_grammars(self, extractor, outdir, qry_off, qid_base=0, threads=1):
        """print_query_GPU_Gappy for the extractor's last batch."""
        r = Result()
        extractor.L.cgx_result(extractor.h, C.byref(r))
        qo = np.ascontiguousarray(qry_off, dtype=np.int32)
        rc = self.H.cgxh_write_grammars(outdir.encode(), C.byref(r), qo.ctypes.data_as(C.POINTER(C.c_int32)), int(qid_base), C.byref(self.src),
                                        C.byref(self.tgt), int(threads))
        if rc:
            raise RuntimeError("cgxh_write_grammars failed")
