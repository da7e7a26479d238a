"""ctypes binding of oracle/libbm2oracle.so (CPU restatement; test infrastructure only)."""
from __future__ import annotations
import ctypes as C, os, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


class BswParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "zdrop", "end_bonus", "vector_quirks")]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "oracle", "libbm2oracle.so")
        src = os.path.join(ROOT, "oracle", "bm2_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"], stdout=subprocess.DEVNULL)
        _LIB = C.CDLL(so)
        _LIB.bm2o_extend_pairs.restype = C.c_int64
        _LIB.bm2o_bsw_extend.restype = C.c_int64
    return _LIB


def bsw_params(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, vector_quirks=1):
    return BswParams(a, b, o_del, e_del, o_ins, e_ins, zdrop, end_bonus, vector_quirks)


def make_pairs(len1, len2, h0, idr, idq):
    from refdump import PAIR_DT
    n = len(len1)
    p = np.zeros(n, PAIR_DT)
    p["len1"] = len1; p["len2"] = len2; p["h0"] = h0; p["idr"] = idr; p["idq"] = idq
    p["id"] = np.arange(n)
    return p


def extend_pairs(pairs, ref, qer, w, params):
    """Runs the oracle in place on a PAIR_DT array; returns banded cell count."""
    ref = np.ascontiguousarray(ref, np.uint8); qer = np.ascontiguousarray(qer, np.uint8)
    assert pairs.flags.c_contiguous
    return lib().bm2o_extend_pairs(pairs.ctypes.data_as(C.c_void_p), ref.ctypes.data_as(C.c_void_p),
                                   qer.ctypes.data_as(C.c_void_p), C.c_int32(len(pairs)), C.c_int32(w), C.byref(params))
