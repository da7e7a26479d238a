"""The two-columns-per-instruction BSW DP (bwa-mem2_b200/csrc/bsw_col2.cuh: columns 2p / 2p+1 of one job in the packed
16-bit halves, what bsw_col2_kernel runs per thread) compiled for the host with portable stand-ins for the packed
instructions (tests/host_emul/bsw_col2_emul.cpp) and checked against the oracle / the reference's golden vectors.
CPU-only; the kernel itself is covered by the `-m gpu` BSW and pipeline tests."""
import ctypes as C, os, subprocess
import numpy as np
import pytest
import oracle_lib as ol
from bsw_util import random_jobs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = ("score", "tle", "gtle", "qle", "gscore", "max_off")
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        d = os.path.join(ROOT, "tests", "host_emul")
        so = os.path.join(d, "libbswcol2.so")
        srcs = [os.path.join(d, "bsw_col2_emul.cpp")] + [os.path.join(ROOT, "bwa-mem2_b200", "csrc", f)
                                                          for f in ("bsw_col2.cuh", "bsw_pair.cuh", "bsw_types.h", "hd.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "bwa-mem2_b200", "csrc"),
                                   "-I" + os.path.join(ROOT, "include"), srcs[0], "-o", so])
        _LIB = C.CDLL(so)
        _LIB.col2_extend_all.restype = C.c_longlong
    return _LIB


def _eligible(len1, len2, h0, a):
    """The routing rule of the column-pair kernel (bsw.cu): 8-bit scores, <= 256 columns (N bases are fine)."""
    return (h0 + np.minimum(len1, len2) * a <= 255) & (len2 <= 256)


def _run(idx, len1, len2, h0, idr, idq, ref, qer, prm, w, qstride=None, tstride=None):
    n = len(idx)
    i64 = lambda x: np.ascontiguousarray(x, np.int64); i32 = lambda x: np.ascontiguousarray(x, np.int32)
    qoff = i64(idq[idx]); toff = i64(idr[idx]); ql = i32(len2[idx]); tl = i32(len1[idx]); hh = i32(h0[idx])
    qs = i32(np.ones(n) if qstride is None else qstride); ts = i32(np.ones(n) if tstride is None else tstride)
    p = i32([prm.a, prm.b, prm.o_del, prm.e_del, prm.o_ins, prm.e_ins, prm.zdrop, prm.end_bonus, w])
    out = np.zeros((n, 6), np.int32)
    qer = np.ascontiguousarray(qer, np.uint8); ref = np.ascontiguousarray(ref, np.uint8)
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    cells = _lib().col2_extend_all(C.c_int(n), P(qoff), P(toff), P(ql), P(tl), P(hh), P(qs), P(ts), P(qer), P(ref), P(p), P(out))
    assert cells >= 0
    return out, cells


def _check(len1, len2, h0, idr, idq, ref, qer, prm, w, min_frac):
    idx = np.nonzero(_eligible(len1, len2, h0, prm.a))[0]
    assert len(idx) >= min_frac * len(len1), (len(idx), len(len1))
    want = ol.make_pairs(len1, len2, h0, idr, idq)
    ol.extend_pairs(want, ref, qer, w, prm)
    got, cells = _run(idx, len1, len2, h0, idr, idq, ref, qer, prm, w)
    for k, f in enumerate(OUT):
        bad = np.nonzero(got[:, k] != want[f][idx])[0]
        assert len(bad) == 0, (f, idx[bad[:5]], got[bad[:5], k], want[f][idx[bad[:5]]], len1[idx[bad[:5]]], len2[idx[bad[:5]]], h0[idx[bad[:5]]])
    return len(idx), cells


def test_col2_dp_matches_reference_golden(golden_dir):
    g = np.load(golden_dir + "/bsw_c0.npz")
    prm = ol.bsw_params(a=int(g["p_a"]), b=int(g["p_b"]), o_del=int(g["p_o_del"]), e_del=int(g["p_e_del"]),
                        o_ins=int(g["p_o_ins"]), e_ins=int(g["p_e_ins"]), zdrop=int(g["p_zdrop"]), end_bonus=int(g["p_end_bonus"]))
    idx = np.nonzero(_eligible(g["len1"], g["len2"], g["h0"], prm.a))[0]
    assert len(idx) > 0.9 * len(g["len1"])
    got, _ = _run(idx, g["len1"], g["len2"], g["h0"], g["idr"], g["idq"], g["ref"], g["qer"], prm, int(g["w"]))
    for k, f in enumerate(OUT):
        assert np.array_equal(got[:, k], g["out_" + f][idx]), f


@pytest.mark.parametrize("seed,qmax,tmax,w,nrate", [(1, 151, 400, 100, 0.002), (2, 151, 400, 100, 0.05), (3, 40, 90, 100, 0.01),
                                                     (4, 151, 300, 10, 0.01), (5, 256, 500, 200, 0.01), (6, 151, 400, 3, 0.0),
                                                     (7, 7, 30, 100, 0.1), (8, 200, 400, 1, 0.01)])
def test_col2_dp_matches_oracle_on_random_jobs(seed, qmax, tmax, w, nrate):
    rng = np.random.default_rng(seed)
    len1, len2, h0, idr, idq, ref, qer = random_jobs(rng, 3000, qmax, tmax, nrate=nrate, h0max=60)
    n, cells = _check(len1, len2, h0, idr, idq, ref, qer, ol.bsw_params(end_bonus=5), w, 0.3)
    assert cells > 0


@pytest.mark.parametrize("scoring", [dict(a=1, b=1, o_del=1, e_del=1, o_ins=1, e_ins=1, zdrop=100, end_bonus=0),
                                     dict(a=2, b=3, o_del=4, e_del=2, o_ins=5, e_ins=1, zdrop=30, end_bonus=7),
                                     dict(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=0, end_bonus=5),
                                     dict(a=3, b=9, o_del=2, e_del=3, o_ins=1, e_ins=2, zdrop=50, end_bonus=10)])
def test_col2_dp_non_default_scoring(scoring):
    rng = np.random.default_rng(17)
    len1, len2, h0, idr, idq, ref, qer = random_jobs(rng, 2000, 80, 300, sim=0.85, nrate=0.002, h0max=40)
    _check(len1, len2, h0, idr, idq, ref, qer, ol.bsw_params(**scoring), 100, 0.2)


def test_col2_dp_reversed_strides():
    """Left extensions read the query and the target backwards in place (stride -1)."""
    rng = np.random.default_rng(23)
    len1, len2, h0, idr, idq, ref, qer = random_jobs(rng, 1500, 120, 300, nrate=0.01, h0max=60)
    prm = ol.bsw_params(end_bonus=5)
    idx = np.nonzero(_eligible(len1, len2, h0, prm.a))[0]
    want = ol.make_pairs(len1, len2, h0, idr, idq)
    ol.extend_pairs(want, ref, qer, 100, prm)
    # the same jobs on buffers stored back to front, addressed with stride -1 from the last base
    qrev = qer[::-1].copy(); rrev = ref[::-1].copy()
    idq_r = len(qer) - 1 - idq; idr_r = len(ref) - 1 - idr
    got, _ = _run(idx, len1, len2, h0, idr_r, idq_r, rrev, qrev, prm, 100, qstride=-np.ones(len(idx)), tstride=-np.ones(len(idx)))
    for k, f in enumerate(OUT):
        assert np.array_equal(got[:, k], want[f][idx]), f
