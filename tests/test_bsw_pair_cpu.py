"""The two-jobs-per-thread BSW DP (bwa-mem2_b200/csrc/bsw_pair.cuh: packed 16-bit halves, what bsw_pair_kernel runs
per thread) compiled for the host with portable stand-ins for the packed instructions (tests/host_emul/bsw_pair_emul.cpp)
and checked against the oracle / the reference's golden vectors.  CPU-only; the kernel itself is covered by the
`-m gpu` BSW and pipeline tests."""
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
        so = os.path.join(d, "libbswpair.so")
        srcs = [os.path.join(d, "bsw_pair_emul.cpp")] + [os.path.join(ROOT, "bwa-mem2_b200", "csrc", f) for f in ("bsw_pair.cuh", "bsw_types.h", "hd.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "bwa-mem2_b200", "csrc"),
                                   "-I" + os.path.join(ROOT, "include"), srcs[0], "-o", so])
        _LIB = C.CDLL(so)
        _LIB.pair_extend_all.restype = C.c_longlong
    return _LIB


def _eligible(len1, len2, h0, idq, qer, a):
    """The routing rule of the pair classes (bsw.cu bsw_class_of): 8-bit scores, <= 255 columns, no N in the query."""
    ok = (h0 + np.minimum(len1, len2) * a <= 255) & (len2 <= 255)
    isn = np.concatenate([[0], np.cumsum(qer > 3)])
    ok &= (isn[idq + len2] - isn[idq]) == 0
    return ok


def _run_pairs(order, len1, len2, h0, idr, idq, ref, qer, prm, w):
    n = len(order)
    i64 = lambda x: np.ascontiguousarray(x, np.int64); i32 = lambda x: np.ascontiguousarray(x, np.int32)
    qoff = i64(idq[order]); toff = i64(idr[order]); ql = i32(len2[order]); tl = i32(len1[order]); hh = i32(h0[order])
    p = i32([prm.a, prm.b, prm.o_del, prm.e_del, prm.o_ins, prm.e_ins, prm.zdrop, prm.end_bonus, w])
    out = np.zeros((n, 6), np.int32)
    qer = np.ascontiguousarray(qer, np.uint8); ref = np.ascontiguousarray(ref, np.uint8)
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    cells = _lib().pair_extend_all(C.c_int(n), P(qoff), P(toff), P(ql), P(tl), P(hh), P(qer), P(ref), P(p), P(out))
    assert cells >= 0
    return out, cells


def _check(len1, len2, h0, idr, idq, ref, qer, prm, w, order_kind, rng, min_frac):
    ok = _eligible(len1, len2, h0, idq, qer, prm.a)
    idx = np.nonzero(ok)[0]
    assert len(idx) >= min_frac * len(len1), (len(idx), len(len1))
    if order_kind == "sorted":       # as the kernel pairs them: target length, then query length, descending
        idx = idx[np.lexsort((-len2[idx], -len1[idx]))]
    else:                            # arbitrary partners: very different bands / exit rows inside a pair
        idx = rng.permutation(idx)
    want = ol.make_pairs(len1, len2, h0, idr, idq)
    wcells = ol.extend_pairs(want, ref, qer, w, prm)
    got, cells = _run_pairs(idx, len1, len2, h0, idr, idq, ref, qer, prm, w)
    for k, f in enumerate(OUT):
        bad = np.nonzero(got[:, k] != want[f][idx])[0]
        assert len(bad) == 0, (f, idx[bad[:5]], got[bad[:5], k], want[f][idx[bad[:5]]], len1[idx[bad[:5]]], len2[idx[bad[:5]]], h0[idx[bad[:5]]])
    return len(idx)


def test_pair_dp_matches_reference_golden(golden_dir):
    g = np.load(golden_dir + "/bsw_c0.npz")
    prm = ol.bsw_params(a=int(g["p_a"]), b=int(g["p_b"]), o_del=int(g["p_o_del"]), e_del=int(g["p_e_del"]),
                        o_ins=int(g["p_o_ins"]), e_ins=int(g["p_e_ins"]), zdrop=int(g["p_zdrop"]), end_bonus=int(g["p_end_bonus"]))
    ok = _eligible(g["len1"], g["len2"], g["h0"], g["idq"], g["qer"], prm.a)
    idx = np.nonzero(ok)[0]
    assert len(idx) > 0.8 * len(ok)
    idx = idx[np.lexsort((-g["len2"][idx], -g["len1"][idx]))]
    got, _ = _run_pairs(idx, g["len1"], g["len2"], g["h0"], g["idr"], g["idq"], g["ref"], g["qer"], prm, int(g["w"]))
    for k, f in enumerate(OUT):
        assert np.array_equal(got[:, k], g["out_" + f][idx]), f


@pytest.mark.parametrize("seed,qmax,tmax,w,order", [(1, 151, 400, 100, "sorted"), (2, 151, 400, 100, "random"), (3, 40, 90, 100, "random"),
                                                     (4, 151, 300, 10, "random"), (5, 250, 500, 200, "sorted"), (6, 151, 400, 3, "random")])
def test_pair_dp_matches_oracle_on_random_jobs(seed, qmax, tmax, w, order):
    rng = np.random.default_rng(seed)
    len1, len2, h0, idr, idq, ref, qer = random_jobs(rng, 3000, qmax, tmax, nrate=0.002, h0max=60)
    _check(len1, len2, h0, idr, idq, ref, qer, ol.bsw_params(end_bonus=5), w, order, rng, 0.3)


@pytest.mark.parametrize("scoring", [dict(a=1, b=1, o_del=1, e_del=1, o_ins=1, e_ins=1, zdrop=100, end_bonus=0),
                                     dict(a=2, b=3, o_del=4, e_del=2, o_ins=5, e_ins=1, zdrop=30, end_bonus=7),
                                     dict(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=0, end_bonus=5)])
def test_pair_dp_non_default_scoring(scoring):
    rng = np.random.default_rng(17)
    len1, len2, h0, idr, idq, ref, qer = random_jobs(rng, 2000, 100, 300, sim=0.85, nrate=0.002, h0max=40)
    _check(len1, len2, h0, idr, idq, ref, qer, ol.bsw_params(**scoring), 100, "random", rng, 0.3)
