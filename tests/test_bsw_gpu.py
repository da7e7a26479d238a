"""GPU parity of seam 1 (bm2_extend_pairs) through the C ABI: bit-exact against the oracle on
seeded inputs, against the reference's golden vectors, and on edge cases."""
import numpy as np
import pytest
import oracle_lib as ol
from bsw_util import random_jobs as _random_jobs

pytestmark = pytest.mark.gpu
OUT = ("score", "tle", "gtle", "qle", "gscore", "max_off")


@pytest.fixture(autouse=True, params=["col2", "cell"])
def bsw_kernel(request, monkeypatch):
    """Every test runs twice: 8-bit-score jobs on the two-columns-per-instruction kernel (bsw_col2_kernel, the default)
    and on the one-cell-per-instruction kernel (BM2_BSW_COL2=0); the library reads the variable at every launch."""
    monkeypatch.setenv("BM2_BSW_COL2", "1" if request.param == "col2" else "0")
    return request.param


def _assert_same(got, want):
    for f in OUT:
        bad = np.nonzero(got[f] != want[f])[0]
        assert len(bad) == 0, (f, bad[:5], got[f][bad[:5]], want[f][bad[:5]], got["len1"][bad[:5]], got["len2"][bad[:5]], got["h0"][bad[:5]])


def test_golden_reference_vectors(pkg, gpu_ctx, golden_dir):
    g = np.load(golden_dir + "/bsw_c0.npz")
    p = np.zeros(len(g["h0"]), pkg.capi.PAIR_DT)
    for f in ("len1", "len2", "h0", "idr", "idq"):
        p[f] = g[f]
    gpu_ctx.extend_pairs(p, g["ref"], g["qer"], int(g["w"]), int(g["p_end_bonus"]))
    for f in OUT:
        assert np.array_equal(p[f], g["out_" + f]), f


@pytest.mark.parametrize("seed,qmax,tmax,w", [(1, 151, 400, 100), (2, 151, 400, 200), (3, 40, 90, 100), (4, 700, 900, 100),
                                               (5, 1500, 1700, 100), (6, 151, 400, 10),
                                               # long queries: the warp-per-job kernel (row-parallel scan), three band widths
                                               (7, 4000, 4300, 100), (8, 3000, 3300, 200), (9, 2500, 2800, 300), (10, 2000, 2200, 600)])
def test_random_jobs_match_oracle(pkg, gpu_ctx, seed, qmax, tmax, w):
    rng = np.random.default_rng(seed)
    n = 3000 if qmax <= 151 else (400 if qmax <= 1500 else 120)
    len1, len2, h0, idr, idq, ref, qer = _random_jobs(rng, n, qmax, tmax)
    p = np.zeros(n, pkg.capi.PAIR_DT)
    p["len1"] = len1; p["len2"] = len2; p["h0"] = h0; p["idr"] = idr; p["idq"] = idq
    want = p.copy()
    gpu_ctx.extend_pairs(p, ref, qer, w, 5)
    ol.extend_pairs(want, ref, qer, w, ol.bsw_params(end_bonus=5))
    _assert_same(p, want)


@pytest.mark.parametrize("seed,qmax,tmax,w", [(21, 151, 400, 100), (22, 60, 200, 100), (23, 151, 300, 7)])
def test_random_low_score_jobs_match_oracle(pkg, gpu_ctx, seed, qmax, tmax, w, monkeypatch):
    # small h0, almost no N: with BM2_BSW_PAIR=1 most jobs take the experimental two-jobs-per-thread kernel
    # (8-bit scores, bsw_pair_kernel; off by default because it measured slower than the thread-per-job kernel)
    monkeypatch.setenv("BM2_BSW_PAIR", "1")
    rng = np.random.default_rng(seed)
    n = 6001
    len1, len2, h0, idr, idq, ref, qer = _random_jobs(rng, n, qmax, tmax, nrate=0.001, h0max=60)
    p = np.zeros(n, pkg.capi.PAIR_DT)
    p["len1"] = len1; p["len2"] = len2; p["h0"] = h0; p["idr"] = idr; p["idq"] = idq
    want = p.copy()
    gpu_ctx.extend_pairs(p, ref, qer, w, 5)
    ol.extend_pairs(want, ref, qer, w, ol.bsw_params(end_bonus=5))
    _assert_same(p, want)


@pytest.mark.parametrize("seed,qmax,tmax,w,nrate,h0max", [(31, 151, 400, 100, 0.05, 100), (32, 256, 500, 100, 0.01, 40), (33, 9, 40, 100, 0.1, 30),
                                                           (34, 151, 300, 1, 0.01, 60), (35, 130, 400, 40, 0.0, 120)])
def test_random_8bit_jobs_match_oracle(pkg, gpu_ctx, seed, qmax, tmax, w, nrate, h0max):
    # the 8-bit-score classes up to 256 columns (bsw_col2_kernel's domain): N-rich queries, tiny jobs, narrow bands
    rng = np.random.default_rng(seed)
    n = 5003
    len1, len2, h0, idr, idq, ref, qer = _random_jobs(rng, n, qmax, tmax, nrate=nrate, h0max=h0max)
    p = np.zeros(n, pkg.capi.PAIR_DT)
    p["len1"] = len1; p["len2"] = len2; p["h0"] = h0; p["idr"] = idr; p["idq"] = idq
    want = p.copy()
    gpu_ctx.extend_pairs(p, ref, qer, w, 5)
    ol.extend_pairs(want, ref, qer, w, ol.bsw_params(end_bonus=5))
    _assert_same(p, want)


def test_edge_cases(pkg, gpu_ctx):
    # empty target, single-base query/target, all-N query, h0 near the int16 class limit, >int16 scores (wide path)
    seq = np.array([0, 1, 2, 3] * 64, np.uint8)
    nq = np.full(64, 4, np.uint8)
    qer = np.concatenate([seq[:4], seq[:1], nq, seq[:200], seq[:200]])
    ref = np.concatenate([seq[:1], seq[:64], seq[:200], seq[:200]])
    p = np.zeros(5, pkg.capi.PAIR_DT)
    p["len1"] = [0, 1, 64, 200, 200]; p["len2"] = [4, 1, 64, 200, 200]
    p["idr"] = [0, 0, 1, 65, 265]; p["idq"] = [0, 4, 5, 69, 269]
    p["h0"] = [17, 3, 9, 32000, 40000]
    want = p.copy()
    gpu_ctx.extend_pairs(p, ref, qer, 100, 5)
    ol.extend_pairs(want, ref, qer, 100, ol.bsw_params(end_bonus=5))
    _assert_same(p, want)
    assert p["score"][0] == 17 and p["gscore"][0] == -1
    assert p["score"][4] == 40200


@pytest.mark.parametrize("sc", [dict(a=1, b=1, o_del=1, e_del=1, o_ins=1, e_ins=1, zdrop=100),     # -x ont2d scoring
                                dict(a=2, b=3, o_del=4, e_del=2, o_ins=5, e_ins=1, zdrop=30)])     # o_del+e_del != o_ins+e_ins
def test_non_default_scoring(pkg, golden_dir, sc):
    o = pkg.capi.default_opt()
    o.a, o.b, o.o_del, o.e_del, o.o_ins, o.e_ins, o.zdrop = sc["a"], sc["b"], sc["o_del"], sc["e_del"], sc["o_ins"], sc["e_ins"], sc["zdrop"]
    ctx = pkg.capi.Context(0, opt=o)
    rng = np.random.default_rng(11)
    len1, len2, h0, idr, idq, ref, qer = _random_jobs(rng, 1500, 300, 500, sim=0.85, h0max=60)
    p = np.zeros(1500, pkg.capi.PAIR_DT)
    p["len1"] = len1; p["len2"] = len2; p["h0"] = h0; p["idr"] = idr; p["idq"] = idq
    want = p.copy()
    ctx.extend_pairs(p, ref, qer, 100, 0)
    ol.extend_pairs(want, ref, qer, 100, ol.bsw_params(end_bonus=0, **sc))
    _assert_same(p, want)
    ctx.close()


@pytest.mark.parametrize("sc", [dict(zdrop=0), dict(zdrop=128), dict(zdrop=200), dict(a=2, b=8, o_del=12, e_del=2, o_ins=12, e_ins=2, zdrop=200),
                                dict(a=3, b=12, o_del=18, e_del=3, o_ins=18, e_ins=3, zdrop=90)])
def test_zdrop_and_band_quirks_of_the_simd_classes(pkg, sc):
    # the reference's SIMD kernels test z-drop on every row without a `zdrop > 0` guard, and its 8-bit class (len < 128,
    # h0 + min(len) * a < 128) keeps the threshold and the band operands in 8 bits: -d 0, -d 128.., -A 2 (which scales -d to 200)
    o = pkg.capi.default_opt()
    for k, v in sc.items():
        setattr(o, k, v)
    ctx = pkg.capi.Context(0, opt=o)
    rng = np.random.default_rng(77)
    n = 4000
    len1, len2, h0, idr, idq, ref, qer = _random_jobs(rng, n, 151, 300, sim=0.9, nrate=0.005, h0max=70)
    p = np.zeros(n, pkg.capi.PAIR_DT)
    p["len1"] = len1; p["len2"] = len2; p["h0"] = h0; p["idr"] = idr; p["idq"] = idq
    want = p.copy()
    ctx.extend_pairs(p, ref, qer, 100, 5 * o.a)
    prm = ol.bsw_params(a=o.a, b=o.b, o_del=o.o_del, e_del=o.e_del, o_ins=o.o_ins, e_ins=o.e_ins, zdrop=o.zdrop, end_bonus=5 * o.a)
    ol.extend_pairs(want, ref, qer, 100, prm)
    _assert_same(p, want)
    assert ((len1 < 128) & (len2 < 128) & (h0 + np.minimum(len1, len2) * o.a < 128)).sum() > 200      # the 8-bit class is covered
    ctx.close()
