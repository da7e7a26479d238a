"""Pins the oracle's extension DP to the reference: golden vectors were produced by the unmodified
reference's AVX-512 getScores8/getScores16 (tests/golden/make_golden.py)."""
import numpy as np
import oracle_lib as ol

OUT = ("score", "tle", "gtle", "qle", "gscore", "max_off")


def _load(golden_dir):
    g = np.load(golden_dir + "/bsw_c0.npz")
    p = ol.make_pairs(g["len1"], g["len2"], g["h0"], g["idr"], g["idq"])
    return g, p


def test_oracle_matches_reference_golden(golden_dir):
    g, p = _load(golden_dir)
    prm = ol.bsw_params(a=int(g["p_a"]), b=int(g["p_b"]), o_del=int(g["p_o_del"]), e_del=int(g["p_e_del"]),
                        o_ins=int(g["p_o_ins"]), e_ins=int(g["p_e_ins"]), zdrop=int(g["p_zdrop"]), end_bonus=int(g["p_end_bonus"]))
    cells = ol.extend_pairs(p, g["ref"], g["qer"], int(g["w"]), prm)
    assert cells > 0
    for f in OUT:
        assert np.array_equal(p[f], g["out_" + f]), f
    assert set(np.unique(g["kind"])) >= {8, 16}     # both SIMD classes of the reference are covered


def test_oracle_scalar_and_vector_band_agree_for_default_scores(golden_dir):
    g, p = _load(golden_dir)
    q = p.copy()
    ol.extend_pairs(p, g["ref"], g["qer"], 100, ol.bsw_params(vector_quirks=1))
    ol.extend_pairs(q, g["ref"], g["qer"], 100, ol.bsw_params(vector_quirks=0))
    for f in OUT:
        assert np.array_equal(p[f], q[f]), f


def test_oracle_edge_cases():
    prm = ol.bsw_params()
    # empty target: score stays h0, nothing aligned
    ref = np.zeros(4, np.uint8); qer = np.array([0, 1, 2, 3], np.uint8)
    p = ol.make_pairs([0], [4], [17], [0], [0])
    ol.extend_pairs(p, ref, qer, 100, prm)
    assert (p["score"][0], p["qle"][0], p["tle"][0], p["gtle"][0], p["gscore"][0], p["max_off"][0]) == (17, 0, 0, 0, -1, 0)
    # perfect match to the end: global score == local score
    seq = np.array([0, 1, 2, 3] * 10, np.uint8)
    p = ol.make_pairs([40], [40], [20], [0], [0])
    ol.extend_pairs(p, seq, seq, 100, prm)
    assert p["score"][0] == 60 and p["qle"][0] == 40 and p["tle"][0] == 40 and p["gscore"][0] == 60 and p["gtle"][0] == 40
    # all-N query: every cell scores -1, extension dies, score == h0
    p = ol.make_pairs([40], [40], [5], [0], [0])
    ol.extend_pairs(p, seq, np.full(40, 4, np.uint8), 100, prm)
    assert p["score"][0] == 5 and p["qle"][0] == 0
