"""Mate-rescue local alignment (groundwork for the next widening step): the oracle's scalar restatement of the
reference's striped SSE2 kernels (ksw_align2 = ksw_u8 / ksw_i16 forward + reversed-prefix pass, src/ksw.cpp:111-381) against
golden vectors made by the UNMODIFIED reference (tests/golden/make_ksw_golden.py) and, when oracle/_ref is built, against the
reference itself on fresh request sets."""
import numpy as np
import pytest
import ksw_util as ku


def _golden(golden_dir):
    g = np.load(golden_dir + "/ksw_c0.npz")
    reqs = [(g["query"][g["qoff"][i]:g["qoff"][i + 1]], g["target"][g["toff"][i]:g["toff"][i + 1]], int(g["xtra"][i])) for i in range(len(g["xtra"]))]
    return reqs, g["out"]


def test_oracle_matches_reference_golden(pkg, golden_dir):
    reqs, want = _golden(golden_dir)
    got = ku.oracle_ksw(reqs, pkg.capi.default_opt())
    assert np.array_equal(got, want)
    assert (want[:, 5] >= 0).sum() > 800 and (want[:, 3] > 0).sum() > 200      # start positions and second-best scores are covered
    assert sum(1 for r in reqs if r[2] & ku.KSW_XBYTE) > 500 and sum(1 for r in reqs if not r[2] & ku.KSW_XBYTE) > 200      # both kernels


@pytest.mark.parametrize("seed,qlens", [(11, (151,)), (12, (36, 50, 76, 100)), (13, (249, 250, 251, 400)), (14, (15, 16, 17, 8, 9))])
def test_oracle_matches_the_live_reference(pkg, seed, qlens):
    if ku.refbin() is None:
        pytest.skip("oracle/_ref not built")
    reqs = ku.make_requests(np.random.default_rng(seed), 1200, qlens=qlens)
    want = ku.reference_ksw(reqs)
    got = ku.oracle_ksw(reqs, pkg.capi.default_opt())
    bad = np.nonzero((got != want).any(1))[0]
    assert len(bad) == 0, (bad[:5], got[bad[:5]], want[bad[:5]])
