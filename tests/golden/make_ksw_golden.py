"""Generates tests/golden/ksw_c0.npz from the UNMODIFIED reference: mate-rescue shaped local-alignment requests (query, reference
window, xtra of mem_matesw) and the kswr_t results of the reference's own ksw_align2 (src/ksw.cpp:324) through
`oracle/_ref/<isa>/ref_driver ksw`.  Run here (container with /root/reference and oracle/_ref built):
    python tests/golden/make_ksw_golden.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import ksw_util as ku  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    reqs = ku.make_requests(rng, 1500, qlens=(151, 151, 151, 100, 76, 36, 250, 300))
    out = ku.reference_ksw(reqs)
    qoff = np.concatenate([[0], np.cumsum([len(r[0]) for r in reqs])]).astype(np.int64)
    toff = np.concatenate([[0], np.cumsum([len(r[1]) for r in reqs])]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "ksw_c0.npz"), query=np.concatenate([r[0] for r in reqs]), target=np.concatenate([r[1] for r in reqs]),
                        qoff=qoff, toff=toff, xtra=np.array([r[2] for r in reqs], np.int32), out=out)
    print(len(reqs), "requests;", int((out[:, 5] >= 0).sum()), "with start positions;", int((out[:, 3] > 0).sum()), "with a second-best score")


if __name__ == "__main__":
    main()
