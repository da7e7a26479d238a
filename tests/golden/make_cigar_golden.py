"""Generates tests/golden/cigar_c0.npz from the UNMODIFIED reference: requests of bwa_gen_cigar2 on the C0 data set (final
alignment regions of the oracle == reference, several band limits, perturbed end points, rejected cases) and the outputs of
the reference's own bwa_gen_cigar2 (src/bwa.cpp:260) through `oracle/_ref/<isa>/ref_driver cigar`.

Run here (container with /root/reference and oracle/_ref built):  python tests/golden/make_cigar_golden.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
import oracle_lib as ol, cigar_util as cu  # noqa: E402


def main():
    capi = load_package().capi
    prefix = os.path.join(HERE, "c0_index", "ref.fa")
    idx = capi.Index(prefix)
    reads = np.load(os.path.join(HERE, "c0_reads.npz"))["reads"]
    codes = reads.reshape(-1); offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    regs, ro, _, rc = ol.seed_chain_extend(idx, capi.default_opt(), codes, offs)
    assert rc == 0
    rng = np.random.default_rng(5)
    reqs = cu.make_requests(capi, rng, regs, ro, reads.shape[1], idx.desc.l_pac)
    # a fixture of a few thousand requests: every perturbed / degenerate one and a sample of the plain ones
    n_plain = 5 * len(regs)
    keep = np.concatenate([np.sort(rng.choice(n_plain, 2500, replace=False)), np.arange(n_plain, len(reqs))])
    reqs = reqs[keep]
    recs, cigar, md = cu.reference_gen_cigar(capi, prefix, codes, offs, reqs)
    np.savez_compressed(os.path.join(HERE, "cigar_c0.npz"), reqs=reqs, recs=recs, cigar=cigar, md=md)
    print(len(reqs), "requests,", int((recs["nm"] < 0).sum()), "rejected,", int((recs["n_cigar"] > 1).sum()), "with indels")


if __name__ == "__main__":
    main()
