#!/usr/bin/env python
"""One-off differential run (usage: bigdiff.py <ref Mbp> <pairs> <seed> [--len=N] [mem options ...])
One-off differential run (CPU only, needs oracle/_ref): a larger synthetic data set than the committed fixtures, optional
`bwa-mem2 mem` options; the UNMODIFIED reference (regs dumped by ref_driver's link-time hooks) against the oracle and against the
kernels' device logic (host emulation), every field of every alignment region.
Usage: bigdiff.py <ref Mbp> <pairs> <seed> [mem options ...]"""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
pkg = load_package(); capi = pkg.capi
import importlib
import oracle_lib as ol, emul_lib as el, refdump, cigar_util as cu
from test_option_surface_cpu import opt_from_cli


def main():
    mbp, pairs, seed = float(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]); args = sys.argv[4:]
    read_len = 151
    if args and args[0].startswith("--len="):            # read length of the synthetic pairs (default 151)
        read_len = int(args[0][6:]); args = args[1:]
    synth = importlib.import_module("bwa_mem2_b200.synth")
    work = tempfile.mkdtemp(prefix="bm2_big_")
    ctg = synth.make_reference(int(mbp * 1e6), seed=seed, n_contigs=5, repeat_frac=0.3)
    synth.write_fasta(work + "/ref.fa", ctg)
    r1, r2 = synth.make_pairs_fast(ctg, pairs, read_len=read_len, seed=seed + 1)
    synth.write_fastq_fast(work + "/r1.fq", r1); synth.write_fastq_fast(work + "/r2.fq", r2)
    bindir = os.path.dirname(cu.refbin())
    subprocess.check_call([bindir + "/bwa-mem2", "index", work + "/ref.fa"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    env = dict(os.environ, BM2_DUMP_PREFIX=work + "/d")
    t0 = time.time()
    subprocess.check_call([cu.refbin(), "mem", "-t", "1", "-K", "1000000000"] + args + [work + "/ref.fa", work + "/r1.fq", work + "/r2.fq"],
                          stdout=open(work + "/o.sam", "w"), stderr=subprocess.DEVNULL, env=env)
    rr, roff = refdump.read_regs(work + "/d.regs.bin")
    reads = np.empty((2 * pairs, r1.shape[1]), np.uint8); reads[0::2] = r1; reads[1::2] = r2
    codes = reads.reshape(-1); offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    idx = capi.Index(work + "/ref.fa"); opt = opt_from_cli(capi, args)
    t1 = time.time()
    regs, ro, cells, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    t2 = time.time()
    bad = ol.regs_equal_to_dump(regs, ro, rr, roff)
    e, eo = el.seed_chain_extend(idx, opt, codes, offs)
    t3 = time.time()
    # SAM level: the oracle's mate rescue + pairing + MAPQ + CIGAR/NM/MD against the reference's SAM lines
    import test_oracle_sam_pe as tp
    opt2 = opt_from_cli(capi, args); opt2.flag |= 0x2
    lh, as_ = tp._pestat(capi, idx, opt2, reads, regs, ro)
    recs, cig, md = tp.oracle_sam_pe(capi, idx, opt2, codes, offs, regs, ro, lh, as_)
    names = [l.split()[1] for i, l in enumerate(open(work + "/ref.fa.ann")) if i % 2 == 1]
    got = tp.fields(recs, cig, md, names); want = tp.parse_sam(open(work + "/o.sam"))
    sam_bad = sum(1 for g, w in zip(got, want) if g != w) + abs(len(got) - len(want))
    # ... and the complete text of every line after QNAME (SEQ / QUAL, all tags); write_fastq_fast writes quality 'I'
    tgot = tp.oracle_sam_text(capi, idx, opt2, codes, offs, regs, ro, lh, as_, names)
    twant = [ln.rstrip("\n").split("\t", 1)[1] for ln in open(work + "/o.sam") if not ln.startswith("@")]
    text_bad = sum(1 for g, w in zip(tgot, twant) if g != w) + abs(len(tgot) - len(twant))
    print(f"SAM lines {len(want)}, differing columns {sam_bad}, differing text {text_bad}, XA tags {sum('XA:Z:' in w for w in twant)}, SA tags {sum('SA:Z:' in w for w in twant)}")
    print(f"{mbp} Mbp, {2 * pairs} reads, options {args}: {len(rr)} regs; reference {t1 - t0:.0f}s oracle {t2 - t1:.0f}s emulation {t3 - t2:.0f}s; "
          f"oracle vs reference: {len(bad)} differing reads {bad[:5]}; device logic == oracle: {e.tobytes() == regs.tobytes() and np.array_equal(eo, ro)}")


if __name__ == "__main__":
    main()
