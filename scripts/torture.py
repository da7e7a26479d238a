#!/usr/bin/env python
"""One-off differential run on hostile input (CPU only, needs oracle/_ref): many tiny contigs, N runs, microsatellites and tandem repeats in the
reference; reads of mixed lengths (25-251), homopolymers, dinucleotide repeats, N-rich reads, reads spanning contig ends.  The UNMODIFIED reference
(regs dumped by ref_driver's hooks, SAM) against the oracle (regs, SAM text) and the kernels' device logic (host emulation).
Usage: torture.py <seed> [mem options ...]      (BM2_TORTURE_LONG=1: a tenth of the pairs are 0.8-3 kb reads, for mem_flt_chained_seeds)"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
capi = load_package().capi
import oracle_lib as ol, emul_lib as el, refdump, cigar_util as cu
from test_option_surface_cpu import opt_from_cli
import test_oracle_sam_pe as tp


def main():
    seed = int(sys.argv[1]); args = sys.argv[2:]
    rng = np.random.default_rng(seed)
    work = tempfile.mkdtemp(prefix="bm2_tort_")
    ctgs = []
    for c in range(60):                                       # tiny contigs
        L = int(rng.integers(300, 6000)); g = rng.integers(0, 4, L).astype(np.uint8); ctgs.append(g)
    big = rng.integers(0, 4, 600_000).astype(np.uint8)
    big[100_000:130_000] = 0                                   # poly-A
    big[200_000:230_000] = np.tile([0, 3], 15_000)             # (AT)n
    unit = rng.integers(0, 4, 37).astype(np.uint8); big[300_000:300_000 + 37 * 600] = np.tile(unit, 600)      # tandem repeat, period 37
    seg = big[400_000:401_000].copy()
    for k in range(40): big[410_000 + 2000 * k: 411_000 + 2000 * k] = seg                                      # 40 identical copies (max_occ)
    ctgs.append(big)
    with open(work + "/ref.fa", "w") as f:
        for i, g in enumerate(ctgs):
            s = "".join("ACGT"[b] for b in g)
            if i % 7 == 3 and len(s) > 500: s = s[:200] + "N" * int(rng.integers(1, 120)) + s[200:]             # N runs (.amb holes)
            f.write(f">t{i}\n" + "\n".join(s[j:j + 70] for j in range(0, len(s), 70)) + "\n")
    bindir = os.path.dirname(cu.refbin())
    subprocess.check_call([bindir + "/bwa-mem2", "index", work + "/ref.fa"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    idx = capi.Index(work + "/ref.fa")
    l_pac = idx.desc.l_pac
    import ctypes as C
    ref = np.ctypeslib.as_array(C.cast(idx.desc.ref_string, C.POINTER(C.c_uint8)), shape=(2 * l_pac,))
    comp = np.array([3, 2, 1, 0, 4], np.uint8)
    reads = []
    n_pairs = int(os.environ.get("BM2_TORTURE_PAIRS", "4000"))
    for p in range(n_pairs):
        L1 = int(rng.choice([25, 36, 50, 76, 101, 151, 151, 151, 200, 251])); L2 = int(rng.choice([25, 50, 76, 151, 151, 251]))
        if os.environ.get("BM2_TORTURE_LONG") and p % 10 == 0: L1 = int(rng.integers(800, 3000)); L2 = int(rng.integers(800, 3000))
        kind = rng.random()
        if kind < 0.70:                                        # a fragment from the text (may span contig ends / N-filled holes)
            ins = int(max(rng.normal(350, 60), max(L1, L2) + 5)); ins = min(ins, int(l_pac) - 10); st = int(rng.integers(0, l_pac - ins))
            frag = ref[st:st + ins].copy(); mut = rng.random(ins) < 0.015; frag[mut] = rng.integers(0, 4, int(mut.sum()))
            if rng.random() < 0.2:
                q = int(rng.integers(5, ins - 5)); d = int(rng.integers(1, 12)); frag = np.concatenate([frag[:q], frag[q + d:]]) if rng.random() < .5 else np.concatenate([frag[:q], rng.integers(0, 4, d).astype(np.uint8), frag[q:]])
            r1 = frag[:L1]; r2 = comp[frag[-L2:][::-1]]
        elif kind < 0.78: r1 = np.full(L1, int(rng.integers(0, 4)), np.uint8); r2 = np.full(L2, int(rng.integers(0, 4)), np.uint8)       # homopolymers
        elif kind < 0.86: r1 = np.tile(rng.integers(0, 4, 2), L1)[:L1].astype(np.uint8); r2 = np.tile(unit, 8)[:L2].astype(np.uint8)         # repeats
        elif kind < 0.93: r1 = rng.integers(0, 4, L1).astype(np.uint8); r2 = rng.integers(0, 4, L2).astype(np.uint8)                          # garbage
        else:
            st = int(rng.integers(0, l_pac - 300)); r1 = ref[st:st + L1].copy(); r2 = comp[ref[st + 100:st + 100 + L2][::-1]]
            r1[rng.random(len(r1)) < 0.3] = 4; r2[rng.random(len(r2)) < 0.05] = 4                                                               # N-rich
        if len(r1) < L1: r1 = np.concatenate([r1, rng.integers(0, 4, L1 - len(r1)).astype(np.uint8)])
        if len(r2) < L2: r2 = np.concatenate([r2, rng.integers(0, 4, L2 - len(r2)).astype(np.uint8)])
        if rng.random() < 0.5: r1, r2 = r2, r1
        reads += [np.ascontiguousarray(r1, np.uint8), np.ascontiguousarray(r2, np.uint8)]
    for k, name in ((0, "r1.fq"), (1, "r2.fq")):
        with open(os.path.join(work, name), "w") as f:
            for i, r in enumerate(reads[k::2]):
                f.write(f"@p{i}\n{''.join('ACGTN'[c] for c in r)}\n+\n{'I' * len(r)}\n")
    env = dict(os.environ, BM2_DUMP_PREFIX=work + "/d")
    subprocess.check_call([cu.refbin(), "mem", "-t", "1", "-K", "1000000000"] + args + [work + "/ref.fa", work + "/r1.fq", work + "/r2.fq"],
                          stdout=open(work + "/o.sam", "w"), stderr=subprocess.DEVNULL, env=env)
    rr, roff = refdump.read_regs(work + "/d.regs.bin")
    codes = np.concatenate(reads); offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    opt = opt_from_cli(capi, args)
    regs, ro, cells, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    bad = ol.regs_equal_to_dump(regs, ro, rr, roff)
    e, eo = el.seed_chain_extend(idx, opt, codes, offs)
    opt2 = opt_from_cli(capi, args); opt2.flag |= 0x2
    class R:                                                   # _pestat only needs len(reads)
        def __len__(self): return len(reads)
    lh, as_ = tp._pestat(capi, idx, opt2, R(), regs, ro)
    pp = capi.pestat(opt2, l_pac, regs, ro)                    # the product's host-side mem_pestat against the oracle's (pinned to the reference's dump)
    pes_ok = all((int(pp[d]["low"]), int(pp[d]["high"]), int(pp[d]["failed"])) == tuple(int(v) for v in lh[3 * d:3 * d + 3]) and
                 (pp[d]["failed"] or (pp[d]["avg"] == as_[2 * d] and pp[d]["std"] == as_[2 * d + 1])) for d in range(4))
    names = [l.split()[1] for i, l in enumerate(open(work + "/ref.fa.ann")) if i % 2 == 1]
    tgot = tp.oracle_sam_text(capi, idx, opt2, codes, offs, regs, ro, lh, as_, names)
    twant = [ln.rstrip("\n").split("\t", 1)[1] for ln in open(work + "/o.sam") if not ln.startswith("@")]
    tbad = [i for i in range(min(len(tgot), len(twant))) if tgot[i] != twant[i]]
    try:                                                       # the SAM stage's device logic (mate rescue, pairing, MAPQ, CIGAR / NM / MD) on the same input
        em = tp.emul_sam_pe(capi, idx, opt2, codes, offs, regs, ro, lh, as_, xa_names=names)
        tp._compare(tp.fields(*em[:3], names), tp.fields(*tp.oracle_sam_pe(capi, idx, opt2, codes, offs, regs, ro, lh, as_), names))
        assert em[3] == tp.xa_of_lines(twant), "XA entries differ from the reference's tags"
        sam_dev = "== oracle, XA == reference"
        import ctypes as C, test_oracle_sam_se as ts
        L = tp._emul()
        for mode, what in ((1, "warp per window"), (2, "thread per window")):      # the staged rescue: job table + batch of local alignments + lookup
            L.emul_sam_set_staged(mode)
            try:
                em2 = tp.emul_sam_pe(capi, idx, opt2, codes, offs, regs, ro, lh, as_)
                st = (C.c_longlong * 4)(); L.emul_sam_stage_stats(st)
            finally:
                L.emul_sam_set_staged(0)
            assert all(np.array_equal(x, y) for x, y in zip(em2[:3], em[:3])), "staged rescue differs"
            sam_dev += f"; staged rescue ({what}) identical (batch {st[0]}, looked up {st[1]}, computed in place {st[2]}, of those window moved {st[3]})"
        al, oc, om = ts.oracle_sam_se(capi, idx, opt, codes, offs, regs, ro)                      # every read as a single-end read
        assert ts.rec_fields(*ts.emul_sam_se(capi, idx, opt, codes, offs, regs, ro), names) == ts.sam_fields(al, oc, om, names, soft_clip_all=bool(opt.flag & 0x200)), "single-end differs"
        sam_dev += "; single-end device logic == oracle"
    except AssertionError as ex:
        sam_dev = "DIFFERS: " + str(ex)[:300]
    print(f"seed {seed} {args}: {len(reads)} reads, {len(rr)} regs (max per read {int(np.diff(roff).max())}); oracle vs reference: {len(bad)} differing reads {bad[:5]}; "
          f"device logic == oracle: {e.tobytes() == regs.tobytes() and np.array_equal(eo, ro)}; SAM lines {len(twant)} vs {len(tgot)}, differing text {len(tbad)} {tbad[:3]}; SAM-stage device logic {sam_dev}; bm2_pestat == oracle: {pes_ok}")
    for i in tbad[:2]:
        g = tgot[i].split("\t"); w = twant[i].split("\t")
        for a, b in zip(g, w):
            if a != b: print("   ours:", a[:200], "| ref:", b[:200])


if __name__ == "__main__":
    main()
