#!/usr/bin/env python
"""Fixture for chains with EQUAL positions (run where /root/reference exists and oracle/_ref is built; the output is committed).

A 56 kb text with short tandem repeats -- (AT)n, (CA)n, (CAG)n, a period-37 repeat -- and reads that lie inside them, cross
their borders or carry a few substitutions.  Such reads give seeds with the same reference start and different query starts,
which the reference's chain B-tree stores as equal keys (src/kbtree.h accepts them, src/bwamem.cpp:916-950); the outcome then
depends on the shape of that tree.  The UNMODIFIED reference is run on the reads (regs dumped by ref_driver's link-time hooks)
and the compared fields of every alignment region are stored.

Writes tests/golden/tandem_index/ (reference-built index), tandem_reads.npz, tandem_regs.npz."""
import os, subprocess, sys, tempfile
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import refdump, cigar_util as cu, oracle_lib as ol


def main():
    rng = np.random.default_rng(20240917)
    g = rng.integers(0, 4, 56_000).astype(np.uint8)
    regions = []
    def put(start, unit, n):
        u = np.array(unit, np.uint8); g[start:start + len(u) * n] = np.tile(u, n); regions.append((start, start + len(u) * n))
    put(20_000, [0, 3], 15_000)                # (AT)n, 30 kb: its own reverse complement; the max_occ = 500 sampled occurrences of a
                                               # piece lie 2 * (s / 500) = 118 bp apart, farther than the chaining band w = 100
    put(3_000, [1, 0], 700)                    # (CA)n
    put(7_000, [1, 0, 2], 400)                 # (CAG)n
    put(11_000, rng.integers(0, 4, 37), 50)    # period 37
    d = os.path.join(HERE, "tandem_index"); os.makedirs(d, exist_ok=True)
    fa = os.path.join(d, "ref.fa")
    with open(fa, "w") as f:
        s = "".join("ACGT"[b] for b in g)
        f.write(">tr1\n" + "\n".join(s[j:j + 60] for j in range(0, 14_000, 60)) + "\n")
        f.write(">tr2\n" + "\n".join(s[j:j + 60] for j in range(14_000, len(s), 60)) + "\n")
    bindir = os.path.dirname(cu.refbin())
    subprocess.check_call([bindir + "/bwa-mem2", "index", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.remove(fa)                               # the index is the fixture (ref.fa.0123 holds the text)
    comp = np.array([3, 2, 1, 0, 4], np.uint8)
    reads = []
    for (b, e) in regions:
        for L in (76, 151, 251):
            for rep in range(3):
                st = int(rng.integers(b, e - L)); r = g[st:st + L].copy()
                if rep == 1: r[rng.integers(0, L, 2)] = rng.integers(0, 4, 2)          # two substitutions
                if rep == 2: r = comp[r[::-1]]
                reads.append(r)
            st = int(rng.integers(b - L + 20, b - 10)); reads.append(g[st:st + L].copy())         # across the left border
            st = int(rng.integers(e - L + 10, e - 20)); reads.append(comp[g[st:st + L][::-1]])    # across the right border
    b, e = regions[0]                            # (AT)n reads cut into several long exact matches: the same suffix-array rows are
    for k in range(16):                          # sampled for each piece, so seeds of different pieces share their reference start
        L = 251 if k < 10 else 151
        st = int(rng.integers(b, e - L)); r = g[st:st + L].copy()
        cuts = [103, 171, 205, 227] if k == 0 else sorted(rng.choice(np.arange(20, L - 20), int(rng.integers(2, 6)), replace=False).tolist())
        for c in cuts: r[c] = (r[c] + int(rng.integers(1, 4))) & 3
        reads.append(comp[r[::-1]] if k % 3 == 2 else r)
    for _ in range(24):                          # ordinary reads
        L = int(rng.choice([76, 151, 251])); st = int(rng.integers(0, len(g) - L)); reads.append(g[st:st + L].copy())
    if len(reads) % 2: reads.append(reads[0].copy())
    reads = [np.ascontiguousarray(r, np.uint8) for r in reads]
    codes = np.concatenate(reads); offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "tandem_reads.npz"), codes=codes, offs=offs)
    work = tempfile.mkdtemp(prefix="bm2_tandem_")
    for k, name in ((0, "r1.fq"), (1, "r2.fq")):
        with open(os.path.join(work, name), "w") as f:
            for i, r in enumerate(reads[k::2]):
                f.write(f"@p{i}\n{''.join('ACGTN'[c] for c in r)}\n+\n{'I' * len(r)}\n")
    env = dict(os.environ, BM2_DUMP_PREFIX=work + "/d")
    subprocess.check_call([cu.refbin(), "mem", "-t", "1", "-K", "1000000000", fa, work + "/r1.fq", work + "/r2.fq"],
                          stdout=open(work + "/o.sam", "w"), stderr=subprocess.DEVNULL, env=env)
    rr, roff = refdump.read_regs(work + "/d.regs.bin")
    keep = np.zeros(len(rr), dtype=[(f, rr.dtype[f]) for f in ol.REG_CMP_FIELDS] + [("n_comp", rr.dtype["n_comp"]), ("is_alt", rr.dtype["is_alt"])])
    for f in keep.dtype.names: keep[f] = rr[f]
    np.savez_compressed(os.path.join(HERE, "tandem_regs.npz"), regs=keep, offs=roff)
    chains = refdump.read_chains(work + "/d.chains.bin")
    print(f"{len(reads)} reads, {len(rr)} regs (max per read {int(np.diff(roff).max())}), chains per read max {max(len(c) for c in chains)}")


if __name__ == "__main__":
    main()
