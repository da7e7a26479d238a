"""The torch index builder (bench/test tooling for the 3 Gbp configurations) must write the same bytes as the
reference binary's `bwa-mem2 index` (golden: tests/golden/c0_index, built by the unmodified reference)."""
import os, tempfile
import numpy as np
import torch


def test_index_files_identical_to_reference(pkg, golden_dir):
    import importlib
    ib = importlib.import_module("bwa_mem2_b200.index_build")
    g = golden_dir + "/c0_index/ref.fa"
    ann = open(g + ".ann").read().split("\n")
    l_pac, n_seqs = int(ann[0].split()[0]), int(ann[0].split()[1])
    t0123 = np.fromfile(g + ".0123", np.uint8)
    fwd = t0123[:l_pac]
    contigs = []
    for i in range(n_seqs):
        name = ann[1 + 2 * i].split()[1]
        off, ln = (int(x) for x in ann[2 + 2 * i].split()[:2])
        contigs.append((name, fwd[off:off + ln]))
    work = tempfile.mkdtemp(prefix="bm2_idx_")
    ib.write_index(work + "/x", contigs, device="cpu")
    for suf in (".bwt.2bit.64", ".0123", ".pac"):
        a = open(work + "/x" + suf, "rb").read(); b = open(g + suf, "rb").read()
        assert a == b, suf
    # .ann differs only in n_ambs bookkeeping of the original FASTA (N runs were already replaced in .0123)
    mine = open(work + "/x.ann").read().split("\n")
    assert mine[0] == ann[0] and [l.split()[:2] for l in mine[2::2] if l] == [l.split()[:2] for l in ann[2::2] if l]


def test_suffix_array_against_naive():
    import importlib
    ib = importlib.import_module("bwa_mem2_b200.index_build")
    rng = np.random.default_rng(3)
    # low-entropy text with long repeats: forces many refinement rounds and end-of-text ties
    unit = rng.integers(0, 4, 50, dtype=np.uint8)
    t = np.concatenate([np.tile(unit, 40), rng.integers(0, 2, 500, dtype=np.uint8), np.zeros(70, np.uint8), np.tile(unit, 7), np.zeros(40, np.uint8)])
    sa = ib.suffix_array(torch.from_numpy(t), max_bucket=256).numpy()
    s = bytes(t + 1)
    naive = sorted(range(len(t)), key=lambda i: s[i:])
    assert list(sa) == naive
