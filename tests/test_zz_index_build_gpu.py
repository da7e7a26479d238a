"""The GPU index builder of the bench (bwa_mem2_b200.index_build: the ~3 Gbp index of the default workload is its output) against the
unmodified reference's `bwa-mem2 index` at 100 Mbp ON THE GPU: .bwt.2bit.64 (Occ checkpoints, sampled SA, sentinel), .0123 and .pac must
be the same bytes.  (tests/test_index_build.py compares on the CPU device at the size of the committed golden set.)"""
import os, subprocess, tempfile
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_built_index_equals_reference_index_at_100mbp(pkg):
    import importlib, torch
    synth = importlib.import_module("bwa_mem2_b200.synth"); ib = importlib.import_module("bwa_mem2_b200.index_build")
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    refbin = os.path.join(ROOT, "oracle", "_ref", isa, "bwa-mem2")
    if not os.path.exists(refbin):
        pytest.skip("oracle/_ref not built")
    work = tempfile.mkdtemp(prefix="bm2_idx100_")
    ctg = synth.make_reference(100_000_000, seed=55, n_contigs=6)
    synth.write_fasta(work + "/ref.fa", ctg)
    subprocess.check_call([refbin, "index", work + "/ref.fa"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1500)
    # the contigs as the reference's index holds them (its .0123: ambiguous bases, if any, already replaced by its random bases)
    g = work + "/ref.fa"
    ann = open(g + ".ann").read().split("\n")
    l_pac, n_seqs = int(ann[0].split()[0]), int(ann[0].split()[1])
    fwd = np.fromfile(g + ".0123", np.uint8)[:l_pac]
    contigs = []
    for i in range(n_seqs):
        off, ln = (int(x) for x in ann[2 + 2 * i].split()[:2])
        contigs.append((ann[1 + 2 * i].split()[1], torch.from_numpy(fwd[off:off + ln].copy()).cuda()))
    ib.write_index(work + "/x", contigs, device="cuda")
    for suf in (".bwt.2bit.64", ".0123", ".pac"):
        a = np.fromfile(work + "/x" + suf, np.uint8); b = np.fromfile(work + "/ref.fa" + suf, np.uint8)
        assert a.shape == b.shape and np.array_equal(a, b), suf
    for f in os.listdir(work):
        os.remove(os.path.join(work, f))
