"""Seams 0 and 5 on the GPU box: bm2_fastq_encode parses and encodes raw FASTQ bytes on the device; the whole flow
FASTQ bytes -> bm2_fastq_encode -> bm2_seed_chain_extend_resident -> bm2_pestat -> bm2_sam_pe -> bm2_sam_format must give the reference's
SAM text (tests/golden/c0.sam) byte for byte."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fastq(names, reads, quals, eol=b"\n", lower=False, last_eol=True):
    out = []
    for nm, r, q in zip(names, reads, quals):
        s = bytes(b"ACGTN"[c] for c in r)
        if lower:
            s = s.lower()
        out.append(b"@" + nm + eol + s + eol + b"+" + eol + q + eol)
    buf = b"".join(out)
    return buf if last_eol else buf[:-len(eol)]


def test_fastq_encode_matches_a_host_parse(pkg):
    capi = pkg.capi
    rng = np.random.default_rng(3)
    n = 257
    lens = rng.integers(1, 300, n)
    reads = [rng.integers(0, 5, L).astype(np.uint8) for L in lens]
    quals = [bytes(rng.integers(33, 74, L).astype(np.uint8)) for L in lens]
    n1 = [b"read%d/1 comment here" % i for i in range(n)]; n2 = [b"read%d/2\tmore" % i for i in range(n)]
    want_names = [b"read%d" % i for i in range(n)]
    ctx = capi.Context(0)
    for eol, lower, last in ((b"\n", False, True), (b"\r\n", True, True), (b"\n", False, False)):
        b1 = _fastq(n1, reads, quals, eol, lower, last); b2 = _fastq(n2, reads[::-1], quals[::-1], eol, lower, last)
        r = ctx.fastq_encode(b1, b2)
        assert r["n_reads"] == 2 * n
        off = r["offsets"]
        for i in range(n):
            for k, (rd, ql) in enumerate(((reads[i], quals[i]), (reads[n - 1 - i], quals[n - 1 - i]))):
                a, b = off[2 * i + k], off[2 * i + k + 1]
                assert np.array_equal(r["codes"][a:b], rd) and bytes(r["quals"][a:b]) == ql
            assert r["names"][2 * i] == want_names[i] and r["names"][2 * i + 1] == want_names[i]
        s = ctx.fastq_encode(b1)                                        # single-end
        assert s["n_reads"] == n and np.array_equal(np.diff(s["offsets"]), lens)
    with pytest.raises(capi.Bm2Error):
        ctx.fastq_encode(b"@a\nACGT\n+\nIII\n")                         # qualities shorter than the sequence
    with pytest.raises(capi.Bm2Error):
        ctx.fastq_encode(b"@a\nACGT\n+\nIIII\n@b\nAC\n")                # truncated record
    with pytest.raises(capi.Bm2Error):
        ctx.fastq_encode(b"@a\nACGT\n+\nIIII\n", b"@a\nACGT\n+\nIIII\n@b\nA\n+\nI\n")      # different record counts
    ctx.close()


@pytest.mark.parametrize("staged", [0, 1])
def test_fastq_bytes_to_sam_text_equals_the_reference(pkg, golden_dir, staged):
    capi = pkg.capi
    idx = capi.Index(golden_dir + "/c0_index/ref.fa")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    want = [ln for ln in open(golden_dir + "/c0.sam") if not ln.startswith("@")]
    contigs = [l.split()[1] for i, l in enumerate(open(golden_dir + "/c0_index/ref.fa.ann")) if i % 2 == 1]
    # QNAMEs and the quality character of the golden run (its FASTQ files are not kept): first line of every read
    first = {}
    for ln in want:
        f = ln.split("\t")
        flag = int(f[1])
        key = (f[0], 1 if flag & 0x80 else 0)
        first.setdefault(key, f)
    qnames = sorted({k[0] for k in first}, key=lambda s: int("".join(ch for ch in s if ch.isdigit())))
    assert len(qnames) == len(reads) // 2
    qc = want[0].split("\t")[10][0].encode()
    L = reads.shape[1]
    b1 = _fastq([q.encode() + b"/1" for q in qnames], reads[0::2], [qc * L] * len(qnames))
    b2 = _fastq([q.encode() + b"/2" for q in qnames], reads[1::2], [qc * L] * len(qnames))
    opt = capi.default_opt(); opt.flag |= 0x2
    ctx = capi.Context(0, index=idx, opt=opt)
    ctx.set_sam_staged(staged)
    fq = ctx.fastq_encode(b1, b2)
    assert np.array_equal(fq["codes"].reshape(-1, L), reads)
    regs, ro = ctx.seed_chain_extend_resident(fq["codes"], fq["offsets"], fq["d_codes"], fq["d_offsets"], True, return_arrays=True)
    pes = capi.pestat(opt, idx.desc.l_pac, regs, ro)
    recs, xa, cig, md = ctx.sam_pe(fq["codes"], fq["offsets"], regs, ro, pes)
    text = capi.sam_format(recs, xa, cig, md, fq["codes"], fq["offsets"], contigs, read_names=fq["names"], quals=fq["quals"], n_threads=3).decode()
    assert text == "".join(want)
    ctx.close(); idx.close()


@pytest.mark.parametrize("K,workers,gz", [(100_000_000, 2, False), (40_000, 1, False), (40_000, 2, True), (15_000, 3, False)])
def test_bm2_mem_program_prints_the_reference_sam(pkg, golden_dir, tmp_path, K, workers, gz):
    """The C++ host program over the C ABI (bwa-mem2_b200/tools/bm2_mem.cpp): FASTQ files in, SAM file out, chunked by -K like the reference's
    reader; against the unmodified reference run live with the same -K (several chunks: mem_pestat per chunk, id offsets across chunks).
    -p workers: chunks in flight at a time, each on its own context (bm2_create_sibling), output in chunk order."""
    import os, subprocess, importlib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tool = os.path.join(root, "bwa-mem2_b200", "bm2_mem")
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    drv = os.path.join(root, "oracle", "_ref", isa, "ref_driver")
    if not os.path.exists(tool) or not os.path.exists(drv):
        pytest.skip("bm2_mem / oracle/_ref not built")
    synth = importlib.import_module("bwa_mem2_b200.synth")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    r1, r2 = str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    synth.write_fastq(r1, reads[0::2], "p"); synth.write_fastq(r2, reads[1::2], "p")
    if gz:          # gzip input, as real read files come (both programs read it through zlib)
        import gzip, shutil
        for pth in (r1, r2):
            with open(pth, "rb") as fi, gzip.open(pth + ".gz", "wb", compresslevel=1) as fo:
                shutil.copyfileobj(fi, fo)
        r1, r2 = r1 + ".gz", r2 + ".gz"
    prefix = golden_dir + "/c0_index/ref.fa"
    out = str(tmp_path / "out.sam")
    o = subprocess.run([tool, "-t", "4", "-K", str(K), "-p", str(workers), "-o", out, prefix, r1, r2], capture_output=True, text=True, timeout=600)
    assert o.returncode == 0, o.stderr[-2000:]
    ref = subprocess.run([drv, "mem", "-t", "4", "-K", str(K), prefix, r1, r2], env=dict(os.environ, BM2_MODE="ref"), capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0
    got = [l for l in open(out).read().splitlines() if not l.startswith("@PG")]
    want = [l for l in ref.stdout.splitlines() if not l.startswith("@PG")]
    assert len(got) == len(want) and len(got) > 1000
    diff = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert diff == [], (len(diff), got[diff[0]], want[diff[0]])
    import json
    st = json.loads(o.stderr.strip().splitlines()[-1])
    assert st["reads"] == len(reads) and st["workers"] == workers and st["chunks"] == (1 if K > 1_000_000 else 4 if K == 40_000 else st["chunks"])
    assert K != 15_000 or st["chunks"] > 8
    assert st["chunk_done_s"] == sorted(st["chunk_done_s"]) and sum(st["chunk_reads"]) == len(reads)
