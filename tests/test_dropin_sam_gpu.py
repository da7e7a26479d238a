"""Drop-in check at the SAM level: the UNMODIFIED reference (oracle/_ref/*/ref_driver = reference main_mem +
link-time hooks) runs with its worker_bwt + worker_aln replaced by libbm2b200.so through the C ABI, keeps its own
mem_pestat + worker_sam, and must print the same SAM as the pure reference run (tests/golden/c0.sam)."""
import os, subprocess, tempfile
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _driver():
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    p = os.path.join(ROOT, "oracle", "_ref", isa, "ref_driver")
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built")
    return p


def _write_fastq(golden_dir, work):
    import importlib
    synth = importlib.import_module("bwa_mem2_b200.synth")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    synth.write_fastq(work + "/r1.fq", reads[0::2], "p"); synth.write_fastq(work + "/r2.fq", reads[1::2], "p")


def _sam_records(text):
    return [l for l in text.splitlines() if not l.startswith("@PG")]


@pytest.mark.parametrize("mode,threads", [("gpu", 1), ("gpu", 4), ("gpu_bsw", 2)])
def test_sam_identical_to_reference(pkg, golden_dir, mode, threads):
    drv = _driver()
    work = tempfile.mkdtemp(prefix="bm2_sam_")
    _write_fastq(golden_dir, work)
    env = dict(os.environ, BM2_MODE=mode, BM2_LIB=pkg.capi.LIB_PATH)
    out = subprocess.run([drv, "mem", "-t", str(threads), "-K", "100000000", golden_dir + "/c0_index/ref.fa", work + "/r1.fq", work + "/r2.fq"],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = _sam_records(out.stdout)
    want = _sam_records(open(golden_dir + "/c0.sam").read())
    assert len(got) == len(want) and len(got) > 1000
    diff = [i for i, (a, b) in enumerate(zip(got, want)) if a != b]
    assert diff == [], (len(diff), got[diff[0]], want[diff[0]])
