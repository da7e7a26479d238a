"""Debug aid: where do sub-batch results differ from the unsplit batch? (second dataset of tests/test_pipeline_gpu.py)"""
import os, sys, subprocess, tempfile, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import oracle_lib as ol
pkg = load_package()
synth = importlib.import_module("bwa_mem2_b200.synth")
refbin = os.path.join(ROOT, "oracle", "_ref", "avx512bw", "bwa-mem2")
work = tempfile.mkdtemp(prefix="bm2_t2_")
ctg = synth.make_reference(2_000_000, seed=5, n_contigs=3)
synth.write_fasta(work + "/ref.fa", ctg)
subprocess.check_call([refbin, "index", work + "/ref.fa"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
r1, r2 = synth.make_pairs(ctg, 3000, seed=6)
reads = np.empty((6000, 151), np.uint8); reads[0::2] = r1; reads[1::2] = r2
codes = reads.reshape(-1); offs = (np.arange(6001) * 151).astype(np.int64)
idx = pkg.capi.Index(work + "/ref.fa")
ctx = pkg.capi.Context(0, index=idx)
regs, ro = ctx.seed_chain_extend(codes, offs)
print("unsplit regs", len(regs))


def diff(tag, regs3, ro3):
    same = regs.tobytes() == regs3.tobytes() and np.array_equal(ro, ro3)
    print(tag, "same" if same else "DIFFERENT", len(regs3))
    if same:
        return
    bad = [r for r in range(6000) if ro3[r + 1] - ro3[r] != ro[r + 1] - ro[r] or regs3[ro3[r]:ro3[r + 1]].tobytes() != regs[ro[r]:ro[r + 1]].tobytes()]
    print("   differing reads:", len(bad), bad[:20])
    r = bad[0]
    a = regs[ro[r]:ro[r + 1]]; b = regs3[ro3[r]:ro3[r + 1]]
    print("   read", r, "n", len(a), len(b))
    for f in a.dtype.names:
        if len(a) == len(b) and not np.array_equal(a[f], b[f]):
            print("     field", f, a[f][:6], b[f][:6])


for serial in ("1", None):
    if serial:
        os.environ["BM2_SUB_BATCHES_SERIAL"] = "1"
    else:
        os.environ.pop("BM2_SUB_BATCHES_SERIAL", None)
    for k in (2, 4, 3, 11, 4):
        ctx.set_sub_batches(k, 512)
        regs3, ro3 = ctx.seed_chain_extend(codes, offs)
        diff(f"serial={serial} k={k}", regs3, ro3)
ctx.set_sub_batches(1)
regs5, ro5 = ctx.seed_chain_extend(codes, offs)
diff("unsplit again", regs5, ro5)
