"""GPU parity of seam 2 through the C ABI: SMEMs, chains and final regs, bit-exact against the reference's
golden stage dumps (C0) and against the oracle on a second, larger seeded input and on ragged input."""
import os, subprocess, tempfile
import numpy as np
import pytest
import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def c0(pkg, golden_dir):
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    codes = reads.reshape(-1)
    offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    st = np.load(golden_dir + "/c0_stages.npz")
    ctx = pkg.capi.Context(0, index=idx)
    yield idx, ctx, codes, offs, st
    ctx.close(); idx.close()


def _norm(a):
    return a[np.lexsort((a["s"], a["l"], a["k"], a["n"], a["m"], a["rid"]))]


def test_smems_match_reference(c0):
    idx, ctx, codes, offs, st = c0
    sm, ro = ctx.collect_smems(codes, offs)
    b = st["smems"]
    assert len(sm) == len(b)
    # our order is the reference's final order (rid, m asc, n asc); the dump was taken before the per-read sort
    a2, b2 = _norm(sm), _norm(b)
    for f in ("rid", "m", "n", "k", "l", "s"):
        assert np.array_equal(a2[f], b2[f]), f
    key = (sm["rid"].astype(np.uint64) << np.uint64(32)) | (sm["m"].astype(np.uint64) << np.uint64(16)) | sm["n"].astype(np.uint64)
    assert np.all(np.diff(key.astype(np.int64)) >= 0)
    assert ro[-1] == len(sm)


def test_chains_match_reference(c0):
    idx, ctx, codes, offs, st = c0
    ch, sd, co = ctx.seed_chain(codes, offs)
    rc, rs = st["chains"], st["seeds"]
    assert np.array_equal(co, st["chain_off"])
    for f, g in (("pos", "pos"), ("rid", "rid"), ("n_seeds", "n"), ("w", "w"), ("kept", "kept"), ("first", "first"), ("frac_rep", "frac_rep"),
                 ("seed_off", "seed_off")):
        assert np.array_equal(ch[f], rc[g]), f
    for f in ("rbeg", "qbeg", "len", "score"):
        assert np.array_equal(sd[f], rs[f]), f


def test_regs_match_reference(c0):
    idx, ctx, codes, offs, st = c0
    regs, ro = ctx.seed_chain_extend(codes, offs)
    assert ol.regs_equal_to_dump(regs, ro, st["regs"], st["reg_off"]) == []
    ms = ctx.stage_ms()
    assert "smem" in ms and "bsw_left" in ms


def test_regs_match_reference_with_pair_bsw_kernel(c0, monkeypatch):
    # same batch with the experimental two-jobs-per-thread extension kernel routed in (off by default)
    monkeypatch.setenv("BM2_BSW_PAIR", "1")
    idx, ctx, codes, offs, st = c0
    regs, ro = ctx.seed_chain_extend(codes, offs)
    assert ol.regs_equal_to_dump(regs, ro, st["regs"], st["reg_off"]) == []


def test_regs_match_reference_with_cell_bsw_kernel(c0, monkeypatch):
    # same batch with the one-cell-per-instruction extension kernel for the 8-bit classes (the default is bsw_col2_kernel)
    monkeypatch.setenv("BM2_BSW_COL2", "0")
    idx, ctx, codes, offs, st = c0
    regs, ro = ctx.seed_chain_extend(codes, offs)
    assert ol.regs_equal_to_dump(regs, ro, st["regs"], st["reg_off"]) == []


@pytest.mark.parametrize("sc", [dict(zdrop=0), dict(a=2, b=8, o_del=12, e_del=2, o_ins=12, e_ins=2, zdrop=200, pen_clip5=10, pen_clip3=10, T=60, pen_unpaired=34)])
def test_regs_match_oracle_with_non_default_options(pkg, c0, sc):
    # -d 0 and -A 2 (update_a of src/fastmap.cpp scales B, O, E, L, T, d, U): the oracle is pinned to the reference run with these
    # options by tests/test_option_surface_cpu.py
    idx, ctx0, codes, offs, st = c0
    o = pkg.capi.default_opt()
    for k, v in sc.items():
        setattr(o, k, v)
    k = 0
    for i in range(4):
        for j in range(4):
            o.mat[k] = o.a if i == j else -o.b; k += 1
        o.mat[k] = -1; k += 1
    ctx = pkg.capi.Context(0, index=idx, opt=o)
    regs, ro = ctx.seed_chain_extend(codes, offs)
    want, wo, _, rc = ol.seed_chain_extend(idx, o, codes, offs)
    assert rc == 0 and np.array_equal(ro, wo) and regs.tobytes() == want.tobytes()
    ctx.close()


def test_ragged_and_degenerate_reads(pkg, c0):
    idx, ctx, codes, offs, st = c0
    reads = codes.reshape(-1, 151)
    parts = [reads[0][:0], reads[1][:10], np.full(60, 4, np.uint8), reads[2][:100], reads[3], np.concatenate([reads[4], reads[5][:70]])]
    c2 = np.concatenate(parts); o2 = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
    regs, ro = ctx.seed_chain_extend(c2, o2)
    want, wo, _, rc = ol.seed_chain_extend(idx, ctx.opt, c2, o2)
    assert np.array_equal(ro, wo)
    for f in ol.REG_CMP_FIELDS + ("n_comp_is_alt",):
        assert np.array_equal(regs[f], want[f]), f
    # empty batch
    regs, ro = ctx.seed_chain_extend(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert len(regs) == 0 and list(ro) == [0]


def test_second_dataset_against_oracle(pkg):
    """2 Mbp reference / 3000 pairs, index built by the reference binary when it is available."""
    isa = "avx512bw" if "avx512bw" in open("/proc/cpuinfo").read() else "avx2"
    refbin = os.path.join(ROOT, "oracle", "_ref", isa, "bwa-mem2")
    if not os.path.exists(refbin):
        pytest.skip("oracle/_ref not built")
    import importlib
    synth = importlib.import_module("bwa_mem2_b200.synth")
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
    want, wo, cells, rc = ol.seed_chain_extend(idx, ctx.opt, codes, offs)
    assert rc == 0 and np.array_equal(ro, wo) and len(regs) > 5000
    for f in ol.REG_CMP_FIELDS + ("n_comp_is_alt",):
        assert np.array_equal(regs[f], want[f]), f
    # idempotence: a second call on the same context gives the same bytes
    regs2, ro2 = ctx.seed_chain_extend(codes, offs)
    assert regs.tobytes() == regs2.tobytes() and np.array_equal(ro, ro2)
    # sub-batches in flight (bm2_set_sub_batches): 6000 reads as 4 / 3 / 11 concurrent sub-batches cut at multiples of 512
    # reads give the same bytes as the unsplit batch, through the host entry and through the device-resident entry
    import torch
    for k in (4, 3, 11):
        ctx.set_sub_batches(k, 512)
        regs3, ro3 = ctx.seed_chain_extend(codes, offs)
        assert regs.tobytes() == regs3.tobytes() and np.array_equal(ro, ro3), k
    d_codes = torch.from_numpy(codes).cuda(); d_offs = torch.from_numpy(offs).cuda()
    torch.cuda.synchronize()
    regs4, ro4 = ctx.seed_chain_extend_resident(codes, offs, d_codes.data_ptr(), d_offs.data_ptr(), True, return_arrays=True)
    assert regs.tobytes() == regs4.tobytes() and np.array_equal(ro, ro4)
    assert ctx.counters()["cells"] > 0 and ctx.stage_ms()["smem"] > 0
    ctx.set_sub_batches(1)
    ctx.close(); idx.close()
