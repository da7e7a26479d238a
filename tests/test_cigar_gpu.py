"""GPU parity of seam 3 (bm2_gen_cigar == bwa_gen_cigar2: CIGAR, NM, MD) through the C ABI: against the golden vectors made by
the UNMODIFIED reference and against the oracle on larger / longer / non-default request sets."""
import numpy as np
import pytest
import oracle_lib as ol
import cigar_util as cu

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c0(pkg, golden_dir):
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    codes = reads.reshape(-1); offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    ctx = pkg.capi.Context(0, index=idx)
    yield idx, ctx, codes, offs, reads.shape[1]
    ctx.close(); idx.close()


def test_golden_reference_vectors(pkg, c0, golden_dir):
    idx, ctx, codes, offs, _ = c0
    g = np.load(golden_dir + "/cigar_c0.npz")
    got = ctx.gen_cigar(codes, offs, g["reqs"])
    assert cu.same(got, (g["recs"], g["cigar"], g["md"])) == []


def test_all_final_alignments_match_oracle(pkg, c0):
    idx, ctx, codes, offs, read_len = c0
    regs, ro = ctx.seed_chain_extend(codes, offs)
    reqs = cu.make_requests(pkg.capi, np.random.default_rng(41), regs, ro, read_len, idx.desc.l_pac, n_extra=4000)
    want = ol.gen_cigar(idx, ctx.opt, codes, offs, reqs)
    assert want[3] == 0
    got = ctx.gen_cigar(codes, offs, reqs)
    assert cu.same(got, want[:3]) == []
    # idempotence and an empty batch
    assert cu.same(ctx.gen_cigar(codes, offs, reqs[:777]), ol.gen_cigar(idx, ctx.opt, codes, offs, reqs[:777])[:3]) == []
    r0 = ctx.gen_cigar(codes, offs, reqs[:0])
    assert len(r0[0]) == 0 and len(r0[1]) == 0 and len(r0[2]) == 0


def test_long_alignments_and_non_default_scoring(pkg, golden_dir):
    # 2-3 kbp queries cut from the reference with substitutions and indels (wide bands, backtrack matrices of megabytes), -x ont2d scoring
    capi = pkg.capi
    idx = capi.Index(golden_dir + "/c0_index/ref.fa")
    o = capi.default_opt()
    o.o_del = o.o_ins = 1; o.e_del = o.e_ins = 1; o.b = 1
    k = 0
    for i in range(4):
        for j in range(4):
            o.mat[k] = o.a if i == j else -o.b; k += 1
        o.mat[k] = -1; k += 1
    ctx = capi.Context(0, index=idx, opt=o)
    l_pac = idx.desc.l_pac
    rng = np.random.default_rng(8)
    import ctypes as C
    ref = np.ctypeslib.as_array(C.cast(idx.desc.ref_string, C.POINTER(C.c_uint8)), shape=(2 * l_pac,))
    reads = []; reqs = []
    for i in range(24):
        L = int(rng.integers(1500, 3000)); rb = int(rng.integers(0, 2 * l_pac - L - 50))
        if rb < l_pac < rb + L + 40:
            rb = l_pac + 10
        t = ref[rb:rb + L].copy()
        mut = rng.random(L) < 0.05; t[mut] = rng.integers(0, 4, int(mut.sum()))
        for _ in range(6):      # a few indels
            p = int(rng.integers(10, len(t) - 10)); d = int(rng.integers(1, 8))
            t = np.concatenate([t[:p], t[p + d:]]) if rng.random() < 0.5 else np.concatenate([t[:p], rng.integers(0, 4, d).astype(np.uint8), t[p:]])
        reads.append(t.astype(np.uint8))
        reqs.append((rb, rb + L, i, 0, len(t), int(rng.choice([50, 200, 400]))))
    codes = np.concatenate(reads); offs = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    reqs = np.array(reqs, dtype=capi.CIGAR_REQ_DT)
    want = ol.gen_cigar(idx, o, codes, offs, reqs)
    assert want[3] == 0 and (want[0]["n_cigar"] > 3).all()
    got = ctx.gen_cigar(codes, offs, reqs)
    assert cu.same(got, want[:3]) == []
    ctx.close(); idx.close()


def test_bad_requests_are_errors(pkg, c0):
    idx, ctx, codes, offs, read_len = c0
    bad = np.zeros(1, pkg.capi.CIGAR_REQ_DT)
    bad["read"] = 10 ** 6; bad["qe"] = 10; bad["re"] = 10
    with pytest.raises(pkg.capi.Bm2Error):
        ctx.gen_cigar(codes, offs, bad)
    bad["read"] = 0; bad["qe"] = read_len + 1
    with pytest.raises(pkg.capi.Bm2Error):
        ctx.gen_cigar(codes, offs, bad)
