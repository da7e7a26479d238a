"""The per-thread device logic of the seam-2 kernels (bwa-mem2_b200/csrc/{fm,chain,ext}_device.cuh),
compiled for the host by tests/host_emul, against the reference's golden stage dumps.  This is how the
kernels' control logic is checked where no GPU exists; `-m gpu` tests check the real kernels."""
import numpy as np
import pytest
import oracle_lib as ol
import emul_lib as el


@pytest.fixture(scope="module")
def c0(pkg, golden_dir):
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa")
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    codes = reads.reshape(-1)
    offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    st = np.load(golden_dir + "/c0_stages.npz")
    yield idx, pkg.capi.default_opt(), codes, offs, st
    idx.close()


def _norm(a):
    return a[np.lexsort((a["s"], a["l"], a["k"], a["n"], a["m"], a["rid"]))]


def test_smem_state_machine(c0):
    idx, opt, codes, offs, st = c0
    sm, n_ext = el.collect_smems(idx, opt, codes, offs)
    a, b = _norm(sm), _norm(st["smems"])
    assert len(a) == len(b)
    for f in ("rid", "m", "n", "k", "l", "s"):
        assert np.array_equal(a[f], b[f]), f
    assert 300 < n_ext / (len(offs) - 1) < 1000      # interval extensions per read (~628)


def test_chain_logic(c0):
    idx, opt, codes, offs, st = c0
    ch, sd, co = el.seed_chain(idx, opt, codes, offs)
    rc, rs = st["chains"], st["seeds"]
    assert np.array_equal(co, st["chain_off"])
    for f, g in (("pos", "pos"), ("rid", "rid"), ("n_seeds", "n"), ("w", "w"), ("kept", "kept"), ("first", "first"), ("frac_rep", "frac_rep")):
        assert np.array_equal(ch[f], rc[g]), f
    for f in ("rbeg", "qbeg", "len", "score"):
        assert np.array_equal(sd[f], rs[f]), f


def test_extension_logic(c0):
    idx, opt, codes, offs, st = c0
    regs, ro = el.seed_chain_extend(idx, opt, codes, offs)
    assert ol.regs_equal_to_dump(regs, ro, st["regs"], st["reg_off"]) == []


def test_ragged_and_degenerate_reads(pkg, golden_dir):
    # empty read, read shorter than the seed length, all-N read, reads of different lengths
    idx = pkg.capi.Index(golden_dir + "/c0_index/ref.fa")
    opt = pkg.capi.default_opt()
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    parts = [reads[0][:0], reads[1][:10], np.full(60, 4, np.uint8), reads[2][:100], reads[3], np.concatenate([reads[4], reads[5][:70]])]
    codes = np.concatenate(parts); offs = np.concatenate([[0], np.cumsum([len(p) for p in parts])]).astype(np.int64)
    regs, ro = el.seed_chain_extend(idx, opt, codes, offs)
    want, wo, _, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    assert rc == 0 and np.array_equal(ro, wo)
    for f in ol.REG_CMP_FIELDS + ("n_comp_is_alt",):
        assert np.array_equal(regs[f], want[f]), f
    assert ro[1] == 0 and ro[2] == 0 and ro[3] == 0     # nothing for empty / too short / all-N reads
    idx.close()


def test_unique_interval_text_shortcut_changes_only_l(pkg, golden_dir):
    """fm_forward with the reference text (the whole-path entries' setting): a one-row interval is extended by comparing the text with the read
    instead of by Occ lookups.  Every SMEM must keep rid, m, n, k, s (l is not maintained in that mode and not read by anything downstream), the
    search must make the same number of extensions, and the alignment regions of the whole path must stay the oracle's."""
    import emul_lib as el, oracle_lib as ol
    capi = pkg.capi
    idx = capi.Index(golden_dir + "/c0_index/ref.fa"); opt = capi.default_opt()
    reads = np.load(golden_dir + "/c0_reads.npz")["reads"]
    codes = reads.reshape(-1); offs = (np.arange(len(reads) + 1) * reads.shape[1]).astype(np.int64)
    base = el.collect_smems(idx, opt, codes, offs)
    el.lib().emul_set_smem_text(1)
    try:
        fast = el.collect_smems(idx, opt, codes, offs)
        regs, ro = el.seed_chain_extend(idx, opt, codes, offs)
    finally:
        el.lib().emul_set_smem_text(0)
    sm0, n0 = base[0], base[-1]; sm1, n1 = fast[0], fast[-1]
    assert len(sm0) == len(sm1) and n0 == n1
    for f in ("rid", "m", "n", "k", "s"):
        assert np.array_equal(sm0[f], sm1[f]), f
    assert (sm0["l"] != sm1["l"]).sum() > 100                      # the shortcut really ran
    want, wo, _, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    assert rc == 0 and np.array_equal(ro, wo) and regs.tobytes() == want.tobytes()
