"""Pins the oracle's FM-index / chaining / extension stages to the UNMODIFIED reference: the golden
stage dumps in tests/golden/c0_stages.npz were written by oracle/_ref/*/ref_driver (link-time
hooks around the reference's own functions; tests/golden/make_golden.py)."""
import numpy as np
import pytest
import oracle_lib as ol


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


def test_smems_match_reference(c0):
    idx, opt, codes, offs, st = c0
    a = _norm(ol.collect_smems(idx, opt, codes, offs)); b = _norm(st["smems"])
    assert len(a) == len(b) and len(a) > 5000
    for f in ("rid", "m", "n", "k", "l", "s"):
        assert np.array_equal(a[f], b[f]), f


def test_sa_lookup_matches_seed_positions(c0):
    # SA values of the first row of every SMEM interval must be genuine occurrences of the read substring
    idx, opt, codes, offs, st = c0
    import ctypes as C
    sm = st["smems"]
    sa = ol.sa_lookup(idx, sm["k"])
    d = idx.desc
    ref = np.ctypeslib.as_array(C.cast(d.ref_string, C.POINTER(C.c_uint8)), shape=(2 * d.l_pac,))
    for i in range(0, len(sm), 37):
        q = codes[offs[sm["rid"][i]] + sm["m"][i]: offs[sm["rid"][i]] + sm["n"][i] + 1]
        assert np.array_equal(ref[sa[i]:sa[i] + len(q)], q)


def test_chains_match_reference(c0):
    idx, opt, codes, offs, st = c0
    ch, sd, co = ol.seed_chain(idx, opt, codes, offs)
    rc, rs, ro = st["chains"], st["seeds"], st["chain_off"]
    assert np.array_equal(co, ro)
    for f, g in (("pos", "pos"), ("rid", "rid"), ("n_seeds", "n"), ("w", "w"), ("kept", "kept"), ("first", "first"),
                 ("frac_rep", "frac_rep")):
        assert np.array_equal(ch[f], rc[g]), f
    assert np.array_equal(ch["seqid"] % 512, rc["seqid"])      # the reference numbers reads within a 512-read block
    assert np.array_equal(ch["seed_off"], rc["seed_off"])
    for f in ("rbeg", "qbeg", "len", "score"):
        assert np.array_equal(sd[f], rs[f]), f


def test_regs_match_reference(c0):
    idx, opt, codes, offs, st = c0
    regs, ro, cells, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    assert rc == 0 and cells > 0
    assert ol.regs_equal_to_dump(regs, ro, st["regs"], st["reg_off"]) == []
