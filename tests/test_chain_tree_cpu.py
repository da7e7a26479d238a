"""Chains with EQUAL positions (reads inside short tandem repeats).  The reference keeps a read's chains in a B-tree that accepts equal
keys (src/kbtree.h; src/bwamem.cpp:916-950), so which of them a later seed is tested against depends on the shape of the tree.  The
oracle restates the tree (ChainTree in oracle/bm2_oracle.cpp), the kernels keep its shape as a level per key of the ordered array
(chain_tree_put_d / chain_tree_equal_d in chain_device.cuh).  Golden: the UNMODIFIED reference's alignment regions on
tests/golden/tandem_* (tests/golden/make_tandem_golden.py)."""
import numpy as np
import pytest
import oracle_lib as ol
import emul_lib as el


@pytest.fixture(scope="module")
def tandem(pkg, golden_dir):
    idx = pkg.capi.Index(golden_dir + "/tandem_index/ref.fa")
    rd = np.load(golden_dir + "/tandem_reads.npz"); gd = np.load(golden_dir + "/tandem_regs.npz")
    yield idx, rd["codes"], rd["offs"], gd["regs"], gd["offs"]
    idx.close()


def test_oracle_matches_reference_regs(pkg, tandem):
    idx, codes, offs, gregs, goffs = tandem
    regs, ro, cells, rc = ol.seed_chain_extend(idx, pkg.capi.default_opt(), codes, offs)
    assert rc == 0
    assert ol.regs_equal_to_dump(regs, ro, gregs, goffs) == []
    assert int(np.diff(ro).max()) >= 400            # the repeats are there


def test_equal_positions_occur(pkg, tandem):
    """The fixture must exercise the case: chains of one read with the same pos."""
    idx, codes, offs, _, _ = tandem
    ch, sd, co = ol.seed_chain(idx, pkg.capi.default_opt(), codes, offs)
    n_dup = 0
    for r in range(len(co) - 1):
        p = np.sort(ch["pos"][co[r]:co[r + 1]])
        n_dup += int((p[1:] == p[:-1]).sum())
    assert n_dup > 0


def test_device_logic_matches_oracle(pkg, tandem):
    idx, codes, offs, _, _ = tandem
    opt = pkg.capi.default_opt()
    regs, ro, cells, rc = ol.seed_chain_extend(idx, opt, codes, offs)
    eregs, ero = el.seed_chain_extend(idx, opt, codes, offs)
    assert np.array_equal(ero, ro) and eregs.tobytes() == regs.tobytes()


@pytest.mark.parametrize("n", [9, 10, 49, 50, 250, 3000])
def test_level_array_is_the_tree(n):
    """The level-per-key array against a node-based B-tree (kbtree's insertion and lookup rules restated in Python), random keys
    from a small range so that equal keys abound.  n around 9, 49, ...: the root splits for the first, second, ... time."""
    lib = el.lib()
    import ctypes as C
    rng = np.random.default_rng(n)
    keys = rng.integers(0, max(3, n // 6), n).astype(np.int64)
    T = 5

    class Node:
        def __init__(s, internal): s.internal = internal; s.key = []; s.child = []
    pos = {}
    root = [Node(False)]

    def slot(x, k):
        ks = [pos[i] for i in x.key]
        if not ks: return -1, 0
        b = int(np.searchsorted(ks, k, side="left"))
        if b == len(ks): return len(ks) - 1, 1
        return (b - 1, -1) if k < ks[b] else (b, 0)

    def split(x, i):
        y = x.child[i]; z = Node(y.internal)
        z.key = y.key[T:]; z.child = y.child[T:] if y.internal else []
        med = y.key[T - 1]; y.key = y.key[:T - 1]
        if y.internal: y.child = y.child[:T]
        x.child.insert(i + 1, z); x.key.insert(i, med)

    def lower(k):
        x = root[0]; low = -1
        while True:
            i, r = slot(x, k)
            if i >= 0 and r == 0: return x.key[i]
            if i >= 0: low = x.key[i]
            if not x.internal: return low
            x = x.child[i + 1]

    def put(ident):
        k = pos[ident]
        if len(root[0].key) == 2 * T - 1:
            s = Node(True); s.child = [root[0]]; root[0] = s; split(s, 0)
        x = root[0]
        while x.internal:
            i = slot(x, k)[0] + 1
            if len(x.child[i].key) == 2 * T - 1:
                split(x, i)
                if k > pos[x.key[i]]: i += 1
            x = x.child[i]
        x.key.insert(slot(x, k)[0] + 1, ident)

    def in_order(x, out):
        for i, kk in enumerate(x.key):
            if x.internal: in_order(x.child[i], out)
            out.append(kk)
        if x.internal: in_order(x.child[-1], out)

    want_lower = []
    for ident, k in enumerate(keys.tolist()):
        want_lower.append(lower(k) if ident else -1)
        pos[ident] = k; put(ident)
    order = []; in_order(root[0], order)
    got_lower = np.zeros(n, np.int32); got_order = np.zeros(n, np.int32)
    lib.bm2e_chain_tree.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bm2e_chain_tree(keys.ctypes.data, n, got_lower.ctypes.data, got_order.ctypes.data)
    assert got_order.tolist() == order
    assert got_lower.tolist() == want_lower
