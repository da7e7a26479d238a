"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/bm2_b200.h declares,
its structs match the header's layout, and compute entries fail loudly without a CUDA device."""
import ctypes as C, os, re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    hdr = open(os.path.join(ROOT, "include", "bm2_b200.h")).read()
    declared = set(re.findall(r"\b(bm2_[a-z0-9_]+)\s*\(", hdr))
    lib = pkg.capi.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert set(pkg.capi.EXPORTS) <= declared
    assert lib.bm2_abi_version() == 1


def test_struct_layouts(pkg):
    capi = pkg.capi
    assert capi.PAIR_DT.itemsize == 56            # SeqPair, reference src/bandedSWA.h:90-99
    assert capi.SMEM_DT.itemsize == 40            # SMEM, reference src/FMI_search.h:75-83
    assert capi.REG_DT.itemsize == 112            # mem_alnreg_t, reference src/bwamem.h:137-160
    assert C.sizeof(capi.MemOpt) == 176           # mem_opt_t, reference src/bwamem.h:76-108


def test_default_options_match_mem_opt_init(pkg):
    o = pkg.capi.default_opt()
    assert (o.a, o.b, o.o_del, o.e_del, o.o_ins, o.e_ins, o.w, o.zdrop) == (1, 4, 6, 1, 6, 1, 100, 100)
    assert (o.min_seed_len, o.max_occ, o.max_chain_gap, o.split_width, o.max_mem_intv) == (19, 500, 10000, 10, 20)
    assert (o.pen_clip5, o.pen_clip3, o.mapQ_coef_fac) == (5, 5, 3)
    mat = np.array(list(o.mat)).reshape(5, 5)
    assert mat[0, 0] == 1 and mat[0, 1] == -4 and (mat[4] == -1).all() and (mat[:, 4] == -1).all()


def test_native_index_loader_reads_reference_index(pkg, golden_dir):
    idx = pkg.capi.Index(os.path.join(golden_dir, "c0_index", "ref.fa"))
    d = idx.desc
    assert d.reference_seq_len == 2 * d.l_pac + 1
    assert d.n_seqs == 4 and d.count[0] == 1 and d.count[4] == d.reference_seq_len
    ref = np.ctypeslib.as_array(C.cast(d.ref_string, C.POINTER(C.c_uint8)), shape=(2 * d.l_pac,))
    assert ref.max() <= 3
    fwd, rev = ref[:d.l_pac], ref[d.l_pac:]
    assert np.array_equal(rev, 3 - fwd[::-1])     # second half is the reverse complement
    idx.close()


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.capi.Bm2Error) as e:
        pkg.capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_sam_stage_switches_validate_their_arguments(pkg):
    """bm2_set_sam_staged / bm2_last_sam_stats: host-only entries; without a context they return an error code, nothing runs."""
    lib = pkg.capi.lib()
    lib.bm2_set_sam_staged.argtypes = [C.c_void_p, C.c_int]
    lib.bm2_last_sam_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    assert lib.bm2_set_sam_staged(None, 1) == 1
    ms = (C.c_double * 4)(); cnt = (C.c_ulonglong * 6)()
    assert lib.bm2_last_sam_stats(None, ms, cnt, 4, 6) == 1
    assert {"bm2_set_sam_staged", "bm2_last_sam_stats"} <= set(pkg.capi.EXPORTS)


def test_sibling_context_needs_a_parent(pkg):
    """bm2_create_sibling (a second context on the parent's index, one per host worker of bm2_mem): NULL arguments are an error code, not a crash."""
    lib = pkg.capi.lib()
    lib.bm2_create_sibling.argtypes = [C.POINTER(C.c_void_p), C.c_void_p]
    out = C.c_void_p(123)
    assert lib.bm2_create_sibling(C.byref(out), None) == 1
    assert lib.bm2_create_sibling(None, None) == 1
    assert "bm2_create_sibling" in pkg.capi.EXPORTS
