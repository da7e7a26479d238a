"""ctypes binding of tests/host_emul/libbm2emul.so: the kernels' per-thread logic compiled for the host
(test-only; see tests/host_emul/emul.cpp)."""
from __future__ import annotations
import ctypes as C, os, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        import oracle_lib
        oracle_lib.lib()
        d = os.path.join(ROOT, "tests", "host_emul")
        so = os.path.join(d, "libbm2emul.so")
        srcs = [os.path.join(d, "emul.cpp")] + [os.path.join(ROOT, "bwa-mem2_b200", "csrc", f) for f in
                                                ("fm_device.cuh", "chain_device.cuh", "ext_device.cuh", "hd.h")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-ffp-contract=off",
                                   "-I" + os.path.join(ROOT, "bwa-mem2_b200", "csrc"), "-I" + os.path.join(ROOT, "include"),
                                   os.path.join(d, "emul.cpp"), "-o", so, "-L" + os.path.join(ROOT, "oracle"), "-lbm2oracle",
                                   "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
        _LIB = C.CDLL(so)
    return _LIB


def _capi():
    from __graft_entry__ import load_package
    return load_package().capi


def _arr(p, n, dt):
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(n, 1) * dt.itemsize,))[:n * dt.itemsize].view(dt).copy()


def _free(*ps):
    import oracle_lib
    for p in ps:
        oracle_lib.lib().bm2o_free(p)


def _batch(codes, offsets):
    capi = _capi()
    codes = np.ascontiguousarray(codes, np.uint8); offsets = np.ascontiguousarray(offsets, np.int64)
    return capi.ReadBatch(len(offsets) - 1, codes.ctypes.data, offsets.ctypes.data), (codes, offsets)


def collect_smems(index, opt, codes, offsets):
    capi = _capi(); rb, keep = _batch(codes, offsets)
    out = C.c_void_p(); n_ext = C.c_int64()
    L = lib(); L.emul_collect_smems.restype = C.c_int64
    n = L.emul_collect_smems(C.byref(index.desc), C.byref(opt), C.byref(rb), C.byref(out), C.byref(n_ext))
    a = _arr(out, n, capi.SMEM_DT); _free(out)
    return a, n_ext.value


def seed_chain(index, opt, codes, offsets):
    capi = _capi(); rb, keep = _batch(codes, offsets)
    ch = C.c_void_p(); sd = C.c_void_p(); off = C.c_void_p(); nc = C.c_int64(); ns = C.c_int64()
    lib().emul_seed_chain(C.byref(index.desc), C.byref(opt), C.byref(rb), C.byref(ch), C.byref(nc), C.byref(sd), C.byref(ns), C.byref(off))
    chains = _arr(ch, nc.value, capi.CHAIN_DT); seeds = _arr(sd, ns.value, capi.SEED_DT)
    offs = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int64)), shape=(rb.n_reads + 1,)).copy()
    _free(ch, sd, off)
    return chains, seeds, offs


def seed_chain_extend(index, opt, codes, offsets):
    capi = _capi(); rb, keep = _batch(codes, offsets)
    regs = C.c_void_p(); off = C.c_void_p(); n = C.c_int64()
    lib().emul_seed_chain_extend(C.byref(index.desc), C.byref(opt), C.byref(rb), C.byref(regs), C.byref(n), C.byref(off))
    a = _arr(regs, n.value, capi.REG_DT)
    offs = np.ctypeslib.as_array(C.cast(off, C.POINTER(C.c_int64)), shape=(rb.n_reads + 1,)).copy()
    _free(regs, off)
    return a, offs
