// next_rows_check.cu — COMPILE CHECK ONLY (the object is not linked into libbm2b200.so): the device logic of the next rows
// (ksw_device.cuh, mate_device.cuh, sam_device.cuh) is instantiated in a kernel so that nvcc / ptxas for
// sm_100a see it every build.  Their semantics are checked on the host (tests/host_emul/{ksw,mate,sam}_emul.cpp); kernels and the C ABI
// around them are the next round's work.
#include "bm2_common.cuh"
#include "sam_device.cuh"

struct NextRowsEmit {
    SamRec *recs; int *n;
    __device__ void operator()(int, int, const SamRec &r, const uint32_t *, const char *) { recs[(*n)++] = r; }
};

__global__ void next_rows_check_kernel(SamParams p, SamTables tb, ContigView cv, MatePes pes, const uint8_t *ref, const uint8_t *s0, const uint8_t *s1, int l0, int l1,
                                       bm2_alnreg_t *a0, bm2_alnreg_t *a1, bm2_alnreg_t *b0, bm2_alnreg_t *b1, int n0, int n1, MateScratch ms, SamScratch sc,
                                       SamRec *recs, int *n_recs, int *overflow)
{
    const uint8_t *seq[2] = { s0, s1 }; const int l_seq[2] = { l0, l1 };
    bm2_alnreg_t *a[2] = { a0, a1 }, *b[2] = { b0, b1 };
    int n[2] = { n0, n1 };
    mate_rescue_pair_d(cv, p.ep, p.min_seed_len, p.pen_unpaired, 50, pes, ref, seq, l_seq, a, n, b, ms, overflow);
    NextRowsEmit emit = { recs, n_recs };
    sam_pe_pair_d(p, tb, cv, pes, ref, seq, l_seq, a, n, (int) blockIdx.x, sc, emit, overflow);
}
