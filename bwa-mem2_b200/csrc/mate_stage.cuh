// mate_stage.cuh — the staged form of mate rescue: job records, their enumeration per pair, and the table lookup of the per-pair logic.
//
// The rescue block of mem_sam_pe (reference src/bwamem_pair.cpp:378-412) calls mem_matesw (:150-283) anchor by anchor; every call is a
// handful of local alignments (ksw_align2, src/ksw.cpp:324-381) - the bulk of the SAM stage's arithmetic.  The reference batches them
// across pairs for its SIMD kernel (mem_sam_pe_batch*, src/bwamem_pair.cpp:930-1248, src/kswv.cpp); here they become jobs of a
// warp-per-window kernel (sam.cu: sam_jobs_kernel -> sam_ksw_jobs_kernel -> sam_kernel with MateKswTable).
// Everything in this header is plain BM2_HD logic: tests/host_emul/sam_emul.cpp runs the same functions on the host
// (tests/test_oracle_sam_pe.py::test_staged_rescue_equals_the_per_pair_block).
#pragma once
#include "mate_device.cuh"

struct MateJob { int64_t rb, re; int32_t pair; int16_t j; int8_t ai, r_rev; };      // r_rev = orientation | is_rev << 2; pair = index in the batch
struct PairJobs { int32_t begin, count; };                                        // a pair's slice of the job table
struct MateStats { unsigned int n_jobs, looked_up, in_place, window_moved; };
struct MateJobRes { int32_t score, te, qe, score2, te2, tb, qb, valid; };          // = bm2_ksw_res with _pad as the "computed" mark

#if defined(__CUDA_ARCH__)
#define BM2_STAT_INC(p) atomicAdd((p), 1u)
#else
#define BM2_STAT_INC(p) (++*(p))
#endif

// upper bound of the jobs of one pair: every anchor asks for at most 4 windows, at most max_matesw anchors per read
BM2_HD long long mate_jobs_bound_d(long long n0, long long n1, int max_matesw) {
    const long long mm = max_matesw > 0 ? max_matesw : 0;
    return 4 * ((n0 < mm ? n0 : mm) + (n1 < mm ? n1 : mm));
}

// the jobs of one pair, counted (out == nullptr) or written; the anchor index j is the index among the anchors that pass the score test,
// as matesw_d's caller numbers them
BM2_HD int mate_jobs_list_d(const ContigView &cv, int min_seed_len, int pen_unpaired, int max_matesw, const MatePes &pes, const int l_seq[2],
                            const bm2_alnreg_t *const a[2], const int n[2], int pair, MateJob *out)
{
    int k = 0;
    auto emit = [&](int ai, int j, int r, int64_t rb, int64_t re, int is_rev) {
        if (out) { MateJob o; o.rb = rb; o.re = re; o.pair = pair; o.j = (int16_t) j; o.ai = (int8_t) ai; o.r_rev = (int8_t) (r | (is_rev << 2)); out[k] = o; }
        ++k;
    };
    mate_jobs_pair_d(cv, min_seed_len, pen_unpaired, max_matesw, pes, l_seq, a, n, emit);
    return k;
}

// how a job reads its query: the mate in place, from its last base and complemented when the window lies on the other strand;
// xtra as mem_matesw sets it (src/bwamem_pair.cpp:186-189)
struct MateJobQuery { const uint8_t *q; int stride, comp, l_ms, tlen, xtra; };
BM2_HD MateJobQuery mate_job_query_d(const MateJob &jb, const uint8_t *codes, const int64_t *offs, int a_match, int min_seed_len) {
    const int64_t mr = 2LL * jb.pair + !jb.ai;                                  // the mate of the anchor's read
    const uint8_t *ms = codes + offs[mr];
    MateJobQuery o;
    o.l_ms = (int) (offs[mr + 1] - offs[mr]); o.tlen = (int) (jb.re - jb.rb);
    const int is_rev = (jb.r_rev >> 2) & 1;
    o.q = is_rev ? ms + o.l_ms - 1 : ms; o.stride = is_rev ? -1 : 1; o.comp = is_rev;
    o.xtra = BM2_KSW_XSUBO | BM2_KSW_XSTART | (o.l_ms * a_match < 250 ? BM2_KSW_XBYTE : 0) | (min_seed_len * a_match);
    return o;
}

// one window by one thread (staged mode 2): the contiguous query ksw_align2_d wants (the reverse complement built in rev) and the one-thread sweep
BM2_HD KswRes mate_job_align_thread_d(const MateJobQuery &q, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                                      int32_t *ksw, int32_t *bsc, int32_t *bpos, int bcap, uint8_t *tmp, uint8_t *rev, int *overflow)
{
    const uint8_t *seq = q.q;
    if (q.comp) { for (int i = 0; i < q.l_ms; ++i) { const uint8_t b = q.q[-i]; rev[i] = b < 4 ? 3 - b : 4; } seq = rev; }      // q.q = the mate's last base
    return ksw_align2_d(q.l_ms, seq, q.tlen, target, mat, o_del, e_del, o_ins, e_ins, q.xtra, ksw, bsc, bpos, bcap, tmp, overflow);
}

// The provider of the local alignment for mate_rescue_pair_d in the staged form: the pair's slice of the job table, else the computation
// in place (a window that moved because an earlier rescue of the pair changed the regions, a job that was not listed or not computed).
struct MateKswTable {
    const MateJob *jobs; const MateJobRes *res; int n; MateKswDirect direct; MateStats *stats;
    BM2_HD KswRes operator()(int ai, int j, int r, int l_ms, const uint8_t *seq, int64_t rb, int64_t re, int xtra) const {
        for (int k = 0; k < n; ++k) {
            const MateJob &jb = jobs[k];
            if (jb.ai != ai || jb.j != j || (jb.r_rev & 3) != r) continue;
            if (jb.rb == rb && jb.re == re && res[k].valid) {
                const MateJobRes &t = res[k];
                KswRes o; o.score = t.score; o.te = t.te; o.qe = t.qe; o.score2 = t.score2; o.te2 = t.te2; o.tb = t.tb; o.qb = t.qb;
                BM2_STAT_INC(&stats->looked_up);
                return o;
            }
            if (jb.rb != rb || jb.re != re) BM2_STAT_INC(&stats->window_moved);
            break;
        }
        BM2_STAT_INC(&stats->in_place);
        return direct(ai, j, r, l_ms, seq, rb, re, xtra);
    }
};
