// bsw_pair_emul.cpp — TEST-ONLY host build of the two-jobs-per-thread BSW DP (bwa-mem2_b200/csrc/bsw_pair.cuh, the
// code bsw_pair_kernel runs per thread) so that it can be checked against the oracle on a machine without a GPU.
// The packed-halfword instructions are replaced by the portable definitions in the header.
#include <vector>
#include <cstdint>
#include <cstring>
static std::vector<int> *g_trace = nullptr;       // optional: the three column segments of every row (for lane-utilisation studies)
static inline void pair_trace(int a, int b, int c, int d, int e, int f) {
    if (g_trace) { int v[6] = {a, b, c, d, e, f}; g_trace->insert(g_trace->end(), v, v + 6); }
}
#define BM2_PAIR_TRACE pair_trace
#include "bsw_pair.cuh"

struct HostPairMem {
    uint32_t *state; uint16_t *selv;
    uint32_t ld(int j) const { return state[j]; }
    void st(int j, uint32_t w) const { state[j] = w; }
    uint32_t ld_half(int j, int l) const { return (state[j] >> (16 * l)) & 0xFFFFu; }
    void st_half(int j, int l, uint32_t v) const { state[j] = (state[j] & ~(0xFFFFu << (16 * l))) | ((v & 0xFFFFu) << (16 * l)); }
    uint32_t sel(int j) const { return selv[j]; }
};

// jobs 2k and 2k+1 of the given order share a thread; out = 6 ints per job; returns the number of DP cells.
// eligible[i] == 0 jobs must not be passed.  Sequences: query[qoff + k], target[toff + k] (stride +1).
// per thread (pair) k: trace_off[k]..trace_off[k+1] rows of 6 ints in trace (call with out_trace = null to skip)
extern "C" long long pair_extend_trace(int n, const int64_t *qoff, const int64_t *toff, const int32_t *qlen, const int32_t *tlen,
                                       const int32_t *h0, const uint8_t *qbuf, const uint8_t *tbuf, const int32_t *prm, int32_t *out,
                                       int32_t *trace, long long trace_cap, int64_t *trace_off);

extern "C" long long pair_extend_all(int n, const int64_t *qoff, const int64_t *toff, const int32_t *qlen, const int32_t *tlen,
                                     const int32_t *h0, const uint8_t *qbuf, const uint8_t *tbuf, const int32_t *prm /*9*/, int32_t *out)
{
    return pair_extend_trace(n, qoff, toff, qlen, tlen, h0, qbuf, tbuf, prm, out, nullptr, 0, nullptr);
}

extern "C" long long pair_extend_trace(int n, const int64_t *qoff, const int64_t *toff, const int32_t *qlen, const int32_t *tlen,
                                       const int32_t *h0, const uint8_t *qbuf, const uint8_t *tbuf, const int32_t *prm, int32_t *out,
                                       int32_t *trace, long long trace_cap, int64_t *trace_off)
{
    std::vector<int> tr;
    g_trace = trace ? &tr : nullptr;
    BswParams p; p.a = prm[0]; p.b = prm[1]; p.o_del = prm[2]; p.e_del = prm[3]; p.o_ins = prm[4]; p.e_ins = prm[5];
    p.zdrop = prm[6]; p.end_bonus = prm[7]; p.w = prm[8];
    if (!p2_params_ok(p)) return -1;
    unsigned long long cells = 0;
    std::vector<uint32_t> state(258);
    std::vector<uint16_t> sel(258);
    for (int k = 0; k < n; k += 2) {
        const int nj = k + 1 < n ? 2 : 1;
        int ql[2] = {qlen[k], nj == 2 ? qlen[k + 1] : 0}, tl[2] = {tlen[k], nj == 2 ? tlen[k + 1] : 0}, hh[2] = {h0[k], nj == 2 ? h0[k + 1] : 0};
        // stale garbage on purpose: the kernel's shared memory is not cleared between jobs either
        for (auto &x : state) x = 0xA5A5A5A5u * (uint32_t) (k + 1);
        const int qmax = ql[0] > ql[1] ? ql[0] : ql[1];
        for (int j = 0; j <= qmax; ++j) {
            const int qa = j < ql[0] ? qbuf[qoff[k] + j] : 0, qb = (nj == 2 && j < ql[1]) ? qbuf[qoff[k + 1] + j] : 0;
            sel[j] = (uint16_t) p2_selector(qa, qb);
        }
        HostPairMem mem{state.data(), sel.data()};
        BswOut o[2];
        if (trace_off) trace_off[k / 2] = (int64_t) tr.size() / 6;
        bsw_pair_extend(mem, tbuf + toff[k], 1, tbuf + (nj == 2 ? toff[k + 1] : toff[k]), 1, ql, tl, hh, nj, p, o, cells);
        for (int l = 0; l < nj; ++l) {
            int32_t *d = out + 6 * (size_t) (k + l);
            d[0] = o[l].score; d[1] = o[l].tle; d[2] = o[l].gtle; d[3] = o[l].qle; d[4] = o[l].gscore; d[5] = o[l].max_off;
        }
    }
    if (trace_off) trace_off[(n + 1) / 2] = (int64_t) tr.size() / 6;
    if (trace) { if ((long long) tr.size() > trace_cap) return -2; memcpy(trace, tr.data(), tr.size() * sizeof(int)); }
    g_trace = nullptr;
    return (long long) cells;
}
