// bsw_col2_emul.cpp — TEST-ONLY host build of the two-columns-per-instruction BSW DP (bwa-mem2_b200/csrc/bsw_col2.cuh,
// the code bsw_col2_kernel runs per thread) so that it can be checked against the oracle on a machine without a GPU.
// The packed-halfword instructions are replaced by the portable definitions in bsw_pair.cuh.
#include <vector>
#include <cstdint>
#include <cstring>
static std::vector<int> *g_trace = nullptr;       // optional: the band [beg, end) of every row (lane-utilisation studies)
static inline void col2_trace(int b, int e) { if (g_trace) { g_trace->push_back(b); g_trace->push_back(e); } }
#define BM2_COL2_TRACE col2_trace
#include "bsw_col2.cuh"

struct HostCol2Mem {
    uint16_t *state; uint8_t *selb;
    uint32_t ldw(int p) const { return (uint32_t) state[2 * p] | ((uint32_t) state[2 * p + 1] << 16); }
    void stw(int p, uint32_t w) const { state[2 * p] = (uint16_t) (w & 0xFFFFu); state[2 * p + 1] = (uint16_t) (w >> 16); }
    uint32_t ldh(int j) const { return state[j]; }
    void sth(int j, uint32_t v) const { state[j] = (uint16_t) v; }
    uint32_t sel16(int q) const { return (uint32_t) selb[2 * q] | ((uint32_t) selb[2 * q + 1] << 8); }
};

// out = 6 ints per job (score, tle, gtle, qle, gscore, max_off); returns the number of DP cells, -1 for unsupported scoring.
// Sequences: query[qoff + k * qstride], target[toff + k * tstride].
// rows[k] = number of rows job k ran; widths: per job and row the band width end - beg (concatenated, cap entries)
extern "C" long long col2_extend_trace(int n, const int64_t *qoff, const int64_t *toff, const int32_t *qlen, const int32_t *tlen,
                                       const int32_t *h0, const int32_t *qstride, const int32_t *tstride, const uint8_t *qbuf, const uint8_t *tbuf,
                                       const int32_t *prm, int32_t *out, int32_t *rows, int16_t *widths, long long cap);

extern "C" long long col2_extend_all(int n, const int64_t *qoff, const int64_t *toff, const int32_t *qlen, const int32_t *tlen,
                                     const int32_t *h0, const int32_t *qstride, const int32_t *tstride, const uint8_t *qbuf, const uint8_t *tbuf,
                                     const int32_t *prm /*9*/, int32_t *out)
{
    return col2_extend_trace(n, qoff, toff, qlen, tlen, h0, qstride, tstride, qbuf, tbuf, prm, out, nullptr, nullptr, 0);
}

extern "C" long long col2_extend_trace(int n, const int64_t *qoff, const int64_t *toff, const int32_t *qlen, const int32_t *tlen,
                                       const int32_t *h0, const int32_t *qstride, const int32_t *tstride, const uint8_t *qbuf, const uint8_t *tbuf,
                                       const int32_t *prm, int32_t *out, int32_t *rows, int16_t *widths, long long cap)
{
    BswParams p; p.a = prm[0]; p.b = prm[1]; p.o_del = prm[2]; p.e_del = prm[3]; p.o_ins = prm[4]; p.e_ins = prm[5];
    p.zdrop = prm[6]; p.end_bonus = prm[7]; p.w = prm[8];
    if (!c2_params_ok(p)) return -1;
    std::vector<int> tr;
    long long wpos = 0;
    unsigned long long cells = 0;
    std::vector<uint16_t> state(264);
    std::vector<uint8_t> sel(268);
    for (int k = 0; k < n; ++k) {
        if (qlen[k] > 256) return -2;
        // stale garbage on purpose: the kernel's shared memory is not cleared between jobs either
        for (size_t x = 0; x < state.size(); ++x) state[x] = (uint16_t) (0xA5A5u * (unsigned) (k + 1) + x * 7u);
        for (size_t x = 0; x < sel.size(); ++x) sel[x] = (uint8_t) c2_selector_byte(4);
        for (int j = 0; j < qlen[k]; ++j) sel[j] = (uint8_t) c2_selector_byte(qbuf[qoff[k] + (int64_t) j * qstride[k]]);
        HostCol2Mem mem{state.data(), sel.data()};
        BswOut o;
        tr.clear(); g_trace = rows ? &tr : nullptr;
        // every instance the kernel compiles, by job index: the three band-shrink forms under equal penalties (0 = scans, the product's
        // default; 1, 2 = edge columns from registers first) and the general form (also under equal penalties)
        const bool same = p.o_del + p.e_del == p.o_ins + p.e_ins;
        if (same && (k & 3) == 0) bsw_col2_extend<true, HostCol2Mem, 0>(mem, tbuf + toff[k], tstride[k], qlen[k], tlen[k], h0[k], p, o, cells);
        else if (same && (k & 3) == 1) bsw_col2_extend<true, HostCol2Mem, 1>(mem, tbuf + toff[k], tstride[k], qlen[k], tlen[k], h0[k], p, o, cells);
        else if (same && (k & 3) == 2) bsw_col2_extend<true, HostCol2Mem, 2>(mem, tbuf + toff[k], tstride[k], qlen[k], tlen[k], h0[k], p, o, cells);
        else bsw_col2_extend<false>(mem, tbuf + toff[k], tstride[k], qlen[k], tlen[k], h0[k], p, o, cells);
        g_trace = nullptr;
        if (rows) {
            rows[k] = (int32_t) (tr.size() / 2);
            for (size_t r = 0; r < tr.size() / 2; ++r) { if (wpos >= cap) return -3; widths[wpos++] = (int16_t) (tr[2 * r + 1] > tr[2 * r] ? tr[2 * r + 1] - tr[2 * r] : 0); }
        }
        int32_t *d = out + 6 * (size_t) k;
        d[0] = o.score; d[1] = o.tle; d[2] = o.gtle; d[3] = o.qle; d[4] = o.gscore; d[5] = o.max_off;
    }
    return (long long) cells;
}
