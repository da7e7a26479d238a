// bsw_rowscan.cpp — TEST-ONLY lane-by-lane simulation of the warp-per-job BSW kernel (bsw_warp_kernel in
// bwa-mem2_b200/csrc/bsw.cu): the row of the banded extension DP is split over 32 lanes; F (the horizontal gap
// state) is obtained with an exclusive max-scan because F(j) depends only on M(k), k < j:
//     F(j) = max(0, max_{beg<=k<j} (max(M(k) - oe_ins, 0) - (j-1-k) * e_ins)).
// The state lives in a circular buffer of Wcap >= 2w+4 columns; never-visited columns are initialised on demand
// with the first-row values.  This file exists to check that formulation against the oracle on the CPU.
#include <vector>
#include <cstdint>
#include <cstdlib>
#include <algorithm>

extern "C" void rowscan_extend(const uint8_t *query, int qlen, const uint8_t *target, int tlen, int w, int h0,
                               int a, int b, int o_del, int e_del, int o_ins, int e_ins, int zdrop, int end_bonus, int *out)
{
    const int LANES = 32;
    const int NEG = -(1 << 29);
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    {   // SIMD-wrapper band arithmetic (as the product kernels)
        unsigned t1 = ((unsigned) (qlen * a) + (unsigned) (end_bonus - o_ins)) & 0xFFFFu;
        int max_ins = (int) (t1 / (unsigned) e_ins) + 1; if (max_ins < 1) max_ins = 1;
        unsigned t2 = ((unsigned) (qlen * a) + (unsigned) (end_bonus - o_del)) & 0xFFFFu;
        int max_del = (int) (t2 / (unsigned) e_del) + 1; if (max_del < 1) max_del = 1;
        if (w > max_ins) w = max_ins;
        if (w > max_del) w = max_del;
    }
    const int Wcap = 2 * w + 8;
    std::vector<int> H(Wcap, 0), E(Wcap, 0);
    auto init_h = [&](int j) {   // first row: H(-1, j-1)
        if (j == 0) return h0;
        int v = h0 - oe_ins - (j - 1) * e_ins;
        // reference: eh[1] = h0 > oe_ins ? h0 - oe_ins : 0, then decreasing by e_ins while > e_ins... => max(.,0) with the
        // same zero crossing: eh[j] = eh[j-1] - e_ins while eh[j-1] > e_ins, else 0
        return v > 0 ? v : 0;
    };
    int max_init = -1;
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    for (int i = 0; i < tlen; ++i) {
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        for (int j = max_init + 1; j <= end; ++j) { H[j % Wcap] = init_h(j); E[j % Wcap] = 0; }
        if (end > max_init) max_init = end;
        const int h1_init = beg == 0 ? std::max(h0 - (o_del + e_del * (i + 1)), 0) : 0;
        const int n = end - beg;
        const int c = (n + LANES - 1) / LANES;
        const int tb = target[i];
        // phase 1 per lane: M and the local scan values
        std::vector<int> M(std::max(n, 0)), Ecur(std::max(n, 0)), pl(std::max(n, 0)), agg(LANES, NEG), hs(std::max(n, 0));
        for (int L = 0; L < LANES; ++L) {
            int run = NEG;
            for (int k = 0; k < c; ++k) {
                int j = beg + L * c + k; if (j >= end) break;
                int hd = H[j % Wcap]; Ecur[j - beg] = E[j % Wcap];
                int qb = query[j];
                int s = (qb > 3 || tb > 3) ? -1 : (qb == tb ? a : -b);
                int m = hd ? hd + s : 0;
                M[j - beg] = m;
                pl[j - beg] = run;
                int t = std::max(m - oe_ins, 0);
                run = std::max(run, t + j * e_ins);
            }
            agg[L] = run;
        }
        // phase 2: exclusive max-scan over lanes
        std::vector<int> lp(LANES, NEG);
        { int run = NEG; for (int L = 0; L < LANES; ++L) { lp[L] = run; run = std::max(run, agg[L]); } }
        // phase 3
        int m_row = 0, mj = -1;
        for (int L = 0; L < LANES; ++L) {
            for (int k = 0; k < c; ++k) {
                int j = beg + L * c + k; if (j >= end) break;
                int P = std::max(lp[L], pl[j - beg]);
                int f = std::max(P - (j - 1) * e_ins, 0);
                if (P == NEG) f = 0;
                int m = M[j - beg], e = Ecur[j - beg];
                int h = std::max(std::max(m, e), f);
                hs[j - beg] = h;
                int t = std::max(m - oe_del, 0);
                E[j % Wcap] = std::max(e - e_del, t);
                if (h >= m_row) { mj = j; }
                if (h > m_row) m_row = h;
            }
        }
        // H(i, j-1) for the next row
        for (int j = beg; j < end; ++j) H[j % Wcap] = j == beg ? h1_init : hs[j - 1 - beg];
        const int h1 = n > 0 ? hs[n - 1] : h1_init;
        H[end % Wcap] = h1; E[end % Wcap] = 0;
        if (end == qlen) {
            if (h1 >= gscore) best_ie = i;
            if (h1 > gscore) gscore = h1;
        }
        if (m_row == 0) break;
        if (m_row > best) {
            best = m_row; best_i = i; best_j = mj;
            int d = std::abs(mj - i); if (d > max_off) max_off = d;
        } else if (zdrop > 0) {
            int di = i - best_i, dj = mj - best_j;
            int pen = di > dj ? di - dj : dj - di;
            if (best - m_row - pen > zdrop) break;
        }
        int j;
        for (j = beg; j < end && H[j % Wcap] == 0 && E[j % Wcap] == 0; ++j) {}
        beg = j;
        for (j = end; j >= beg && H[j % Wcap] == 0 && E[j % Wcap] == 0; --j) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    out[0] = best; out[1] = best_j + 1; out[2] = best_i + 1; out[3] = best_ie + 1; out[4] = gscore; out[5] = max_off;
}
