// bsw_pair.cuh — two extension jobs per thread in packed 16-bit halves (DPX: VIADDMNMX.S16x2, VIMNMX3.S16x2).
//
// Same DP as bsw_extend_one (bsw.cu; reference src/bandedSWA.cpp:116-237 with the SIMD-wrapper band and z-drop,
// :2905-2926) for jobs whose scores fit 8 bits (h0 + min(qlen, tlen) * a <= 255: every job of a 2x151 bp read),
// whose query has no N and is at most 255 columns long.  Job A lives in the low half of every packed register,
// job B in the high half; each keeps its OWN band [beg, end), exit row and outputs - a row is run as up to three
// column segments (only the job that starts first / both / only the job that ends last) through one cell loop.  Per pair of cells: 1 LDS + 1 STS of the packed state {H_A, E_A, H_B, E_B} (4 x 8 bit), one PRMT for both
// substitution scores (the per-column selector is precomputed from the two queries), 3 VIADDMNMX.S16x2, 1 VIMNMX3,
// and the row maximum + its last column as one unsigned key (h << 8 | j) per half.
//
// Written as BM2_HD so that tests/host_emul/bsw_pair_emul.cpp runs the very same code on the CPU against the oracle.
#pragma once
#include "hd.h"
#include "bsw_types.h"

#if defined(__CUDA_ARCH__)
BM2_D uint32_t p2_prmt(uint32_t a, uint32_t b, uint32_t s) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d; }
BM2_D uint32_t p2_addmin_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2_relu(a, b, c); }   // max(min(a + b, c), 0)
BM2_D uint32_t p2_addmax_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }   // max(a + b, c, 0)
BM2_D uint32_t p2_max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
BM2_D uint32_t p2_maxu(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
BM2_D uint32_t p2_addmaxu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_u16x2(a, b, c); }               // max(a + b mod 2^16, c), unsigned halves
BM2_D uint32_t p2_add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
// a * b + c as ONE multiply-add on the FMA pipe (the compiler would turn constant multiplies into ALU-pipe shift/mask pairs)
BM2_D uint32_t p2_mad(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
#else
inline uint32_t p2_mad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
inline uint32_t p2_prmt(uint32_t a, uint32_t b, uint32_t s) {            // PTX prmt.b32, default mode
    const uint64_t src = ((uint64_t) b << 32) | a;
    uint32_t d = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (s >> (4 * i)) & 0xF;
        uint32_t byte = (uint32_t) (src >> (8 * (n & 7))) & 0xFF;
        if (n & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        d |= byte << (8 * i);
    }
    return d;
}
inline int16_t p2_lo(uint32_t x) { return (int16_t) (x & 0xFFFF); }
inline int16_t p2_hi(uint32_t x) { return (int16_t) (x >> 16); }
inline uint32_t p2_mk(int lo, int hi) { return ((uint32_t) lo & 0xFFFF) | (((uint32_t) hi & 0xFFFF) << 16); }
inline int p2_i16(int v) { return (int16_t) v; }
inline uint32_t p2_addmin_relu(uint32_t a, uint32_t b, uint32_t c) {
    auto f = [](int x, int y, int z) { int t = p2_i16(x + y); t = t < z ? t : z; return t > 0 ? t : 0; };
    return p2_mk(f(p2_lo(a), p2_lo(b), p2_lo(c)), f(p2_hi(a), p2_hi(b), p2_hi(c)));
}
inline uint32_t p2_addmax_relu(uint32_t a, uint32_t b, uint32_t c) {
    auto f = [](int x, int y, int z) { int t = p2_i16(x + y); t = t > z ? t : z; return t > 0 ? t : 0; };
    return p2_mk(f(p2_lo(a), p2_lo(b), p2_lo(c)), f(p2_hi(a), p2_hi(b), p2_hi(c)));
}
inline uint32_t p2_max3(uint32_t a, uint32_t b, uint32_t c) {
    auto f = [](int x, int y, int z) { int t = x > y ? x : y; return t > z ? t : z; };
    return p2_mk(f(p2_lo(a), p2_lo(b), p2_lo(c)), f(p2_hi(a), p2_hi(b), p2_hi(c)));
}
inline uint32_t p2_maxu(uint32_t a, uint32_t b) {
    const uint32_t lo = (a & 0xFFFF) > (b & 0xFFFF) ? (a & 0xFFFF) : (b & 0xFFFF), hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
inline uint32_t p2_addmaxu(uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t lo = (a + b) & 0xFFFFu, hi = ((a >> 16) + (b >> 16)) & 0xFFFFu, cl = c & 0xFFFFu, ch = c >> 16;
    return (lo > cl ? lo : cl) | ((hi > ch ? hi : ch) << 16);
}
inline uint32_t p2_add(uint32_t a, uint32_t b) { return ((a + b) & 0xFFFF) | ((((a >> 16) + (b >> 16)) & 0xFFFF) << 16); }
#endif

// Scores of one target base against query bases 0..3 as four signed bytes (tb > 3: all -1).
BM2_HD uint32_t p2_score_table(int tb, int a, int b) {
    const uint32_t sb8 = (uint32_t) (-b) & 0xFFu, sa8 = (uint32_t) a & 0xFFu;
    const uint32_t t = sb8 * 0x01010101u;
    return tb > 3 ? 0xFFFFFFFFu : ((t & ~(0xFFu << (8 * tb))) | (sa8 << (8 * tb)));
}
// PRMT selector of one column: result = {score_A (16 bit, sign-extended), score_B}; table A is operand 1, table B operand 2.
BM2_HD uint32_t p2_selector(int qa, int qb) { return (uint32_t) qa * 0x11u + (uint32_t) qb * 0x1100u + 0xC480u; }

// can a job go to the pair kernel?  (query N bases are checked by the caller)
BM2_HD bool p2_params_ok(const BswParams &p) {
    return p.a > 0 && p.a <= 127 && p.b >= 0 && p.b <= 127 && p.o_del >= 0 && p.e_del > 0 && p.o_ins >= 0 && p.e_ins > 0 &&
           p.o_del + p.e_del < 16384 && p.o_ins + p.e_ins < 16384;
}

struct PairConsts { uint32_t n_oe_del, n_e_del, n_oe_ins, n_e_ins; };     // negated penalties in both halves

// Which halves a column segment updates (act = 0xFFFF per active half).  ONE code path for "both", "A only" and
// "B only" (the lanes of a warp are in different situations; separate loops would serialise them): the idle half is
// forced to 0 on input (hmask / esel) so that no carry crosses the halves, and its state is written back unchanged
// (one LOP3 blend).
struct PairMode { uint32_t hmask, esel, act; };
BM2_HD PairMode p2_mode(bool a_on, bool b_on) {
    PairMode m;
    m.hmask = (a_on ? 0x000000FFu : 0u) | (b_on ? 0x00FF0000u : 0u);
    m.esel = (a_on ? 0x0001u : 0x0004u) | 0x0040u | (b_on ? 0x0300u : 0x0400u) | 0x4000u;     // E bytes 1 / 3 or the zero byte
    m.act = (a_on ? 0x0000FFFFu : 0u) | (b_on ? 0xFFFF0000u : 0u);
    return m;
}

// The two jobs of one thread.  n_jobs = 1 runs job A alone.  Mem: ld/st (packed state word of a column),
// ld_half/st_half (one job's {H, E << 8}), sel (column selector).
template <class Mem>
BM2_HD void bsw_pair_extend(const Mem &mem, const uint8_t *tptrA, int tstrideA, const uint8_t *tptrB, int tstrideB,
                            const int qlen_[2], const int tlen_[2], const int h0_[2], int n_jobs, const BswParams &p,
                            BswOut out[2], unsigned long long &cells)
{
    const int oe_ins = p.o_ins + p.e_ins, e_ins = p.e_ins, e_del = p.e_del;
    PairConsts c;
    c.n_oe_del = ((uint32_t) (-(p.o_del + p.e_del)) & 0xFFFFu) * 0x10001u;
    c.n_e_del = ((uint32_t) (-p.e_del) & 0xFFFFu) * 0x10001u;
    c.n_oe_ins = ((uint32_t) (-oe_ins) & 0xFFFFu) * 0x10001u;
    c.n_e_ins = ((uint32_t) (-p.e_ins) & 0xFFFFu) * 0x10001u;
    int qlen[2], tlen[2], h0[2], w[2], beg[2], end[2], best[2], best_i[2], best_j[2], best_ie[2], gscore[2], max_off[2], zthr[2];
    bool alive[2];
    unsigned long long ncell = 0;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        qlen[l] = qlen_[l]; tlen[l] = tlen_[l]; h0[l] = h0_[l];
        alive[l] = l < n_jobs && tlen[l] > 0;
        best[l] = h0[l]; best_i[l] = -1; best_j[l] = -1; best_ie[l] = -1; gscore[l] = -1; max_off[l] = 0;
        beg[l] = 0; end[l] = qlen[l];
        // band (SIMD wrapper arithmetic, bandedSWA.cpp:2905-2926)
        int ww = p.w;
        const BswQuirk qk = bsw_quirk(qlen[l], tlen[l], h0[l], p);
        zthr[l] = qk.zthr;
        unsigned t1 = ((unsigned) (qlen[l] * p.a) + (unsigned) (p.end_bonus - p.o_ins)) & qk.band_mask;
        int max_ins = (int) (t1 / (unsigned) e_ins) + 1; if (max_ins < 1) max_ins = 1;
        unsigned t2 = ((unsigned) (qlen[l] * p.a) + (unsigned) (p.end_bonus - p.o_del)) & qk.band_mask;
        int max_del = (int) (t2 / (unsigned) e_del) + 1; if (max_del < 1) max_del = 1;
        if (ww > max_ins) ww = max_ins;
        if (ww > max_del) ww = max_del;
        w[l] = ww;
        // first row (bandedSWA.cpp:141-144): columns 0..qlen, E = 0
        if (l < n_jobs) {
            int h = h0[l];
            mem.st_half(0, l, (uint32_t) h);
            h = h0[l] > oe_ins ? h0[l] - oe_ins : 0;
            for (int j = 1; j <= qlen[l]; ++j) {
                mem.st_half(j, l, (uint32_t) h);
                h = h > e_ins ? h - e_ins : 0;
            }
        }
    }
    for (int i = 0; alive[0] || alive[1]; ++i) {
        int b0 = 0, e0 = 0, b1 = 0, e1 = 0;
        uint32_t tbl0 = 0xFFFFFFFFu, tbl1 = 0xFFFFFFFFu, h1 = 0;
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (alive[l]) {
                if (beg[l] < i - w[l]) beg[l] = i - w[l];
                if (end[l] > i + w[l] + 1) end[l] = i + w[l] + 1;
                if (end[l] > qlen[l]) end[l] = qlen[l];
                int hi = 0;
                if (beg[l] == 0) { hi = h0[l] - (p.o_del + e_del * (i + 1)); if (hi < 0) hi = 0; }
                h1 |= (uint32_t) hi << (16 * l);
                const int tb = l == 0 ? tptrA[(long long) i * tstrideA] : tptrB[(long long) i * tstrideB];
                const uint32_t tv = p2_score_table(tb, p.a, p.b);
                const int bv = beg[l], ev = end[l] > beg[l] ? end[l] : beg[l];
                if (l == 0) { tbl0 = tv; b0 = bv; e0 = ev; } else { tbl1 = tv; b1 = bv; e1 = ev; }
            }
        }
        if (!alive[0]) b0 = e0 = e1;                   // an idle job contributes an empty interval at the other's end
        if (!alive[1]) b1 = e1 = e0;
        uint32_t f = 0, mkey = 0;
        {
            // three column segments: only the job that starts first / both / only the job that ends last
            const bool a_first = b0 <= b1, a_last = e0 >= e1;
            const int bmin = a_first ? b0 : b1, bmax = a_first ? b1 : b0, e_first = a_first ? e0 : e1;
            const int emin = a_last ? e1 : e0, emax = a_last ? e0 : e1;
            const int s1e = bmax < e_first ? bmax : e_first;
            const int s3b = emin > bmax ? emin : bmax;
#ifdef BM2_PAIR_TRACE
            BM2_PAIR_TRACE(bmin, s1e, bmax, emin, s3b, emax);
#endif
            // ONE cell loop per row: a lane that reaches the end of its segment switches to its next one inside the loop
            // (a short divergent block), so that the 32 lanes of a warp - each at its own segment - keep sharing the loop.
            // (Measured: a loop per segment runs the lanes apart, 14 of 32 active; simulated on real jobs the per-row
            // cost is max over lanes of the SUM of the segments, 0.90-0.98 of ideal, instead of the sum of the maxima, 0.62.)
            int sg = -1, j = 0, jend = 0;
            PairMode md; md.hmask = 0x00FF00FFu; md.esel = 0x4341u; md.act = 0xFFFFFFFFu;
            uint32_t sf = 0, sh = 0, sk = 0, jj = 0;
            auto next_segment = [&]() {       // only called while cells remain: a non-empty segment follows
                f = (f & md.act) | sf; h1 = (h1 & md.act) | sh; mkey = (mkey & md.act) | sk;      // the idle half gets its running values back
                int j0 = 0, j1 = 0;
                for (++sg; sg < 2; ++sg) {
                    j0 = sg == 0 ? bmin : bmax; j1 = sg == 0 ? s1e : emin;
                    if (j0 < j1) break;
                }
                if (sg == 2) { j0 = s3b; j1 = emax; }
                const bool a_on = sg == 1 || (sg == 0 ? a_first : a_last), b_on = sg == 1 || (sg == 0 ? !a_first : !a_last);
                md = p2_mode(a_on, b_on);
                sf = f & ~md.act; sh = h1 & ~md.act; sk = mkey & ~md.act;
                f &= md.act; h1 &= md.act; mkey &= md.act;
                j = j0; jend = j1; jj = (uint32_t) j0 * 0x10001u;
            };
            const int total = (s1e > bmin ? s1e - bmin : 0) + (emin > bmax ? emin - bmax : 0) + (emax > s3b ? emax - s3b : 0);
            if (total > 0) {
                next_segment();
                uint32_t wn = mem.ld(j), sn = mem.sel(j);
#pragma unroll 1
                for (int k = 0; k < total; ++k) {
                    const uint32_t w = wn, sl = sn;
                    wn = mem.ld(j + 1); sn = mem.sel(j + 1);          // next column's state and selector, ahead of this cell's arithmetic
                    const uint32_t hd = w & md.hmask;
                    const uint32_t e = p2_prmt(w, 0u, md.esel);
                    const uint32_t s = p2_prmt(tbl0, tbl1, sl);
                    const uint32_t M = p2_addmin_relu(hd, s, p2_mad(hd, 128u, 0u));          // hd ? max(hd + s, 0) : 0   (s <= 127)
                    const uint32_t h = p2_max3(M, e, f);
                    const uint32_t en = p2_addmax_relu(e, c.n_e_del, p2_add(M, c.n_oe_del));
                    const uint32_t wv = p2_mad(en, 256u, h1);
                    mem.st(j, (wv & md.act) | (w & ~md.act));
                    f = p2_addmax_relu(f, c.n_e_ins, p2_add(M, c.n_oe_ins));
                    h1 = h;
                    mkey = p2_maxu(mkey, p2_mad(h, 256u, jj));
                    jj += 0x10001u;
                    ++j;
                    if (j == jend && k + 1 < total) {                // rare, divergent, short
                        next_segment();
                        wn = mem.ld(j); sn = mem.sel(j);
                    }
                }
                f = (f & md.act) | sf; h1 = (h1 & md.act) | sh; mkey = (mkey & md.act) | sk;
            }
        }
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            if (!alive[l]) continue;
            const int key = (int) ((mkey >> (16 * l)) & 0xFFFFu), hl = (int) ((h1 >> (16 * l)) & 0xFFFFu);
            const int m = key >> 8, mj = key & 0xFF;
            const int jfin = end[l] > beg[l] ? end[l] : beg[l];
            if (end[l] > beg[l]) ncell += (unsigned) (end[l] - beg[l]);
            mem.st_half(end[l], l, (uint32_t) hl);
            if (jfin == qlen[l]) {
                if (hl >= gscore[l]) best_ie[l] = i;
                if (hl > gscore[l]) gscore[l] = hl;
            }
            if (m == 0) { alive[l] = false; continue; }
            if (m > best[l]) {
                best[l] = m; best_i[l] = i; best_j[l] = mj;
                int d = mj - i; d = d < 0 ? -d : d;
                if (d > max_off[l]) max_off[l] = d;
                if (0 > zthr[l]) { alive[l] = false; continue; }
            } else {
                const int di = i - best_i[l], dj = mj - best_j[l];
                const int pen = di > dj ? di - dj : dj - di;       // SIMD z-drop: no gap-extension factor, no `zdrop > 0` guard (ZSCORE8/16)
                if (best[l] - m - pen > zthr[l]) { alive[l] = false; continue; }
            }
            int j;
            for (j = beg[l]; j < end[l] && mem.ld_half(j, l) == 0u; ++j) {}
            beg[l] = j;
            for (j = end[l]; j >= beg[l] && mem.ld_half(j, l) == 0u; --j) {}
            end[l] = j + 2 < qlen[l] ? j + 2 : qlen[l];
            if (i + 1 >= tlen[l]) alive[l] = false;
        }
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        out[l].score = best[l]; out[l].qle = best_j[l] + 1; out[l].tle = best_i[l] + 1; out[l].gtle = best_ie[l] + 1;
        out[l].gscore = gscore[l]; out[l].max_off = max_off[l];
    }
    cells += ncell;
}
