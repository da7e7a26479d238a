// bsw_col2.cuh — the extension DP of ONE job with TWO ADJACENT COLUMNS per packed 16-bit instruction
// (DPX: VIADDMNMX.S16x2, VIMNMX3.S16x2, VIADD.16x2, VIMNMX.U16x2).
//
// Same DP as bsw_extend_one (bsw.cu; reference src/bandedSWA.cpp:116-237 with the SIMD-wrapper band and z-drop,
// :2905-2926) for jobs whose scores fit 8 bits (h0 + min(qlen, tlen) * a <= 255: every job of a 2x151 bp read) and
// whose query is at most 256 columns long.  Columns 2p (low half) and 2p+1 (high half) of a row share every packed
// register.  M, E, and the row maximum are independent per column; only F runs along the row,
//     F(j+1) = max(F(j) - e_ins, M(j) - oe_ins, 0),
// and crosses from the low to the high half inside a pair and from the high half of a pair to the low half of the
// next (shifts written as multiplies: FMA pipe).  Both columns belong to the SAME job, so - unlike two jobs per thread
// (bsw_pair.cuh) - the band, the exit row and the trip count are shared: the lanes of a warp diverge no more than in
// the one-cell-per-instruction kernel.  Odd band edges are single cells (at most two per row).
//
// Per pair of cells: 1 LDS + 1 STS of the state {H_lo, E_lo, H_hi, E_hi} (4 x 8 bit), one PRMT for both substitution
// scores (the per-column selector byte is precomputed from the query), 9-10 ALU-pipe + 12 FMA-pipe instructions
// (the ALU pipe is the busier one: unpacking H, moving F between the halves and all packing are multiply-adds).
//
// Written as BM2_HD so that tests/host_emul/bsw_col2_emul.cpp runs the very same code on the CPU against the oracle.
#pragma once
#include "bsw_pair.cuh"          // p2_* packed-halfword primitives (device intrinsics + portable host definitions)

#if defined(__CUDA_ARCH__)
BM2_D uint32_t c2_shr16(uint32_t x) { return __umulhi(x, 65536u); }      // x >> 16 on the FMA pipe
#else
inline uint32_t c2_shr16(uint32_t x) { return x >> 16; }
#endif

// selector byte of one query base (codes > 4 count as 4 = N): PRMT nibbles {qb | 8 (sign of that byte), qb}
BM2_HD uint32_t c2_selector_byte(int qb) { const uint32_t q = qb > 4 ? 4u : (uint32_t) qb; return q * 0x11u + 0x80u; }

// can a job go to the column-pair kernel?  (+ the caller's class rule: 8-bit scores, qlen <= 256)
BM2_HD bool c2_params_ok(const BswParams &p) { return p2_params_ok(p); }

// Mem: ldw/stw (state word of column pair p = columns 2p, 2p+1), ldh/sth (16-bit state {H, E << 8} of one column),
// qsel(k) (selector bytes of columns 4k .. 4k+3).
// SAME_OE: o_del + e_del == o_ins + e_ins (the default scoring), one packed add per pair less.
template <bool SAME_OE, class Mem>
BM2_HD void bsw_col2_extend(const Mem &mem, const uint8_t *tptr, int tstride, int qlen, int tlen, int h0, const BswParams &p,
                            BswOut &o, unsigned long long &cells)
{
    const int oe_del = p.o_del + p.e_del, oe_ins = p.o_ins + p.e_ins, e_del = p.e_del, e_ins = p.e_ins;
    const uint32_t n_oe_del = ((uint32_t) (-oe_del) & 0xFFFFu) * 0x10001u, n_e_del = ((uint32_t) (-e_del) & 0xFFFFu) * 0x10001u;
    const uint32_t n_oe_ins = ((uint32_t) (-oe_ins) & 0xFFFFu) * 0x10001u, n_e_ins = ((uint32_t) (-e_ins) & 0xFFFFu) * 0x10001u;
    const uint32_t one = qlen >= 0 ? 1u : 0u;          // 1, opaque to the compiler: x * one + y stays a multiply-add (FMA pipe)
    // first row (bandedSWA.cpp:141-144): columns 0..qlen, E = 0
    {
        int h = h0;
        mem.sth(0, (uint32_t) h);
        h = h0 > oe_ins ? h0 - oe_ins : 0;
        for (int j = 1; j <= qlen; ++j) {
            mem.sth(j, (uint32_t) h);
            h = h > e_ins ? h - e_ins : 0;
        }
    }
    // band (SIMD wrapper arithmetic, bandedSWA.cpp:2905-2926)
    int w = p.w;
    const BswQuirk qk = bsw_quirk(qlen, tlen, h0, p);
    {
        unsigned t1 = ((unsigned) (qlen * p.a) + (unsigned) (p.end_bonus - p.o_ins)) & qk.band_mask;
        int max_ins = (int) (t1 / (unsigned) e_ins) + 1; if (max_ins < 1) max_ins = 1;
        unsigned t2 = ((unsigned) (qlen * p.a) + (unsigned) (p.end_bonus - p.o_del)) & qk.band_mask;
        int max_del = (int) (t2 / (unsigned) e_del) + 1; if (max_del < 1) max_del = 1;
        if (w > max_ins) w = max_ins;
        if (w > max_del) w = max_del;
    }
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    unsigned long long ncell = 0;
    int tb_next = tlen > 0 ? (int) tptr[0] : 0;           // the target base of a row is fetched one row ahead (global-memory latency)
    for (int i = 0; i < tlen; ++i) {
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1;
        if (beg == 0) { h1 = h0 - (p.o_del + e_del * (i + 1)); if (h1 < 0) h1 = 0; }
        else h1 = 0;
        const int tb = tb_next;
        if (i + 1 < tlen) tb_next = (int) tptr[(long long) (i + 1) * tstride];
        const uint32_t tbl = p2_score_table(tb, p.a, p.b);
        int f = 0, key1 = 0;
        uint32_t key2 = 0;
        // one cell (odd band edges): plain 32-bit arithmetic
        auto cell1 = [&](const int jj) {
            const uint32_t v = mem.ldh(jj);
            const int hd = (int) (v & 0xFFu);
            int e = (int) (v >> 8);
            const uint32_t qb = (mem.qsel(jj >> 2) >> (8 * (jj & 3))) & 0xFu;
            const int s = (int) p2_prmt(tbl, 0xFFFFFFFFu, qb * 0x1111u + 0x8880u);
            int M = hd ? hd + s : 0; if (M < 0) M = 0;
            int h = M > e ? M : e; if (f > h) h = f;
            int t = M - oe_del; if (t < 0) t = 0;
            e -= e_del; if (t > e) e = t;
            mem.sth(jj, (uint32_t) (h1 | (e << 8)));
            t = M - oe_ins; if (t < 0) t = 0;
            f -= e_ins; if (t > f) f = t;
            h1 = h;
            const int k = h * 256 + jj;
            if (k > key1) key1 = k;
        };
        int j = beg;
        if (j < end && (j & 1)) { cell1(j); ++j; }
        {
            int pp = j >> 1;
            const int pe = end >> 1;
            if (pp < pe) {
                // F, H1: running F / H(i, j-1) of the next column in the LOW half (high half 0)
                uint32_t F = (uint32_t) f, H1 = (uint32_t) h1, jj2 = (uint32_t) (2 * pp) * 0x10001u + 0x10000u;
                auto pair = [&](const int q, const uint32_t sel) {
                    const uint32_t wv = mem.ldw(q);
                    const uint32_t e = p2_prmt(wv, 0u, 0x4341u);
                    const uint32_t hd = p2_mad(e, 0xFFFFFF00u, wv);                          // wv - (e << 8) on the FMA pipe
                    const uint32_t s = p2_prmt(tbl, 0xFFFFFFFFu, sel);                       // low 16 selector bits: columns 2q, 2q+1
                    const uint32_t M = p2_addmin_relu(hd, s, p2_mad(hd, 128u, 0u));          // hd ? max(hd + s, 0) : 0   (s <= 127)
                    const uint32_t td = p2_add(M, n_oe_del);
                    const uint32_t en = p2_addmax_relu(e, n_e_del, td);
                    const uint32_t ti = SAME_OE ? td : p2_add(M, n_oe_ins);
                    // F(2q+1) = max(F(2q) - e_ins, M(2q) - oe_ins, 0) in the HIGH half: both operands moved up by multiplies
                    const uint32_t fa = p2_addmax_relu(p2_mad(F, 65536u, 0u), n_e_ins, p2_mad(ti, 65536u, 0u));      // low half: 0
                    const uint32_t f2 = p2_mad(F, one, fa);                                  // {F(2q), F(2q+1)}
                    const uint32_t fb = p2_addmax_relu(f2, n_e_ins, ti);                     // high half: F(2q+2)
                    const uint32_t h = p2_max3(M, e, f2);
                    const uint32_t hs = p2_mad(h, 65536u, H1);                               // {H(i, 2q-1), H(i, 2q)}
                    mem.stw(q, p2_mad(en, 256u, hs));
                    F = c2_shr16(fb);
                    H1 = c2_shr16(h);
                    key2 = p2_maxu(key2, p2_mad(h, 256u, jj2));
                    jj2 = p2_mad(one, 0x00020002u, jj2);
                };
                auto quad = [&](const int k) {                                              // columns 4k .. 4k+3
                    const uint32_t sw = mem.qsel(k);
                    pair(2 * k, sw);
                    pair(2 * k + 1, c2_shr16(sw));
                };
                if (pp & 1) { pair(pp, c2_shr16(mem.qsel(pp >> 1))); ++pp; }
                int k = pp >> 1;
                const int ke = pe >> 1;
                for (; k + 2 <= ke; k += 2) { quad(k); quad(k + 1); }
                if (k < ke) { quad(k); ++k; }
                pp = 2 * k;
                if (pp < pe) pair(pp, mem.qsel(pp >> 1));
                f = (int) F; h1 = (int) H1;
                j = 2 * pe;
            }
        }
        if (j < end) { cell1(j); ++j; }
        {
            const int ka = (int) (key2 & 0xFFFFu), kb = (int) (key2 >> 16);
            if (ka > key1) key1 = ka;
            if (kb > key1) key1 = kb;
        }
        const int m = key1 >> 8, mj = key1 & 0xFF;
        if (end > beg) ncell += (unsigned) (end - beg);
#ifdef BM2_COL2_TRACE
        BM2_COL2_TRACE(beg, end);                              // lane-utilisation studies (tests/host_emul/bsw_col2_emul.cpp)
#endif
        mem.sth(end, (uint32_t) h1);
        if (j == qlen) {
            if (h1 >= gscore) best_ie = i;
            if (h1 > gscore) gscore = h1;
        }
        if (m == 0) break;
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            int d = mj - i; d = d < 0 ? -d : d;
            if (d > max_off) max_off = d;
            if (0 > qk.zthr) break;
        } else {
            const int di = i - best_i, dj = mj - best_j;
            const int pen = di > dj ? di - dj : dj - di;       // SIMD z-drop: no e_del/e_ins factor, no `zdrop > 0` guard (ZSCORE8/16)
            if (best - m - pen > qk.zthr) break;
        }
        for (j = beg; j < end && mem.ldh(j) == 0u; ++j) {}
        beg = j;
        j = end;                                               // column `end` holds {h1, 0} (written above): no load for the usual case
        if (h1 == 0) for (--j; j >= beg && mem.ldh(j) == 0u; --j) {}
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    o.score = best; o.qle = best_j + 1; o.tle = best_i + 1; o.gtle = best_ie + 1; o.gscore = gscore; o.max_off = max_off;
    cells += ncell;
}
