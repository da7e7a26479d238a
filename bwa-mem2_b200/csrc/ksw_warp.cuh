// ksw_warp.cuh — the local alignment of mate rescue, one WINDOW PER WARP.  Launched by ksw.cu (bm2_ksw_align2, a batch of requests) and by sam.cu's
// staged rescue (sam_ksw_jobs_kernel); the per-pair kernel's own fallback stays ksw_device.cuh's one-thread sweep.
//
// Same function as ksw_pass_d (ksw_device.cuh: the reference's striped ksw_u8 / ksw_i16, src/ksw.cpp:111-316, with its first-pass E /
// row maximum and its padding).  There a row is one sequential sweep over the query with two insertion registers: F of the segment
// (restarts at every segment start of the reference's striping) and the complete F.  Both are max-plus prefix scans over
//     base[k] = max(H(i-1, k-1) + S, E(i, k))            (an opening from an F-derived H never beats extending that F, o_ins >= 0)
//     F(k) = max(0, max over j < k of base[j] - o_ins - e_ins * (k - j))     (for the segment F: j inside k's segment)
// so a row splits over the 32 lanes: lane l owns C = ceil(nlen / 32) neighbouring columns (H, E in registers), computes base[] and the
// F values leaving its block (zero entering).  The decay between two blocks is e_ins per column and the columns are static, so with
//     w_j = v_j + e_ins * (end column of block j)            F entering block l = max(0, max_{j<l} w_j - e_ins * col0_l)
// the scan is a plain exclusive prefix MAXIMUM over the lanes (one shuffle + one max per step; no (decay, value) pairs).  The segment F only
// sees lanes since the last segment start, and those are static too: keys w_j + BIG * (segment starts in columns [0, end of block j)) make
// every lane of an older segment lose against any lane of the current one, so the same prefix maximum serves - two 32-bit scans per row.
// A second local pass finishes the cells.  The row maximum is a warp max (redux.sync).
// Rows stay synchronous, so the reference's early stops, its score-2 list and the copy of the best row are as in the sweep.
//
// The per-lane phases are plain BM2_HD functions; tests/host_emul/ksw_warp_emul.cpp drives them with a loop over 32 lane states and
// plain loops for the scan / reductions (tests/test_oracle_ksw.py: equal to the oracle and to the reference's golden vectors);
// ksw_pass_warp_d drives the same phases with __shfl_sync.
#pragma once
#include "ksw_device.cuh"

#define BM2_KSW_CMAX 16                          // columns per lane: queries up to 32 * 16 - 15 = 497 bases
#define BM2_KSW_BIG (1 << 20)                    // segment offset of the scan keys: above any w (ksw_scan_ok_d)
#define BM2_KSW_NONE (-(1 << 30))                // scan identity

// T = compile-time capacity of a lane (columns): every loop over the lane's columns is `for c < T, if c < ncol` fully unrolled, so the
// lane's state stays in registers (a runtime trip count would put the arrays in local memory).  ksw_lane_width_d picks T from the shape.
template <int T> struct KswLaneT {
    int32_t H[T], E[T], Hbest[T];                // completed H of the previous row, E, H of the best row
    int32_t base[T];
    uint32_t prof[T];                            // query profile of the lane's columns: byte t = score of the column against target base t (0..3);
    uint32_t profn[(T + 3) / 4];                 // byte c & 3 of word c >> 2 = score against target base 4 (N); padding columns (beyond qlen) score 0
    int col0, ncol;                              // first column, columns owned (0 for lanes beyond nlen)
    unsigned segmask;                            // bit c: column col0 + c starts a segment of the reference's striping
    int endcol, seg_in, seg_end;                 // col0 + ncol; segment starts in columns [0, col0) and [0, endcol)
    int hlast;                                   // completed H of the lane's last column (the diagonal input of the next lane's next row)
};
typedef KswLaneT<BM2_KSW_CMAX> KswLane;
#if defined(__CUDA_ARCH__)
#define BM2_UNROLL _Pragma("unroll")
#else
#define BM2_UNROLL
#endif
// the lane capacities compiled: 151-bp reads need 5 columns per lane (160 padded columns), 250-bp reads 8
BM2_HD int ksw_lane_width_d(int C) { return C <= 5 ? 5 : C <= 8 ? 8 : BM2_KSW_CMAX; }
// does a query of qlen bases fit lanes of capacity tmax in either score class (16 or 8 columns per segment group)?
BM2_HD bool ksw_lane_fits_d(int qlen, int tmax) { return qlen > 0 && ((qlen + 15) / 16 * 16 + 31) / 32 <= tmax && ((qlen + 7) / 8 * 8 + 31) / 32 <= tmax; }
// the kernel instance (5, 8 or BM2_KSW_CMAX) for queries up to max_qlen bases; 0: too long for this formulation
BM2_HD int ksw_kernel_width_d(int max_qlen) { return ksw_lane_fits_d(max_qlen, 5) ? 5 : ksw_lane_fits_d(max_qlen, 8) ? 8 : ksw_lane_fits_d(max_qlen, BM2_KSW_CMAX) ? BM2_KSW_CMAX : 0; }
struct KswShape { int size, qlen, p, slen, nlen, C, shift; int oe_del, e_del, oe_ins, e_ins; };
struct KswSummary { int v_seg, v_full; };        // F of the segment / complete F: leaving a block (phase A), scan keys, entering a block (phase B)
// the scan keys stay below BM2_KSW_BIG: scores < 2^15 (both classes), decay e_ins per column
BM2_HD bool ksw_scan_ok_d(int e_ins, int qlen) { return e_ins > 0 && (long long) e_ins * (qlen + 48) + 32768 < BM2_KSW_BIG; }

BM2_HD KswShape ksw_shape_d(int size, int qlen, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins) {
    KswShape s; s.size = size; s.qlen = qlen; s.p = size == 1 ? 16 : 8;
    s.slen = (qlen + s.p - 1) / s.p; s.nlen = s.slen * s.p; s.C = (s.nlen + 31) / 32;
    int shift = 127;
    for (int a = 0; a < 25; ++a) if (mat[a] < shift) shift = mat[a];
    s.shift = (256 - shift) & 0xff;
    s.oe_del = o_del + e_del; s.e_del = e_del; s.oe_ins = o_ins + e_ins; s.e_ins = e_ins;
    return s;
}

// comp: the query is read complemented (with qstride < 0 from its last base: the reverse complement, without a copy)
// mat is read here only (once per pass): the kernels stage it in shared memory, the rows then work from the profile registers
template <int T>
BM2_HD void ksw_lane_init_d(const KswShape &s, int lane, const uint8_t *query, int qstride, const int8_t *mat, KswLaneT<T> &L, int comp = 0) {
    L.col0 = lane * s.C;
    L.ncol = L.col0 >= s.nlen ? 0 : (s.nlen - L.col0 < s.C ? s.nlen - L.col0 : s.C);
    L.hlast = 0; L.segmask = 0;
    L.endcol = L.col0 + L.ncol;
    L.seg_in = (L.col0 + s.slen - 1) / s.slen; L.seg_end = (L.endcol + s.slen - 1) / s.slen;
    BM2_UNROLL
    for (int c = 0; c < T; ++c) {
        L.H[c] = 0; L.E[c] = 0; L.Hbest[c] = 0; L.base[c] = 0; L.prof[c] = 0;
        if ((c & 3) == 0) L.profn[c >> 2] = 0;
        if (c >= L.ncol) continue;
        const int k = L.col0 + c;
        if (k % s.slen == 0) L.segmask |= 1u << c;
        if (k >= s.qlen) continue;                                               // padding column, substitution score 0
        int b = query[(long long) k * qstride];
        if (b > 4) b = 4;
        if (comp) b = b < 4 ? 3 - b : 4;
        uint32_t w = 0;
        for (int t = 0; t < 4; ++t) w |= (uint32_t) (uint8_t) mat[t * 5 + b] << (8 * t);
        L.prof[c] = w;
        L.profn[c >> 2] |= (uint32_t) (uint8_t) mat[20 + b] << (8 * (c & 3));
    }
}

// phase A: base[] of the row and the block's scan summary.  diag_in = H(i-1) of the column left of the block (0 for lane 0).
template <int T>
BM2_HD KswSummary ksw_lane_phase_a_d(const KswShape &s, int tbase, int diag_in, KswLaneT<T> &L) {      // tbase: the row's target code (0..4)
    KswSummary m;
    int diag = L.col0 == 0 ? 0 : diag_in, fs = 0, ff = 0;
    BM2_UNROLL
    for (int c = 0; c < T; ++c) {
        if (c >= L.ncol) continue;
        if ((L.segmask >> c) & 1) fs = 0;                                      // a segment starts here: nothing from the left survives
        const int sc = tbase < 4 ? (int) (int8_t) (L.prof[c] >> (8 * tbase)) : (int) (int8_t) (L.profn[c >> 2] >> (8 * (c & 3)));
        int h = diag; diag = L.H[c];
        if (s.size == 1) { h = h + sc + s.shift; if (h > 255) h = 255; h -= s.shift; if (h < 0) h = 0; }
        else { h = h + sc; if (h > 32767) h = 32767; }
        const int e = L.E[c];
        if (e > h) h = e;
        L.base[c] = h;
        const int open = h - s.oe_ins > 0 ? h - s.oe_ins : 0;
        fs -= s.e_ins; if (fs < 0) fs = 0; if (open > fs) fs = open;
        ff -= s.e_ins; if (ff < 0) ff = 0; if (open > ff) ff = open;
    }
    m.v_seg = fs; m.v_full = ff;
    return m;
}

// scan keys of a block (see the header) and the way back from the exclusive prefix maxima to the F values entering a block
template <int T>
BM2_HD KswSummary ksw_scan_keys_d(const KswShape &s, const KswLaneT<T> &L, const KswSummary &m) {
    KswSummary k;
    k.v_full = L.ncol ? m.v_full + s.e_ins * L.endcol : BM2_KSW_NONE;
    k.v_seg = L.ncol ? m.v_seg + s.e_ins * L.endcol + BM2_KSW_BIG * L.seg_end : BM2_KSW_NONE;
    return k;
}
template <int T>
BM2_HD KswSummary ksw_scan_entering_d(const KswShape &s, const KswLaneT<T> &L, const KswSummary &pm) {      // pm: maxima over the lanes to the left
    KswSummary in;
    in.v_full = pm.v_full - s.e_ins * L.col0;
    in.v_seg = pm.v_seg - BM2_KSW_BIG * L.seg_in - s.e_ins * L.col0;      // a lane of an older segment gives a negative value here: 0 after phase B's clamp
    return in;
}

// phase B: the cells of the row with the F values entering the block (in.v_seg / in.v_full of the exclusive scan).  Returns the
// block's maximum of the first-pass H; leaves the completed H in L.H and the new E in L.E.
template <int T>
BM2_HD int ksw_lane_phase_b_d(const KswShape &s, const KswSummary &in, KswLaneT<T> &L) {
    int fs = in.v_seg > 0 ? in.v_seg : 0, ff = in.v_full > 0 ? in.v_full : 0, rowmax = 0, hl = 0;
    BM2_UNROLL
    for (int c = 0; c < T; ++c) {
        if (c >= L.ncol) continue;
        if ((L.segmask >> c) & 1) fs = 0;
        int h = L.base[c];
        if (fs > h) h = fs;                                                      // first-pass H
        if (h > rowmax) rowmax = h;
        { int t = h - s.oe_del; if (t < 0) t = 0; int e = L.E[c] - s.e_del; if (e < 0) e = 0; L.E[c] = e > t ? e : t; }
        const int open = h - s.oe_ins > 0 ? h - s.oe_ins : 0;
        fs -= s.e_ins; if (fs < 0) fs = 0; if (open > fs) fs = open;
        const int hc = ff > h ? ff : h;                                          // completed H
        L.H[c] = hc; hl = hc;
        ff -= s.e_ins; if (ff < 0) ff = 0; if (open > ff) ff = open;
    }
    L.hlast = hl;
    return rowmax;
}

// row bookkeeping shared by both drivers.  Every lane keeps the same row state; the last list entry lives in the state (registers), so no lane
// reads the list while another may be writing it, and only the writer lane (lane 0 on the device) stores.
struct KswRowState { int gmax, te, n_b, last_sc, last_pos; bool stop; };
BM2_HD void ksw_row_end_d(const KswShape &s, int i, int rowmax, int minsc, int endsc, KswRowState &st, int32_t *bsc, int32_t *bpos, int bcap, int *overflow, bool *took,
                          bool writer) {
    *took = false;
    if (rowmax >= minsc) {
        if (st.n_b == 0 || st.last_pos + 1 != i) {
            if (st.n_b < bcap) { if (writer) { bsc[st.n_b] = rowmax; bpos[st.n_b] = i; } ++st.n_b; st.last_sc = rowmax; st.last_pos = i; } else *overflow |= 32;
        } else if (st.last_sc < rowmax) { if (writer) { bsc[st.n_b - 1] = rowmax; bpos[st.n_b - 1] = i; } st.last_sc = rowmax; st.last_pos = i; }
    }
    if (rowmax > st.gmax) {
        st.gmax = rowmax; st.te = i; *took = true;
        if (s.size == 1 ? (st.gmax + s.shift >= 255 || st.gmax >= endsc) : (st.gmax >= endsc)) st.stop = true;
    }
}

#if defined(__CUDACC__)
// One pass by a full warp (all 32 lanes call it with the same arguments).  bsc / bpos: the warp's score-2 list (global or shared).
// rev_upto >= 0: the rows run over target[rev_upto], target[rev_upto - 1], ..., target[0], target[rev_upto + 1], ... (ksw_align2's second pass
// reverses the prefix in place and still walks all tlen rows, src/ksw.cpp:366-371).
template <int T>
__device__ __forceinline__ KswRes ksw_pass_warp_t(int size, int qlen, const uint8_t *query, int qstride, int comp, int tlen, const uint8_t *target, int rev_upto, const int8_t *mat,
                                         int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *bsc, int32_t *bpos, int bcap, int *overflow)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const KswShape s = ksw_shape_d(size, qlen, mat, o_del, e_del, o_ins, e_ins);
    int qmax = 0;
    for (int a = 0; a < 25; ++a) if (mat[a] > qmax) qmax = mat[a];
    const int minsc = (xtra & BM2_KSW_XSUBO) ? xtra & 0xffff : 0x10000, endsc = (xtra & BM2_KSW_XSTOP) ? xtra & 0xffff : 0x10000;
    KswLaneT<T> L;
    ksw_lane_init_d(s, lane, query, qstride, mat, L, comp);
    KswRowState st; st.gmax = 0; st.te = -1; st.n_b = 0; st.last_sc = 0; st.last_pos = -2; st.stop = false;
    int ov = 0;
    for (int i = 0; i < tlen && !st.stop; ++i) {
        const int tbase = target[i <= rev_upto ? rev_upto - i : i];
        const int diag_in = __shfl_up_sync(full, L.hlast, 1);                    // H(i-1) of the left neighbour's last column
        KswSummary m = ksw_lane_phase_a_d(s, tbase > 4 ? 4 : tbase, diag_in, L);
        const KswSummary key = ksw_scan_keys_d(s, L, m);
        KswSummary pm;                                                           // exclusive prefix maxima: shift by one lane, then Hillis-Steele
        pm.v_seg = __shfl_up_sync(full, key.v_seg, 1); pm.v_full = __shfl_up_sync(full, key.v_full, 1);
        if (lane == 0) { pm.v_seg = BM2_KSW_NONE; pm.v_full = BM2_KSW_NONE; }
        BM2_UNROLL
        for (int d = 1; d < 32; d <<= 1) {
            const int os = __shfl_up_sync(full, pm.v_seg, d), of = __shfl_up_sync(full, pm.v_full, d);
            if (lane >= d) { pm.v_seg = os > pm.v_seg ? os : pm.v_seg; pm.v_full = of > pm.v_full ? of : pm.v_full; }
        }
        const KswSummary in = ksw_scan_entering_d(s, L, pm);
        int rowmax = ksw_lane_phase_b_d(s, in, L);
        rowmax = __reduce_max_sync(full, rowmax);
        bool took;                                                               // all lanes keep the same row state; lane 0 writes the list
        ksw_row_end_d(s, i, rowmax, minsc, endsc, st, bsc, bpos, bcap, &ov, &took, lane == 0);
        if (took) {
            BM2_UNROLL
            for (int c = 0; c < T; ++c) L.Hbest[c] = L.H[c];
        }
        __syncwarp(full);
    }
    KswRes r; r.score = size == 1 ? (st.gmax + s.shift < 255 ? st.gmax : 255) : st.gmax; r.te = st.te; r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = -1; r.qb = -1;
    if (size == 2 || r.score != 255) {
        int mx = -1, pos = 0x7fffffff;
        BM2_UNROLL
        for (int c = 0; c < T; ++c) if (c < L.ncol && L.Hbest[c] > mx) { mx = L.Hbest[c]; pos = L.col0 + c; }
        for (int d = 16; d > 0; d >>= 1) {
            const int omx = __shfl_xor_sync(full, mx, d), opos = __shfl_xor_sync(full, pos, d);
            if (omx > mx || (omx == mx && opos < pos)) { mx = omx; pos = opos; }
        }
        r.qe = pos;
        if (st.n_b) {
            const int dd = (r.score + qmax - 1) / qmax, low = st.te - dd, high = st.te + dd;
            int s2 = -1, t2 = -1;
            for (int k = lane; k < st.n_b; k += 32) if ((bpos[k] < low || bpos[k] > high) && bsc[k] > s2) { s2 = bsc[k]; t2 = bpos[k]; }
            for (int d = 16; d > 0; d >>= 1) {                                       // first entry among equal scores, as the serial scan
                const int os = __shfl_xor_sync(full, s2, d), ot = __shfl_xor_sync(full, t2, d);
                if (os > s2 || (os == s2 && os >= 0 && ot < t2)) { s2 = os; t2 = ot; }
            }
            r.score2 = s2; r.te2 = t2;
        }
    }
    if (lane == 0 && ov) *overflow |= ov;
    return r;
}

// TMAX: the largest lane capacity this kernel instance carries (its register budget); the caller guarantees ksw_lane_fits_d(qlen, TMAX)
template <int TMAX>
__device__ __forceinline__ KswRes ksw_pass_warp_d(int size, int qlen, const uint8_t *query, int qstride, int comp, int tlen, const uint8_t *target, int rev_upto,
                                         const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *bsc, int32_t *bpos, int bcap, int *overflow)
{
    const int p = size == 1 ? 16 : 8, nlen = (qlen + p - 1) / p * p;
    const int w = ksw_lane_width_d((nlen + 31) / 32);                         // warp-uniform
    if (TMAX > 8 && w > 8)
        return ksw_pass_warp_t<(TMAX > 8) ? BM2_KSW_CMAX : 5>(size, qlen, query, qstride, comp, tlen, target, rev_upto, mat, o_del, e_del, o_ins, e_ins, xtra, bsc, bpos, bcap, overflow);
    if (TMAX > 5 && w > 5)
        return ksw_pass_warp_t<(TMAX > 5) ? 8 : 5>(size, qlen, query, qstride, comp, tlen, target, rev_upto, mat, o_del, e_del, o_ins, e_ins, xtra, bsc, bpos, bcap, overflow);
    return ksw_pass_warp_t<5>(size, qlen, query, qstride, comp, tlen, target, rev_upto, mat, o_del, e_del, o_ins, e_ins, xtra, bsc, bpos, bcap, overflow);
}

// ksw_align2 (src/ksw.cpp:324-381) by a full warp: forward pass, then the reversed prefixes to find the start.
// The query is query[0], query[qstride], ... (complemented if comp): (ms + l_ms - 1, -1, 1) is the reverse complement of ms.
template <int TMAX>
__device__ __forceinline__ KswRes ksw_align2_warp_d(int qlen, const uint8_t *query, int qstride, int comp, int tlen, const uint8_t *target, const int8_t *mat, int o_del, int e_del, int o_ins,
                                           int e_ins, int xtra, int32_t *bsc, int32_t *bpos, int bcap, int *overflow)
{
    const int size = (xtra & BM2_KSW_XBYTE) ? 1 : 2;
    KswRes r = ksw_pass_warp_d<TMAX>(size, qlen, query, qstride, comp, tlen, target, -1, mat, o_del, e_del, o_ins, e_ins, xtra, bsc, bpos, bcap, overflow);
    if ((xtra & BM2_KSW_XSTART) == 0 || ((xtra & BM2_KSW_XSUBO) && r.score < (xtra & 0xffff))) return r;
    int ov2 = 0;
    __syncwarp(0xffffffffu);
    const KswRes rr = ksw_pass_warp_d<TMAX>(size, r.qe + 1, query + (long long) r.qe * qstride, -qstride, comp, tlen, target, r.te, mat, o_del, e_del, o_ins, e_ins, BM2_KSW_XSTOP | r.score, bsc, bpos, bcap, &ov2);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return r;
}
#endif
