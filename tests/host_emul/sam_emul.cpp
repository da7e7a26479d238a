// sam_emul.cpp — TEST-ONLY host build of the SAM-stage device logic (bwa-mem2_b200/csrc/sam_device.cuh + mate_device.cuh): mate rescue,
// primary marking, pairing, MAPQ, CIGAR / NM / MD and the SAM columns of every line of a batch of pairs, in the record layout of the
// oracle's bm2o_sam_pe so that the two can be compared field by field.  The libm tables are filled here with the host's log / erfc, as a
// host driver would.  Never part of the product.
#include <vector>
#include <string>
#include <cmath>
#include <cstring>
#include "sam_layout.cuh"
#include "../../oracle/bm2_oracle.h"

// ---- staged form of the rescue (the shape of the next kernel version): all local alignments a chunk can ask for are enumerated from the
// regions before any rescue (mate_jobs_pair_d), computed as one batch - here by the warp formulation, ksw_warp.cuh through
// ksw_warp_emul.cpp - and the per-pair block then looks its alignments up, computing one itself only if the batch does not hold it.
extern "C" int emul_ksw_warp_align2_q(int32_t qlen, const uint8_t *query, int32_t qstride, int32_t comp, int32_t tlen, const uint8_t *target, const int8_t *mat,
                                      int32_t o_del, int32_t e_del, int32_t o_ins, int32_t e_ins, int32_t xtra, int32_t *out);
#include "mate_stage.cuh"
#include "ksw_warp.cuh"
static int g_staged = 0;
static long long g_stage_stats[8];            // jobs in the batch, looked up, computed in place (not in the batch), windows that differed;
                                              // [4..6]: arena bytes, output-stripe bytes, units (pairs) of the last call
extern "C" void emul_sam_set_staged(int on) { g_staged = on; }
extern "C" void emul_sam_stage_stats(long long *out) { for (int k = 0; k < 4; ++k) out[k] = g_stage_stats[k]; }
extern "C" void emul_sam_layout_stats(long long *out) { for (int k = 0; k < 3; ++k) out[k] = g_stage_stats[4 + k]; }

// one XA entry: printed with every record of `read` whose rec_reg equals `reg`
struct EmXa { int32_t read, reg, rid, is_rev, nm, n_cigar; int64_t pos, cigar_off; };

extern "C" int emul_sam_pe(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_alnreg_t *regs, const int64_t *read_off,
                           const int32_t *lh, const double *as, int64_t id_base, bm2o_samrec **recs_out, int64_t *n_recs, uint32_t **cigar_out,
                           int64_t *n_ops_out, char **md_out, int64_t *n_md_out, int32_t **rec_reg_out, EmXa **xa_out, int64_t *n_xa_out, uint32_t **xa_cigar_out,
                           int64_t *n_xa_ops_out)
{
    ContigView cv; cv.l_pac = idx->l_pac; cv.n_seqs = idx->n_seqs; cv.ann_off = idx->ann_offset; cv.ann_len = idx->ann_len; cv.ann_alt = idx->ann_is_alt;
    SamParams p;
    p.ep.a = opt->a; p.ep.b = opt->b; p.ep.o_del = opt->o_del; p.ep.e_del = opt->e_del; p.ep.o_ins = opt->o_ins; p.ep.e_ins = opt->e_ins; p.ep.w = opt->w;
    p.ep.pen_clip5 = opt->pen_clip5; p.ep.pen_clip3 = opt->pen_clip3; p.ep.max_chain_gap = opt->max_chain_gap; p.ep.mask_level_redun = opt->mask_level_redun;
    memcpy(p.ep.mat, opt->mat, 25);
    p.T = opt->T; p.flag = opt->flag; p.min_seed_len = opt->min_seed_len; p.pen_unpaired = opt->pen_unpaired; p.mask_level = opt->mask_level;
    p.drop_ratio = opt->drop_ratio; p.mapQ_coef_len = opt->mapQ_coef_len; p.mapQ_coef_fac = opt->mapQ_coef_fac;
    p.XA_drop_ratio = opt->XA_drop_ratio; p.max_XA_hits = opt->max_XA_hits; p.max_XA_hits_alt = opt->max_XA_hits_alt;
    MatePes pes;
    for (int d = 0; d < 4; ++d) { pes.low[d] = lh[3 * d]; pes.high[d] = lh[3 * d + 1]; pes.failed[d] = lh[3 * d + 2]; }
    // host-filled libm tables
    std::vector<double> logt(1 << 16); for (size_t k = 0; k < logt.size(); ++k) logt[k] = log((double) k);      /* log(0) = -inf, as the reference would compute */
    std::vector<double> term[4];
    SamTables tb; tb.log_tab = logt.data(); tb.n_log = (int) logt.size();
    for (int d = 0; d < 4; ++d) {
        tb.pair_lo[d] = pes.low[d]; tb.pair_hi[d] = pes.failed[d] ? pes.low[d] - 1 : pes.high[d];
        for (int64_t dist = tb.pair_lo[d]; dist <= tb.pair_hi[d]; ++dist) {
            const double ns = (dist - as[2 * d]) / as[2 * d + 1];
            term[d].push_back(.721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a);
        }
        tb.pair_term[d] = term[d].data();
    }
    std::vector<bm2o_samrec> out; std::vector<uint32_t> ops_all; std::string md_all;
    std::vector<int32_t> rec_reg; std::vector<EmXa> xa; std::vector<uint32_t> xa_ops;
    int overflow = 0;
    int layout_bad = 0;
    // the staged form as sam.cu runs it: the job table of the chunk (sam_jobs_kernel = mate_jobs_list_d per pair), the alignments of the
    // table (sam_ksw_jobs_kernel = mate_job_query_d + the warp formulation), then the per-pair block with MateKswTable
    std::vector<MateJob> jobs; std::vector<MateJobRes> jres; std::vector<PairJobs> pjobs((size_t) (reads->n_reads >> 1) + 1);
    MateStats mstats = { 0, 0, 0, 0 };
    for (int k = 0; k < 8; ++k) g_stage_stats[k] = 0;
    if (g_staged && !(opt->flag & 0x20)) {
        for (int pr = 0; pr < reads->n_reads >> 1; ++pr) {
            const bm2_alnreg_t *a2[2]; int n2[2], l2[2];
            for (int i = 0; i < 2; ++i) {
                const int r = 2 * pr + i;
                a2[i] = regs + read_off[r]; n2[i] = (int) (read_off[r + 1] - read_off[r]);
                l2[i] = (int) (reads->offsets[r + 1] - reads->offsets[r]);
            }
            const int count = mate_jobs_list_d(cv, opt->min_seed_len, opt->pen_unpaired, opt->max_matesw, pes, l2, a2, n2, pr, nullptr);
            if (count > mate_jobs_bound_d(n2[0], n2[1], opt->max_matesw)) layout_bad |= 8;
            pjobs[(size_t) pr].begin = (int32_t) jobs.size(); pjobs[(size_t) pr].count = count;
            jobs.resize(jobs.size() + (size_t) count);
            mate_jobs_list_d(cv, opt->min_seed_len, opt->pen_unpaired, opt->max_matesw, pes, l2, a2, n2, pr, jobs.data() + pjobs[(size_t) pr].begin);
        }
        jres.resize(jobs.size() + 1);
        for (size_t k = 0; k < jobs.size(); ++k) {
            const MateJobQuery q = mate_job_query_d(jobs[k], reads->codes, reads->offsets, opt->a, opt->min_seed_len);
            MateJobRes o; o.score = 0; o.te = -1; o.qe = -1; o.score2 = -1; o.te2 = -1; o.tb = -1; o.qb = -1; o.valid = 0;
            if (g_staged == 2) {             // one window per thread (sam_ksw_jobs_thread_kernel): the same function, the same scratch shapes
                const int lcap = q.tlen / 2 + 2;
                std::vector<int32_t> ksw((size_t) 3 * (q.l_ms + 16)), bsc((size_t) lcap), bpos((size_t) lcap);
                std::vector<uint8_t> tmp((size_t) q.tlen + 16), rev((size_t) q.l_ms + 1);
                int ov = 0;
                const KswRes al = mate_job_align_thread_d(q, idx->ref_string + jobs[k].rb, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, ksw.data(), bsc.data(),
                                                          bpos.data(), lcap, tmp.data(), rev.data(), &ov);
                o.score = al.score; o.te = al.te; o.qe = al.qe; o.score2 = al.score2; o.te2 = al.te2; o.tb = al.tb; o.qb = al.qb; o.valid = ov ? 0 : 1;
                ++g_stage_stats[0];
            } else if (ksw_lane_fits_d(q.l_ms, BM2_KSW_CMAX) && ksw_scan_ok_d(opt->e_ins, q.l_ms)) {
                int32_t o7[7];
                emul_ksw_warp_align2_q(q.l_ms, q.q, q.stride, q.comp, q.tlen, idx->ref_string + jobs[k].rb, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, q.xtra, o7);
                o.score = o7[0]; o.te = o7[1]; o.qe = o7[2]; o.score2 = o7[3]; o.te2 = o7[4]; o.tb = o7[5]; o.qb = o7[6]; o.valid = 1;
                ++g_stage_stats[0];
            }
            jres[k] = o;
        }
    }
    for (int pr = 0; pr < reads->n_reads >> 1; ++pr) {
        const uint8_t *seq[2]; int l_seq[2], n[2];
        SamPairShape shape;
        for (int i = 0; i < 2; ++i) {
            const int r = 2 * pr + i;
            seq[i] = reads->codes + reads->offsets[r]; l_seq[i] = (int) (reads->offsets[r + 1] - reads->offsets[r]);
            n[i] = (int) (read_off[r + 1] - read_off[r]);
            sam_shape_read_d(shape, i, l_seq[i], regs + read_off[r], n[i], opt->w);
        }
        // the arena a kernel would get: capacities from sam_layout.cuh, guard words between the pieces
        const SamPairCaps caps = sam_pair_caps_d(shape, pes, opt->max_matesw, !(opt->flag & 0x20), opt->a, opt->e_del);
        g_stage_stats[4] += (long long) caps.scratch_bytes; g_stage_stats[6] += 1;
        g_stage_stats[5] += (long long) caps.recs_cap * 96 + (long long) caps.xa_cap * 40 + caps.out_ops * 4 + caps.out_md;
        std::vector<uint8_t> arena(caps.scratch_bytes + 64);
        SamArena ar;
        sam_arena_carve_d(arena.data(), caps, 1, &ar);
        for (int i = 0; i < 2; ++i) memcpy(ar.a[i], regs + read_off[2 * pr + i], sizeof(bm2_alnreg_t) * (size_t) n[i]);
        bm2_alnreg_t *ap[2] = { ar.a[0], ar.a[1] }, *bp[2] = { ar.b[0], ar.b[1] };
        if (!(opt->flag & 0x20)) {
            if (g_staged) {
                MateKswTable look = { jobs.data() + pjobs[(size_t) pr].begin, jres.data() + pjobs[(size_t) pr].begin, pjobs[(size_t) pr].count,
                                      { &p.ep, idx->ref_string, &ar.ms, &overflow }, &mstats };
                mate_rescue_pair_d(cv, p.ep, opt->min_seed_len, opt->pen_unpaired, opt->max_matesw, pes, idx->ref_string, seq, l_seq, ap, n, bp, ar.ms, look, &overflow);
            } else mate_rescue_pair_d(cv, p.ep, opt->min_seed_len, opt->pen_unpaired, opt->max_matesw, pes, idx->ref_string, seq, l_seq, ap, n, bp, ar.ms, &overflow);
        }
        if (n[0] > caps.acap[0] || n[1] > caps.acap[1]) layout_bad |= 1;
        const SamScratch &sc = ar.sc;
        long long pair_recs = 0, pair_xa = 0, pair_ops = 0, pair_md = 0;
        auto emit = [&](int i, int k, const SamRec &r, const uint32_t *ops, const char *md) {
            bm2o_samrec o; memset(&o, 0, sizeof(o));
            o.read = 2 * pr + i; o.flag = r.flag; o.rid = r.rid; o.mapq = r.mapq; o.rnext = r.rnext; o.tlen_valid = 1; o.nm = r.nm; o.score = r.score; o.sub = r.sub; o._pad = r.alt_sc;        /* _pad carries alt_sc (pa tag) in this test build */
            o.n_cigar = r.n_cigar; o.pos = r.pos; o.pnext = r.pnext; o.tlen = r.tlen; o.cigar_off = (int64_t) ops_all.size(); o.md_off = (int64_t) md_all.size();
            ops_all.insert(ops_all.end(), ops, ops + r.n_cigar + r.n_mc);          // the MC operations follow the record's own
            if (r.n_cigar) md_all += md;
            md_all.push_back('\0');
            o.n_md = (int32_t) (md_all.size() - (size_t) o.md_off);
            out.push_back(o); rec_reg.push_back(r.reg); rec_reg.push_back(r.is_alt); rec_reg.push_back(r.n_mc); ++pair_recs; pair_ops += r.n_cigar + r.n_mc; pair_md += o.n_md;
        };
        auto emit_xa = [&](int i, int reg, const SamAln &t) {
            EmXa e; e.read = 2 * pr + i; e.reg = reg; e.rid = t.rid; e.is_rev = t.is_rev; e.nm = t.nm; e.n_cigar = t.n_cigar; e.pos = t.pos; e.cigar_off = (int64_t) xa_ops.size();
            xa_ops.insert(xa_ops.end(), t.cigar, t.cigar + t.n_cigar);
            xa.push_back(e); ++pair_xa; pair_ops += t.n_cigar;
        };
        sam_pe_pair_d(p, tb, cv, pes, idx->ref_string, seq, l_seq, ap, n, (int) (id_base + pr), sc, emit, emit_xa, &overflow);
        if (!sam_arena_guards_ok_d(ar)) layout_bad |= 2;
        if (pair_recs > caps.recs_cap || pair_xa > caps.xa_cap || pair_ops > caps.out_ops || pair_md > caps.out_md) layout_bad |= 4;
    }
    g_stage_stats[1] = mstats.looked_up; g_stage_stats[2] = mstats.in_place; g_stage_stats[3] = mstats.window_moved;
    const size_t nr = out.size();
    *recs_out = (bm2o_samrec *) malloc(sizeof(bm2o_samrec) * (nr + 1)); memcpy(*recs_out, out.data(), sizeof(bm2o_samrec) * nr);
    *cigar_out = (uint32_t *) malloc(4 * (ops_all.size() + 1)); memcpy(*cigar_out, ops_all.data(), 4 * ops_all.size());
    *md_out = (char *) malloc(md_all.size() + 1); memcpy(*md_out, md_all.data(), md_all.size());
    *rec_reg_out = (int32_t *) malloc(12 * (nr + 1)); memcpy(*rec_reg_out, rec_reg.data(), 12 * nr);        /* (reg, is_alt, n_mc) per record */
    *xa_out = (EmXa *) malloc(sizeof(EmXa) * (xa.size() + 1)); memcpy(*xa_out, xa.data(), sizeof(EmXa) * xa.size());
    *xa_cigar_out = (uint32_t *) malloc(4 * (xa_ops.size() + 1)); memcpy(*xa_cigar_out, xa_ops.data(), 4 * xa_ops.size());
    *n_xa_out = (int64_t) xa.size(); *n_xa_ops_out = (int64_t) xa_ops.size();
    *n_recs = (int64_t) nr; *n_ops_out = (int64_t) ops_all.size(); *n_md_out = (int64_t) md_all.size();
    return layout_bad ? 0x1000 | layout_bad : overflow ? 0x100 | overflow : 0;
}

// The single-end branch (sam_se_read_d) in the same record layout; one read per arena (shape: read 0 of a pair without a mate).
extern "C" int emul_sam_se(const bm2_index_desc *idx, const bm2_mem_opt_t *opt, const bm2_read_batch *reads, const bm2_alnreg_t *regs, const int64_t *read_off,
                           int64_t id_base, bm2o_samrec **recs_out, int64_t *n_recs, uint32_t **cigar_out, int64_t *n_ops_out, char **md_out, int64_t *n_md_out,
                           int32_t **rec_reg_out, EmXa **xa_out, int64_t *n_xa_out, uint32_t **xa_cigar_out, int64_t *n_xa_ops_out)
{
    ContigView cv; cv.l_pac = idx->l_pac; cv.n_seqs = idx->n_seqs; cv.ann_off = idx->ann_offset; cv.ann_len = idx->ann_len; cv.ann_alt = idx->ann_is_alt;
    SamParams p;
    p.ep.a = opt->a; p.ep.b = opt->b; p.ep.o_del = opt->o_del; p.ep.e_del = opt->e_del; p.ep.o_ins = opt->o_ins; p.ep.e_ins = opt->e_ins; p.ep.w = opt->w;
    p.ep.pen_clip5 = opt->pen_clip5; p.ep.pen_clip3 = opt->pen_clip3; p.ep.max_chain_gap = opt->max_chain_gap; p.ep.mask_level_redun = opt->mask_level_redun;
    memcpy(p.ep.mat, opt->mat, 25);
    p.T = opt->T; p.flag = opt->flag; p.min_seed_len = opt->min_seed_len; p.pen_unpaired = opt->pen_unpaired; p.mask_level = opt->mask_level;
    p.drop_ratio = opt->drop_ratio; p.mapQ_coef_len = opt->mapQ_coef_len; p.mapQ_coef_fac = opt->mapQ_coef_fac;
    p.XA_drop_ratio = opt->XA_drop_ratio; p.max_XA_hits = opt->max_XA_hits; p.max_XA_hits_alt = opt->max_XA_hits_alt;
    MatePes pes; for (int d = 0; d < 4; ++d) { pes.low[d] = 0; pes.high[d] = 0; pes.failed[d] = 1; }
    std::vector<double> logt(1 << 16); for (size_t k = 0; k < logt.size(); ++k) logt[k] = log((double) k);
    SamTables tb; tb.log_tab = logt.data(); tb.n_log = (int) logt.size();
    for (int d = 0; d < 4; ++d) { tb.pair_lo[d] = 0; tb.pair_hi[d] = -1; tb.pair_term[d] = 0; }
    std::vector<bm2o_samrec> out; std::vector<uint32_t> ops_all; std::string md_all;
    std::vector<int32_t> rec_reg; std::vector<EmXa> xa; std::vector<uint32_t> xa_ops;
    int overflow = 0, layout_bad = 0;
    for (int r = 0; r < reads->n_reads; ++r) {
        const uint8_t *seq = reads->codes + reads->offsets[r]; const int l_seq = (int) (reads->offsets[r + 1] - reads->offsets[r]);
        const int n = (int) (read_off[r + 1] - read_off[r]);
        SamPairShape shape;
        sam_shape_read_d(shape, 0, l_seq, regs + read_off[r], n, opt->w);
        sam_shape_read_d(shape, 1, 0, regs, 0, opt->w);
        const SamPairCaps caps = sam_pair_caps_d(shape, pes, opt->max_matesw, false, opt->a, opt->e_del);
        std::vector<uint8_t> arena(caps.scratch_bytes + 64);
        SamArena ar;
        sam_arena_carve_d(arena.data(), caps, 1, &ar);
        memcpy(ar.a[0], regs + read_off[r], sizeof(bm2_alnreg_t) * (size_t) n);
        long long n_rec = 0, n_x = 0, n_o = 0, n_m = 0;
        auto emit = [&](int, int, const SamRec &q, const uint32_t *ops, const char *md) {
            bm2o_samrec o; memset(&o, 0, sizeof(o));
            o.read = r; o.flag = q.flag; o.rid = q.rid; o.mapq = q.mapq; o.rnext = q.rnext; o.tlen_valid = 1; o.nm = q.nm; o.score = q.score; o.sub = q.sub; o._pad = q.alt_sc;
            o.n_cigar = q.n_cigar; o.pos = q.pos; o.pnext = q.pnext; o.tlen = q.tlen; o.cigar_off = (int64_t) ops_all.size(); o.md_off = (int64_t) md_all.size();
            ops_all.insert(ops_all.end(), ops, ops + q.n_cigar + q.n_mc);
            if (q.n_cigar) md_all += md;
            md_all.push_back('\0');
            o.n_md = (int32_t) (md_all.size() - (size_t) o.md_off);
            out.push_back(o); rec_reg.push_back(q.reg); rec_reg.push_back(q.is_alt); rec_reg.push_back(q.n_mc); ++n_rec; n_o += q.n_cigar + q.n_mc; n_m += o.n_md;
        };
        auto emit_xa = [&](int, int reg, const SamAln &t) {
            EmXa e; e.read = r; e.reg = reg; e.rid = t.rid; e.is_rev = t.is_rev; e.nm = t.nm; e.n_cigar = t.n_cigar; e.pos = t.pos; e.cigar_off = (int64_t) xa_ops.size();
            xa_ops.insert(xa_ops.end(), t.cigar, t.cigar + t.n_cigar);
            xa.push_back(e); ++n_x; n_o += t.n_cigar;
        };
        sam_se_read_d(p, tb, cv, idx->ref_string, seq, l_seq, ar.a[0], n, id_base + r, ar.sc, emit, emit_xa, &overflow);
        if (!sam_arena_guards_ok_d(ar)) layout_bad |= 2;
        if (n_rec > caps.recs_cap || n_x > caps.xa_cap || n_o > caps.out_ops || n_m > caps.out_md) layout_bad |= 4;
    }
    const size_t nr = out.size();
    *recs_out = (bm2o_samrec *) malloc(sizeof(bm2o_samrec) * (nr + 1)); memcpy(*recs_out, out.data(), sizeof(bm2o_samrec) * nr);
    *cigar_out = (uint32_t *) malloc(4 * (ops_all.size() + 1)); memcpy(*cigar_out, ops_all.data(), 4 * ops_all.size());
    *md_out = (char *) malloc(md_all.size() + 1); memcpy(*md_out, md_all.data(), md_all.size());
    *rec_reg_out = (int32_t *) malloc(12 * (nr + 1)); memcpy(*rec_reg_out, rec_reg.data(), 12 * nr);        /* (reg, is_alt, n_mc) per record */
    *xa_out = (EmXa *) malloc(sizeof(EmXa) * (xa.size() + 1)); memcpy(*xa_out, xa.data(), sizeof(EmXa) * xa.size());
    *xa_cigar_out = (uint32_t *) malloc(4 * (xa_ops.size() + 1)); memcpy(*xa_cigar_out, xa_ops.data(), 4 * xa_ops.size());
    *n_xa_out = (int64_t) xa.size(); *n_xa_ops_out = (int64_t) xa_ops.size();
    *n_recs = (int64_t) nr; *n_ops_out = (int64_t) ops_all.size(); *n_md_out = (int64_t) md_all.size();
    return layout_bad ? 0x1000 | layout_bad : overflow ? 0x100 | overflow : 0;
}
