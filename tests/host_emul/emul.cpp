// emul.cpp — TEST-ONLY host emulation of the seam-2 GPU pipeline.
//
// Compiles the per-thread device logic of bwa-mem2_b200/csrc/{fm,chain,ext}_device.cuh with g++ and
// runs it read after read with the SAME flat data layout and stage order as pipeline.cu, so that the
// kernels' control logic can be checked against the oracle on a machine without a GPU.  The banded
// extension DP itself is NOT emulated here (its CUDA kernel is checked on the GPU by
// tests/test_bsw_gpu.py); this harness calls the oracle's DP for that step.  Never part of the product.
#include <vector>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include "fm_device.cuh"
#include "chain_device.cuh"
static long long g_gs_calls = 0, g_gs_cells = 0;      // statistics of the score-only global alignment inside the tail
#define BM2_TRACE_GLOBAL_SCORE(qlen, tlen, w) do { ++g_gs_calls; g_gs_cells += (long long) (tlen) * ((2 * (w) + 1) < (qlen) ? (2 * (w) + 1) : (qlen)); } while (0)
#include "ext_device.cuh"
#include "../../oracle/bm2_oracle.h"

static int g_smem_text = 0;      // emul_set_smem_text: the unique-interval shortcut of fm_forward (fm.text) in the host runs
extern "C" void emul_set_smem_text(int on) { g_smem_text = on; }

namespace {
struct Views { FmIndexView fm; ContigView cv; SmemParams sp; ChainParams cp; ExtParams ep; SwParams sw; const bm2_index_desc *idx; const bm2_mem_opt_t *opt; };

Views make_views(const bm2_index_desc *idx, const bm2_mem_opt_t *o) {
    Views v;
    v.fm.cp_occ = idx->cp_occ; v.fm.sa_ms = idx->sa_ms_byte; v.fm.sa_ls = idx->sa_ls_word; v.fm.sentinel = idx->sentinel_index;
    for (int i = 0; i < 5; ++i) v.fm.count[i] = idx->count[i];
    v.cv.l_pac = idx->l_pac; v.cv.n_seqs = idx->n_seqs; v.cv.ann_off = idx->ann_offset; v.cv.ann_len = idx->ann_len; v.cv.ann_alt = idx->ann_is_alt;
    v.sp.min_seed_len = o->min_seed_len; v.sp.split_len = (int) (o->min_seed_len * o->split_factor + .499);
    v.sp.split_width = o->split_width; v.sp.max_mem_intv = (int) o->max_mem_intv;
    v.cp.w = o->w; v.cp.max_chain_gap = o->max_chain_gap; v.cp.max_occ = o->max_occ; v.cp.min_chain_weight = o->min_chain_weight;
    v.cp.max_chain_extend = o->max_chain_extend; v.cp.min_seed_len = o->min_seed_len; v.cp.mask_level = o->mask_level; v.cp.drop_ratio = o->drop_ratio;
    v.ep.a = o->a; v.ep.b = o->b; v.ep.o_del = o->o_del; v.ep.e_del = o->e_del; v.ep.o_ins = o->o_ins; v.ep.e_ins = o->e_ins; v.ep.w = o->w;
    v.ep.pen_clip5 = o->pen_clip5; v.ep.pen_clip3 = o->pen_clip3; v.ep.max_chain_gap = o->max_chain_gap; v.ep.mask_level_redun = o->mask_level_redun;
    memcpy(v.ep.mat, o->mat, 25);
    v.sw.a = o->a; v.sw.o_del = o->o_del; v.sw.e_del = o->e_del; v.sw.o_ins = o->o_ins; v.sw.e_ins = o->e_ins; memcpy(v.sw.mat, o->mat, 25);
    if (g_smem_text || getenv("BM2_EMUL_SMEM_TEXT")) { v.fm.text = idx->ref_string; v.fm.text_len = 2 * idx->l_pac; }      // (the env form: a whole test run with the shortcut on)
    v.idx = idx; v.opt = o;
    return v;
}

struct Stage1 {      // sorted SMEMs + per-read ranges + SA of every seed slot
    std::vector<bm2_smem> smems; std::vector<int64_t> read_smem_off; std::vector<int64_t> slot_off; std::vector<int64_t> sa;
    int64_t n_ext = 0, n_lf = 0;
};

void stage_smem_sa(const Views &v, const bm2_mem_opt_t *o, const bm2_read_batch *rb, Stage1 &s) {
    int max_len = 0;
    for (int r = 0; r < rb->n_reads; ++r) max_len = std::max<int>(max_len, (int) (rb->offsets[r + 1] - rb->offsets[r]));
    std::vector<FmPrev> scratch(max_len + 2);
    struct Search { int x, min_intv; std::vector<FmPrev> list; };
    for (int r = 0; r < rb->n_reads; ++r) {
        const uint8_t *q = rb->codes + rb->offsets[r]; int len = (int) (rb->offsets[r + 1] - rb->offsets[r]);
        unsigned n_ext = 0;
        QPlain qq = { q };
        // same phase order as the kernels: forward chain of pass 1, backward tasks, re-seed forward, backward, pass 3
        std::vector<Search> tasks;
        auto sink = [&](int x, int min_intv, const FmPrev *list, int nl) { Search t; t.x = x; t.min_intv = min_intv; t.list.assign(list, list + nl); tasks.push_back(t); };
        fm_forward(v.fm, qq, len, 0, 1, false, scratch.data(), sink, n_ext);
        std::vector<std::pair<int, int>> reseeds;
        bool pass1 = true;
        auto emit = [&](int m, int n, int64_t k, int64_t l, int64_t ss) {
            bm2_smem x; x.rid = r; x.m = m; x.n = n; x.k = k; x.l = l; x.s = ss; s.smems.push_back(x);
            if (pass1 && n + 1 - m >= v.sp.split_len && ss <= v.sp.split_width) reseeds.push_back(std::make_pair((n + 1 + m) >> 1, (int) (ss + 1)));
        };
        for (Search &t : tasks) fm_backward_rows(v.fm, qq, t.x, t.min_intv, v.sp.min_seed_len, t.list.data(), (int) t.list.size(), emit, n_ext);
        pass1 = false;
        std::vector<Search> tasks2;
        auto sink2 = [&](int x, int min_intv, const FmPrev *list, int nl) { Search t; t.x = x; t.min_intv = min_intv; t.list.assign(list, list + nl); tasks2.push_back(t); };
        for (auto &rs : reseeds) fm_forward(v.fm, qq, len, rs.first, rs.second, true, scratch.data(), sink2, n_ext);
        for (Search &t : tasks2) fm_backward_rows(v.fm, qq, t.x, t.min_intv, v.sp.min_seed_len, t.list.data(), (int) t.list.size(), emit, n_ext);
        fm_smem_pass3(v.fm, qq, len, v.sp, emit, n_ext);
        s.n_ext += n_ext;
    }
    std::stable_sort(s.smems.begin(), s.smems.end(), [](const bm2_smem &a, const bm2_smem &b) {
        uint64_t ka = (uint64_t) a.rid << 32 | (uint64_t) a.m << 16 | a.n, kb = (uint64_t) b.rid << 32 | (uint64_t) b.m << 16 | b.n;
        return ka < kb;
    });
    s.read_smem_off.assign(rb->n_reads + 1, 0);
    for (auto &x : s.smems) s.read_smem_off[x.rid + 1]++;
    for (int r = 0; r < rb->n_reads; ++r) s.read_smem_off[r + 1] += s.read_smem_off[r];
    s.slot_off.assign(s.smems.size() + 1, 0);
    for (size_t i = 0; i < s.smems.size(); ++i) s.slot_off[i + 1] = s.slot_off[i] + std::min<int64_t>(s.smems[i].s, o->max_occ);
    s.sa.resize(s.slot_off.back());
    for (size_t i = 0; i < s.smems.size(); ++i) {
        const bm2_smem &x = s.smems[i];
        int64_t step = x.s > o->max_occ ? x.s / o->max_occ : 1;
        for (int64_t t = 0; t < s.slot_off[i + 1] - s.slot_off[i]; ++t) { int lf = 0; s.sa[s.slot_off[i] + t] = fm_sa_of_row(v.fm, x.k + t * step, &lf); s.n_lf += lf; }
    }
}

struct Stage2 {      // finalized chains per read (flat, per-read stripes compacted here)
    std::vector<bm2_chain> chains; std::vector<bm2_seed> seeds; std::vector<int64_t> read_chain_off;
    std::vector<int> n_left, n_right;
};

void stage_chain(const Views &v, const bm2_read_batch *rb, const Stage1 &s1, Stage2 &s2) {
    s2.read_chain_off.assign(rb->n_reads + 1, 0); s2.n_left.assign(rb->n_reads, 0); s2.n_right.assign(rb->n_reads, 0);
    for (int r = 0; r < rb->n_reads; ++r) {
        int len = (int) (rb->offsets[r + 1] - rb->offsets[r]);
        int64_t sb = s1.read_smem_off[r], se = s1.read_smem_off[r + 1];
        // reference quirk: a 512-read block with exactly one SMEM yields no chain (src/bwamem.cpp:835)
        int b0 = (r / 512) * 512, b1 = std::min(rb->n_reads, b0 + 512);
        bool skip = (s1.read_smem_off[b1] - s1.read_smem_off[b0]) <= 1;
        if (se > sb && !skip && len >= v.sp.min_seed_len) {
            int64_t slots = s1.slot_off[se] - s1.slot_off[sb];
            std::vector<WSeed> ws(slots + 1); std::vector<WChain> wc(slots + 1); std::vector<int32_t> ord(slots + 1), srt(slots + 1), kv(slots + 1); std::vector<int64_t> ordpos(slots + 1); std::vector<FltRec> flt(slots + 1);
            ChainStripe st = { ws.data(), wc.data(), ord.data(), ordpos.data(), srt.data(), kv.data(), flt.data() };
            float frac = 0;
            int nk = chain_read_d(v.cv, v.cp, s1.smems.data() + sb, (int) (se - sb), s1.sa.data() + s1.slot_off[sb], len, st, &frac);
            {
                const double min_l = v.opt->min_chain_weight ? 1.1f * v.opt->min_chain_weight : 5.5f * log((double) len);
                if (!(min_l > 0.05f * len))
                    chain_flt_seeds_d(v.cv, v.sw, v.idx->ref_string, len, rb->codes + rb->offsets[r], (int) (v.opt->a * min_l + .499), st, nk);
            }
            std::vector<bm2_chain> oc(nk + 1); std::vector<bm2_seed> os(slots + 1);
            int ns = 0, nl = 0, nr = 0;
            chain_finalize_d(st, nk, frac, r, len, oc.data(), os.data(), &ns, &nl, &nr);
            int64_t seed_base = (int64_t) s2.seeds.size();
            for (int k = 0; k < nk; ++k) { oc[k].seed_off += (int32_t) seed_base; s2.chains.push_back(oc[k]); }
            for (int k = 0; k < ns; ++k) { os[k].chain += (int32_t) s2.read_chain_off[r]; s2.seeds.push_back(os[k]); }
            s2.n_left[r] = nl; s2.n_right[r] = nr;
        }
        s2.read_chain_off[r + 1] = (int64_t) s2.chains.size();
    }
}
}  // namespace

extern "C" {

int64_t emul_collect_smems(const bm2_index_desc *idx, const bm2_mem_opt_t *o, const bm2_read_batch *rb, bm2_smem **out, int64_t *n_ext) {
    Views v = make_views(idx, o); Stage1 s1; stage_smem_sa(v, o, rb, s1);
    *out = (bm2_smem *) malloc(sizeof(bm2_smem) * (s1.smems.size() + 1));
    memcpy(*out, s1.smems.data(), sizeof(bm2_smem) * s1.smems.size());
    if (n_ext) *n_ext = s1.n_ext;
    return (int64_t) s1.smems.size();
}

int emul_seed_chain(const bm2_index_desc *idx, const bm2_mem_opt_t *o, const bm2_read_batch *rb, bm2_chain **chains, int64_t *n_chains,
                    bm2_seed **seeds, int64_t *n_seeds, int64_t **read_off) {
    Views v = make_views(idx, o); Stage1 s1; stage_smem_sa(v, o, rb, s1); Stage2 s2; stage_chain(v, rb, s1, s2);
    *chains = (bm2_chain *) malloc(sizeof(bm2_chain) * (s2.chains.size() + 1)); memcpy(*chains, s2.chains.data(), sizeof(bm2_chain) * s2.chains.size());
    *seeds = (bm2_seed *) malloc(sizeof(bm2_seed) * (s2.seeds.size() + 1)); memcpy(*seeds, s2.seeds.data(), sizeof(bm2_seed) * s2.seeds.size());
    *read_off = (int64_t *) malloc(8 * (rb->n_reads + 1)); memcpy(*read_off, s2.read_chain_off.data(), 8 * (rb->n_reads + 1));
    *n_chains = (int64_t) s2.chains.size(); *n_seeds = (int64_t) s2.seeds.size();
    return 0;
}

int emul_seed_chain_extend(const bm2_index_desc *idx, const bm2_mem_opt_t *o, const bm2_read_batch *rb, bm2_alnreg_t **regs_out,
                           int64_t *n_regs, int64_t **read_off) {
    Views v = make_views(idx, o); Stage1 s1; stage_smem_sa(v, o, rb, s1); Stage2 s2; stage_chain(v, rb, s1, s2);
    const int n = rb->n_reads;
    // reg / job offsets (device: exclusive scans)
    std::vector<int64_t> reg_off(n + 1, 0), left_off(n + 1, 0), right_off(n + 1, 0);
    int max_chain = 1, max_len = 1;
    for (int r = 0; r < n; ++r) {
        int64_t nr = 0;
        for (int64_t c = s2.read_chain_off[r]; c < s2.read_chain_off[r + 1]; ++c) { nr += s2.chains[c].n_seeds; max_chain = std::max(max_chain, s2.chains[c].n_seeds); }
        reg_off[r + 1] = reg_off[r] + nr; left_off[r + 1] = left_off[r] + s2.n_left[r]; right_off[r + 1] = right_off[r] + s2.n_right[r];
        max_len = std::max<int>(max_len, (int) (rb->offsets[r + 1] - rb->offsets[r]));
    }
    std::vector<bm2_alnreg_t> regs(reg_off[n] + 1); std::vector<int32_t> reg_chain(reg_off[n] + 1), reg_seed(reg_off[n] + 1);
    std::vector<ExtJobRec> left(left_off[n] + 1), right(right_off[n] + 1); std::vector<int32_t> left_reg(left_off[n] + 1), right_reg(right_off[n] + 1);
    std::vector<uint64_t> srt(max_chain + 1);
    for (int r = 0; r < n; ++r) {
        int64_t cb = s2.read_chain_off[r], ce = s2.read_chain_off[r + 1];
        if (ce == cb) continue;
        // chains of the read address seeds through absolute seed_off: pass seeds base 0
        ext_build_read_d(v.cv, v.ep, s2.chains.data() + cb, (int) (ce - cb), s2.seeds.data(), (int) (rb->offsets[r + 1] - rb->offsets[r]),
                         rb->offsets[r], cb, reg_off[r], regs.data() + reg_off[r], reg_chain.data() + reg_off[r], reg_seed.data() + reg_off[r],
                         left.data() + left_off[r], left_reg.data() + left_off[r], right.data() + right_off[r], right_reg.data() + right_off[r], srt.data());
    }
    auto run_phase = [&](std::vector<ExtJobRec> &jobs, std::vector<int32_t> &job_reg, int64_t nj, int is_right) {
        std::vector<int> todo(nj); for (int64_t i = 0; i < nj; ++i) todo[i] = (int) i;
        for (int t = 0; t < 2 && !todo.empty(); ++t) {
            int w = o->w << t;
            std::vector<int> retry;
            for (int ji : todo) {
                ExtJobRec &j = jobs[ji];
                bm2_alnreg_t &a = regs[job_reg[ji]];
                if (is_right && t == 0) j.h0 = a.score;
                std::vector<uint8_t> q(j.qlen), tg(j.tlen);
                for (int i = 0; i < j.qlen; ++i) q[i] = rb->codes[j.qoff + (int64_t) i * j.qstride];
                for (int i = 0; i < j.tlen; ++i) tg[i] = idx->ref_string[j.toff + (int64_t) i * j.tstride];
                bm2o_bsw_params bp = { o->a, o->b, o->o_del, o->e_del, o->o_ins, o->e_ins, o->zdrop, is_right ? o->pen_clip3 : o->pen_clip5, 1 };
                int32_t out[6];
                bm2o_bsw_extend(q.data(), j.qlen, tg.data(), j.tlen, w, j.h0, &bp, out);
                const bm2_chain &c = s2.chains[reg_chain[job_reg[ji]]];
                int rd = c.seqid; int l_query = (int) (rb->offsets[rd + 1] - rb->offsets[rd]);
                bool ok = ext_fold_d(v.ep, a, is_right, j.h0, out[0], out[1], out[2], out[3], out[4], out[5], w, t == 1, l_query,
                                     s2.seeds.data() + c.seed_off, c.n_seeds);
                if (!ok) retry.push_back(ji);
            }
            todo.swap(retry);
        }
    };
    run_phase(left, left_reg, left_off[n], 0);
    run_phase(right, right_reg, right_off[n], 1);
    std::vector<bm2_alnreg_t> out; std::vector<int64_t> off(n + 1, 0);
    std::vector<int32_t> srt2(reg_off[n] + max_chain + 1), he(2 * (max_len + 2)); std::vector<PfBox> box(reg_off[n] + 1);
    for (int r = 0; r < n; ++r) {
        int64_t cb = s2.read_chain_off[r], ce = s2.read_chain_off[r + 1];
        int nreg = (int) (reg_off[r + 1] - reg_off[r]);
        int l_query = (int) (rb->offsets[r + 1] - rb->offsets[r]);
        if (ce > cb) {
            g_gs_calls = 0; g_gs_cells = 0;
            const auto t0 = std::chrono::steady_clock::now();
            ext_postfilter_read_d(v.ep, s2.chains.data() + cb, (int) (ce - cb), s2.seeds.data(), l_query, regs.data() + reg_off[r], nreg,
                                  reg_seed.data() + reg_off[r], srt2.data() + reg_off[r], box.data() + reg_off[r]);
            const auto t1 = std::chrono::steady_clock::now();
            int m = ext_tail_read_d(v.cv, v.ep, idx->ref_string, rb->codes + rb->offsets[r], regs.data() + reg_off[r], nreg, he.data(), srt2.data() + reg_off[r], reinterpret_cast<TailSortKey *>(box.data() + reg_off[r]));
            const auto t2 = std::chrono::steady_clock::now();
            if (getenv("BM2_EMUL_TAIL_STATS") && nreg > 24)
                fprintf(stderr, "TAILSTAT %d nreg %d final %d pf_us %.1f tail_us %.1f gs_calls %lld gs_cells %lld\n", r, nreg, m,
                        std::chrono::duration<double, std::micro>(t1 - t0).count(), std::chrono::duration<double, std::micro>(t2 - t1).count(), g_gs_calls, g_gs_cells);
            for (int i = 0; i < m; ++i) out.push_back(regs[reg_off[r] + i]);
        }
        off[r + 1] = (int64_t) out.size();
    }
    *regs_out = (bm2_alnreg_t *) malloc(sizeof(bm2_alnreg_t) * (out.size() + 1)); memcpy(*regs_out, out.data(), sizeof(bm2_alnreg_t) * out.size());
    *read_off = (int64_t *) malloc(8 * (n + 1)); memcpy(*read_off, off.data(), 8 * (n + 1));
    *n_regs = (int64_t) out.size();
    return 0;
}

// the chain tree alone: keys in insertion order; lower[i] = id of the key kb_intervalp would return before key i is inserted
// (-1: none), order[] = ids in key order at the end
void bm2e_chain_tree(const int64_t *keys, int n, int32_t *lower, int32_t *order)
{
    std::vector<int32_t> ord(n + 1); std::vector<int64_t> ordpos(n + 1);
    int height = 0, root_n = 0;
    for (int m = 0; m < n; ++m) {
        const int64_t k = keys[m];
        int lo = 0, hi = m;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (ordpos[mid] < k) lo = mid + 1; else hi = mid; }
        const int at_low = (lo < m && ordpos[lo] == k) ? chain_tree_equal_d(ord.data(), ordpos.data(), m, height, k, lo) : lo - 1;
        lower[m] = at_low >= 0 ? chain_ord_id(ord[at_low]) : -1;
        const int at = chain_tree_put_d(ord.data(), ordpos.data(), m, height, root_n, k, lo);
        for (int j = m; j > at; --j) { ord[j] = ord[j - 1]; ordpos[j] = ordpos[j - 1]; }
        ord[at] = m; ordpos[at] = k;
    }
    for (int m = 0; m < n; ++m) order[m] = chain_ord_id(ord[m]);
}
}
