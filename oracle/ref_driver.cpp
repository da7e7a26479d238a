/* ref_driver.cpp — TEST INFRASTRUCTURE (oracle/): our own driver around the UNMODIFIED reference.
 *
 * Links the reference's objects (oracle/_ref/<isa>/libbwa.a, built from /root/reference by
 * oracle/Makefile) and calls the reference's own `main_mem` (src/fastmap.cpp:616).  Reference
 * sources are never patched: the seams are hooked at LINK time with `-Wl,--wrap=<symbol>`:
 *
 *   kt_for                         (src/kthread.cpp:81)   the three phases of mem_process_seqs
 *   FMI_search::sortSMEMs          (src/FMI_search.cpp:1008) end of mem_collect_smem -> SMEM dump
 *   BandedPairWiseSW::getScores16 / getScores8 / scalarBandedSWAWrapper (src/bandedSWA.cpp)
 *
 * Modes (environment):
 *   BM2_MODE=ref      (default) pass-through; optional dumps + timing of the hot path
 *   BM2_MODE=hotpath  run only worker_bwt + worker_aln (the hot path), skip worker_sam (timing)
 *   BM2_MODE=gpu      replace worker_bwt + worker_aln by libbm2b200.so:bm2_seed_chain_extend,
 *                     keep the reference's pestat + worker_sam  => drop-in SAM check
 *   BM2_MODE=gpu_bsw  replace only the BSW calls by bm2_extend_pairs (config 2)
 *   BM2_LIB=<path to libbm2b200.so>       (gpu modes)
 *   BM2_DUMP_PREFIX=<p>   write <p>.smem.bin <p>.chains.bin <p>.regs.bin <p>.bsw.bin  (use -t 1)
 *                         + <p>.pestat.bin (mem_pestat's result per chunk)
 *   BM2_DUMP_REGS=<file>  write ONLY the regs dump (same format as <p>.regs.bin); written by the kt_for hook between the phases,
 *                         so it is valid with any -t (the other dumps come from inside the worker threads)
 *   BM2_STATS=<file>      JSON with wall seconds of the phases
 *   BM2_REPEAT=<K>        (hotpath mode) run worker_bwt + worker_aln K times on every chunk inside ONE process (one index load);
 *                         BM2_STATS then carries "rep_s": the K wall times of (worker_bwt + worker_aln)
 *
 * `ref_driver cigar <index prefix> <requests.bin> <out.bin>` calls the reference's own bwa_gen_cigar2 (src/bwa.cpp:260)
 * on every request of a binary file (pins the CIGAR/NM/MD restatement of the oracle):
 *   requests.bin: int64 n; then per request int64 rb, re; int32 w, l_query; uint8 query[l_query]
 *   out.bin:      per request int32 score, n_cigar, nm, n_md; uint32 cigar[n_cigar]; char md[n_md]
 *                 (score = INT32_MIN when the reference returns without setting it)
 *
 * `ref_driver ksw <requests.bin> <out.bin>` calls the reference's own ksw_align2 (src/ksw.cpp:324, the local alignment of mate
 * rescue) with the default scoring on every request:
 *   requests.bin: int64 n; then per request int32 qlen, tlen, xtra; uint8 query[qlen]; uint8 target[tlen]
 *   out.bin:      per request 7 int32: score, te, qe, score2, te2, tb, qb  (kswr_t)
 *
 * `ref_driver matesw <index prefix> <in.bin> <out.bin>` runs the rescue block of mem_sam_pe (src/bwamem_pair.cpp:378-412, MATE_SORT=0)
 * with the reference's own mem_matesw (:150) on read pairs of a file (mem_matesw is called inside its own translation unit, so a
 * link-time wrap cannot see it):
 *   in.bin:  4 x (int32 low, high, failed); int64 n_pairs; per pair, for read 0 then read 1: int32 l_seq; uint8 seq[l_seq]; int32 n; mem_alnreg_t regs[n]
 *   out.bin: per mem_matesw call: int32 pair, i, j, returned n, n_after; mem_alnreg_t regs[n_after] (the mate's regs after the call)
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include <sstream>
#include <fstream>
#include <iostream>
#include <algorithm>
#include <chrono>
#include <dlfcn.h>
#include <unistd.h>
#include <pthread.h>
#include <zlib.h>

#define private public
#define protected public
#include "FMI_search.h"
#include "bandedSWA.h"
#include "bwamem.h"
#include "fastmap.h"
#include "kthread.h"
#include "main.h"
#include "bwa.h"
#include "bntseq.h"
#include "ksw.h"
#undef private
#undef protected

#include "bm2_b200.h"

extern uint64_t proc_freq, tprof[LIM_R][LIM_C];
extern char *bwa_pg;

static_assert(sizeof(bm2_alnreg_t) == sizeof(mem_alnreg_t), "alnreg layout");
static_assert(offsetof(bm2_alnreg_t, score) == offsetof(mem_alnreg_t, score), "alnreg.score");
static_assert(offsetof(bm2_alnreg_t, seedlen0) == offsetof(mem_alnreg_t, seedlen0), "alnreg.seedlen0");
static_assert(offsetof(bm2_alnreg_t, frac_rep) == offsetof(mem_alnreg_t, frac_rep), "alnreg.frac_rep");
static_assert(offsetof(bm2_alnreg_t, hash) == offsetof(mem_alnreg_t, hash), "alnreg.hash");
static_assert(offsetof(bm2_alnreg_t, flg) == offsetof(mem_alnreg_t, flg), "alnreg.flg");
static_assert(sizeof(bm2_seqpair) == sizeof(SeqPair), "SeqPair layout");
static_assert(sizeof(bm2_mem_opt_t) == sizeof(mem_opt_t), "mem_opt_t layout");
static_assert(offsetof(bm2_mem_opt_t, mat) == offsetof(mem_opt_t, mat), "mem_opt_t.mat");
static_assert(offsetof(bm2_mem_opt_t, max_mem_intv) == offsetof(mem_opt_t, max_mem_intv), "mem_opt_t.max_mem_intv");
static_assert(offsetof(bm2_mem_opt_t, mask_level) == offsetof(mem_opt_t, mask_level), "mem_opt_t.mask_level");
static_assert(sizeof(bm2_smem) == sizeof(SMEM), "SMEM layout");
static_assert(sizeof(bm2_cp_occ) == sizeof(CP_OCC), "CP_OCC layout");

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

enum Mode { M_REF, M_HOTPATH, M_GPU, M_GPU_BSW };
static Mode g_mode = M_REF;
static std::string g_dump, g_stats;
static FILE *f_smem = 0, *f_chain = 0, *f_regs = 0, *f_bsw = 0;
static int g_phase = 0;
static double t_phase[3] = {0, 0, 0};
static int64_t g_reads = 0, g_bsw_pairs = 0;
static double t_bsw = 0;
static int64_t g_smem_read_base = 0;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static int g_repeat = 1;
static std::vector<double> g_rep_s;
static void (*g_func0)(void *, int, int, int) = 0;
static double g_t0_first = 0;

/* ---- libbm2b200.so bindings (exactly the stub INTEGRATION.md shows) ---- */
static void *g_lib = 0;
static bm2_ctx *g_ctx = 0;
static decltype(&bm2_create) p_create;
static decltype(&bm2_destroy) p_destroy;
static decltype(&bm2_last_error) p_last_error;
static decltype(&bm2_extend_pairs) p_extend_pairs;
static decltype(&bm2_seed_chain_extend) p_seed_chain_extend;

static void load_lib() {
    if (g_lib) return;
    const char *lib = getenv("BM2_LIB");
    if (!lib) { fprintf(stderr, "[ref_driver] BM2_LIB not set\n"); exit(2); }
    g_lib = dlopen(lib, RTLD_NOW | RTLD_GLOBAL);
    if (!g_lib) { fprintf(stderr, "[ref_driver] dlopen %s: %s\n", lib, dlerror()); exit(2); }
#define SYM(p, name) p = (decltype(p)) dlsym(g_lib, name); if (!p) { fprintf(stderr, "missing %s\n", name); exit(2); }
    SYM(p_create, "bm2_create"); SYM(p_destroy, "bm2_destroy"); SYM(p_last_error, "bm2_last_error");
    SYM(p_extend_pairs, "bm2_extend_pairs"); SYM(p_seed_chain_extend, "bm2_seed_chain_extend");
#undef SYM
}

static void make_ctx(const mem_opt_t *opt, FMI_search *fmi, uint8_t *ref_string) {
    if (g_ctx) return;
    load_lib();
    bm2_index_desc d; memset(&d, 0, sizeof(d));
    std::vector<int64_t> off; std::vector<int32_t> len, alt;
    if (fmi) {
        d.reference_seq_len = fmi->reference_seq_len;
        for (int i = 0; i < 5; i++) d.count[i] = fmi->count[i];
        d.sentinel_index = fmi->sentinel_index;
        d.cp_occ = (const bm2_cp_occ *) fmi->cp_occ;
        d.sa_ms_byte = fmi->sa_ms_byte; d.sa_ls_word = fmi->sa_ls_word;
        d.ref_string = ref_string;
        const bntseq_t *bns = fmi->idx->bns;
        d.l_pac = bns->l_pac; d.n_seqs = bns->n_seqs;
        for (int i = 0; i < bns->n_seqs; i++) {
            off.push_back(bns->anns[i].offset); len.push_back(bns->anns[i].len); alt.push_back(bns->anns[i].is_alt);
        }
        d.ann_offset = off.data(); d.ann_len = len.data(); d.ann_is_alt = alt.data();
    }
    int rc = p_create(&g_ctx, 0, fmi ? &d : NULL, (const bm2_mem_opt_t *) opt);
    if (rc) { fprintf(stderr, "[ref_driver] bm2_create failed: %s\n", p_last_error(NULL)); exit(3); }
}

/* ---- kt_for hook --------------------------------------------------------------------------- */
extern "C" void __real__Z6kt_forPFvPviiiES_i(void (*func)(void *, int, int, int), void *data, int n);

static void dump_chains(worker_t *w, int n) {
    for (int i = 0; i < n; i++) {
        mem_chain_v *cv = &w->chain_ar[i];
        int32_t nc = cv->n; fwrite(&nc, 4, 1, f_chain);
        for (int j = 0; j < nc; j++) {
            mem_chain_t *c = &cv->a[j];
            int32_t hdr[8] = { c->n, c->rid, (int32_t) c->w, (int32_t) c->kept, c->first, (int32_t) c->is_alt, c->seqid, 0 };
            fwrite(hdr, 4, 8, f_chain); fwrite(&c->frac_rep, 4, 1, f_chain); fwrite(&c->pos, 8, 1, f_chain);
            for (int k = 0; k < c->n; k++) {
                mem_seed_t *s = &c->seeds[k];
                fwrite(&s->rbeg, 8, 1, f_chain);
                int32_t v[3] = { s->qbeg, s->len, s->score }; fwrite(v, 4, 3, f_chain);
            }
        }
    }
    fflush(f_chain);
}

static void dump_regs(worker_t *w, int n) {
    for (int i = 0; i < n; i++) {
        mem_alnreg_v *rv = &w->regs[i];
        int32_t nr = rv->n; fwrite(&nr, 4, 1, f_regs);
        for (int j = 0; j < nr; j++) {
            mem_alnreg_t *a = &rv->a[j];
            int64_t r[2] = { a->rb, a->re }; fwrite(r, 8, 2, f_regs);
            int32_t v[16] = { a->qb, a->qe, a->rid, a->score, a->truesc, a->sub, a->alt_sc, a->csub, a->sub_n,
                              a->w, a->seedcov, a->secondary, a->secondary_all, a->seedlen0, a->n_comp, a->is_alt };
            fwrite(v, 4, 16, f_regs); fwrite(&a->frac_rep, 4, 1, f_regs); fwrite(&a->hash, 8, 1, f_regs);
        }
    }
    fflush(f_regs);
}

static void write_stats() {
    if (g_stats.empty()) return;
    FILE *f = fopen(g_stats.c_str(), "w");
    if (!f) return;
    fprintf(f, "{\"reads\": %ld, \"t_bwt\": %.6f, \"t_aln\": %.6f, \"t_sam\": %.6f, \"bsw_pairs\": %ld, \"t_bsw\": %.6f, \"rep_s\": [",
            (long) g_reads, t_phase[0], t_phase[1], t_phase[2], (long) g_bsw_pairs, t_bsw);
    for (size_t i = 0; i < g_rep_s.size(); ++i) fprintf(f, "%s%.6f", i ? ", " : "", g_rep_s[i]);
    fprintf(f, "]}\n");
    fclose(f);
}

static void gpu_hotpath(worker_t *w, int n) {
    const mem_opt_t *opt = w->opt;
    make_ctx(opt, w->fmi, w->ref_string);
    std::vector<int64_t> off(n + 1, 0);
    for (int i = 0; i < n; i++) off[i + 1] = off[i] + w->seqs[i].l_seq;
    std::vector<uint8_t> codes(off[n] + 1);
    for (int i = 0; i < n; i++) {       /* src/bwamem.cpp:992-1000: in-place 2-bit encoding */
        char *seq = w->seqs[i].seq; int len = w->seqs[i].l_seq;
        for (int j = 0; j < len; j++) {
            seq[j] = seq[j] < 4 ? seq[j] : nst_nt4_table[(int) seq[j]];
            codes[off[i] + j] = (uint8_t) seq[j];
        }
    }
    bm2_read_batch rb = { n, codes.data(), off.data() };
    bm2_reg_result rr;
    int rc = p_seed_chain_extend(g_ctx, &rb, &rr);
    if (rc) { fprintf(stderr, "[ref_driver] bm2_seed_chain_extend: %s\n", p_last_error(g_ctx)); exit(3); }
    for (int i = 0; i < n; i++) {
        int64_t b = rr.read_off[i], e = rr.read_off[i + 1];
        mem_alnreg_v *rv = &w->regs[i];
        rv->n = rv->m = e - b;
        rv->a = (mem_alnreg_t *) calloc(rv->m ? rv->m : 1, sizeof(mem_alnreg_t));
        if (e > b) memcpy(rv->a, rr.regs + b, (e - b) * sizeof(mem_alnreg_t));
    }
}

extern "C" void __wrap__Z6kt_forPFvPviiiES_i(void (*func)(void *, int, int, int), void *data, int n) {
    worker_t *w = (worker_t *) data;
    int ph = g_phase; g_phase = (g_phase + 1) % 3;
    double t0 = now_s();
    if (g_mode == M_GPU && ph == 0) gpu_hotpath(w, n);
    else if (g_mode == M_GPU && ph == 1) { /* regs already filled */ }
    else if (g_mode == M_HOTPATH && ph == 2) {
        for (int i = 0; i < n; i++) { free(w->regs[i].a); w->regs[i].a = 0; w->regs[i].n = 0; }
    } else {
        __real__Z6kt_forPFvPviiiES_i(func, data, n);
    }
    t_phase[ph] += now_s() - t0;
    if (ph == 0) { g_reads += n; g_func0 = func; g_t0_first = now_s() - t0; if (f_chain && g_mode != M_GPU) dump_chains(w, n); }
    if (ph == 1 && g_mode == M_HOTPATH && g_repeat > 1 && g_func0) {
        /* timing repetitions of the hot path on this chunk: the reads are already encoded in place (idempotent), mem_kernel2_core
           frees the chains and re-initialises regs[] itself (src/bwamem.cpp:1104-1107, :1126-1139); only regs[].a is ours to free */
        if (g_rep_s.empty()) g_rep_s.push_back(g_t0_first + (now_s() - t0));
        else g_rep_s[0] += g_t0_first + (now_s() - t0);
        for (int r = 1; r < g_repeat; ++r) {
            for (int i = 0; i < n; i++) { free(w->regs[i].a); w->regs[i].a = 0; w->regs[i].n = 0; w->regs[i].m = 0; }
            double t1 = now_s();
            __real__Z6kt_forPFvPviiiES_i(g_func0, data, n);
            __real__Z6kt_forPFvPviiiES_i(func, data, n);
            double dt = now_s() - t1;
            if ((int) g_rep_s.size() <= r) g_rep_s.push_back(dt); else g_rep_s[r] += dt;
        }
    }
    if (ph == 1 && f_regs) dump_regs(w, n);
    if (ph == 2) write_stats();
}

/* ---- SMEM hook ----------------------------------------------------------------------------- */
extern "C" void __real__ZN10FMI_search9sortSMEMsEP11smem_structPliii(FMI_search *self, SMEM *a, int64_t *num, int32_t nreads, int32_t rl, int nt);
extern "C" void __wrap__ZN10FMI_search9sortSMEMsEP11smem_structPliii(FMI_search *self, SMEM *a, int64_t *num, int32_t nreads, int32_t rl, int nt) {
    __real__ZN10FMI_search9sortSMEMsEP11smem_structPliii(self, a, num, nreads, rl, nt);
    if (f_smem) {
        pthread_mutex_lock(&g_mu);
        int64_t hdr[3] = { g_smem_read_base, nreads, num[0] };
        fwrite(hdr, 8, 3, f_smem);
        fwrite(a, sizeof(SMEM), num[0], f_smem);
        g_smem_read_base += nreads;
        fflush(f_smem);
        pthread_mutex_unlock(&g_mu);
    }
}

/* ---- BSW hooks ----------------------------------------------------------------------------- */
static void bsw_dump(int kind, BandedPairWiseSW *self, SeqPair *p, const std::vector<SeqPair> &before,
                     uint8_t *ref, uint8_t *qer, int n, int w) {
    pthread_mutex_lock(&g_mu);
    int32_t hdr[12] = { 0x31575342, kind, n, w, self->end_bonus, self->zdrop, self->o_del, self->e_del, self->o_ins,
                        self->e_ins, self->w_match, self->w_mismatch };
    fwrite(hdr, 4, 12, f_bsw);
    for (int i = 0; i < n; i++) {
        int32_t v[9] = { before[i].len1, before[i].len2, before[i].h0, p[i].score, p[i].tle, p[i].gtle, p[i].qle,
                         p[i].gscore, p[i].max_off };
        fwrite(v, 4, 9, f_bsw);
        fwrite(ref + before[i].idr, 1, before[i].len1, f_bsw);
        fwrite(qer + before[i].idq, 1, before[i].len2, f_bsw);
    }
    fflush(f_bsw);
    pthread_mutex_unlock(&g_mu);
}

#define BSW_HOOK(MANGLED, KIND, NTTYPE)                                                                     \
    extern "C" void __real_##MANGLED(BandedPairWiseSW *self, SeqPair *p, uint8_t *ref, uint8_t *qer,       \
                                     int32_t n, NTTYPE nt, int32_t w);                                     \
    extern "C" void __wrap_##MANGLED(BandedPairWiseSW *self, SeqPair *p, uint8_t *ref, uint8_t *qer,       \
                                     int32_t n, NTTYPE nt, int32_t w) {                                    \
        std::vector<SeqPair> before;                                                                        \
        if (f_bsw) before.assign(p, p + n);                                                                 \
        double t0 = now_s();                                                                                \
        if (g_mode == M_GPU_BSW) {                                                                          \
            /* one bm2_ctx: calls are serialised by the caller (contract in INTEGRATION.md) */             \
            static pthread_mutex_t gpu_mu = PTHREAD_MUTEX_INITIALIZER;                                      \
            pthread_mutex_lock(&gpu_mu);                                                                    \
            make_ctx(NULL, NULL, NULL);                                                                     \
            /* bm2_create(opt=NULL) takes defaults; scoring comes from the reference object */             \
            int rc = p_extend_pairs(g_ctx, (bm2_seqpair *) p, ref, qer, n, w, self->end_bonus);             \
            if (rc) { fprintf(stderr, "[ref_driver] bm2_extend_pairs: %s\n", p_last_error(g_ctx)); exit(3); } \
            pthread_mutex_unlock(&gpu_mu);                                                                  \
        } else {                                                                                            \
            __real_##MANGLED(self, p, ref, qer, n, nt, w);                                                  \
        }                                                                                                   \
        pthread_mutex_lock(&g_mu); t_bsw += now_s() - t0; g_bsw_pairs += n; pthread_mutex_unlock(&g_mu);    \
        if (f_bsw && n > 0) bsw_dump(KIND, self, p, before, ref, qer, n, w);                                \
    }

BSW_HOOK(_ZN16BandedPairWiseSW11getScores16EP10dnaSeqPairPhS2_iti, 16, uint16_t)
BSW_HOOK(_ZN16BandedPairWiseSW10getScores8EP10dnaSeqPairPhS2_iti, 8, uint16_t)
BSW_HOOK(_ZN16BandedPairWiseSW22scalarBandedSWAWrapperEP10dnaSeqPairPhS2_iii, 1, int)

/* ---- mate rescue: mem_pestat (src/bwamem_pair.cpp:88) and mem_matesw (:150) wrapped at link time ---- */
static FILE *f_pestat = 0;
extern "C" void __real__Z10mem_pestatPK9mem_opt_tliPK12mem_alnreg_vP12mem_pestat_t(const mem_opt_t *, int64_t, int, const mem_alnreg_v *, mem_pestat_t *);
extern "C" void __wrap__Z10mem_pestatPK9mem_opt_tliPK12mem_alnreg_vP12mem_pestat_t(const mem_opt_t *opt, int64_t l_pac, int n, const mem_alnreg_v *regs, mem_pestat_t *pes) {
    __real__Z10mem_pestatPK9mem_opt_tliPK12mem_alnreg_vP12mem_pestat_t(opt, l_pac, n, regs, pes);
    if (f_pestat) {
        int32_t nn = n; fwrite(&nn, 4, 1, f_pestat);
        for (int d = 0; d < 4; ++d) { int32_t v[3] = { pes[d].low, pes[d].high, pes[d].failed }; double w[2] = { pes[d].avg, pes[d].std }; fwrite(v, 4, 3, f_pestat); fwrite(w, 8, 2, f_pestat); }
        fflush(f_pestat);
    }
}
static int cigar_mode(int argc, char *argv[]) {
    if (argc < 5) { fprintf(stderr, "usage: ref_driver cigar <index prefix> <requests.bin> <out.bin>\n"); return 1; }
    bntseq_t *bns = bns_restore(argv[2]);
    if (!bns) return 1;
    std::vector<uint8_t> pac((size_t) (bns->l_pac / 4 + 1));
    if (fread(pac.data(), 1, pac.size(), bns->fp_pac) == 0) return 1;
    mem_opt_t *opt = mem_opt_init();
    FILE *fi = fopen(argv[3], "rb"), *fo = fopen(argv[4], "wb");
    if (!fi || !fo) return 1;
    int64_t n = 0;
    if (fread(&n, 8, 1, fi) != 1) return 1;
    std::vector<uint8_t> q;
    for (int64_t i = 0; i < n; ++i) {
        int64_t rb, re; int32_t w, lq;
        if (fread(&rb, 8, 1, fi) != 1 || fread(&re, 8, 1, fi) != 1 || fread(&w, 4, 1, fi) != 1 || fread(&lq, 4, 1, fi) != 1) return 1;
        q.resize((size_t) (lq > 0 ? lq : 0) + 1);
        if (lq > 0 && fread(q.data(), 1, (size_t) lq, fi) != (size_t) lq) return 1;
        int score = INT32_MIN, n_cigar = 0, nm = 0;
        uint32_t *cigar = bwa_gen_cigar2(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, w, bns->l_pac, pac.data(), lq, q.data(), rb, re,
                                         &score, &n_cigar, &nm);
        /* the MD string follows the n_cigar operations in the returned block (src/bwa.cpp:307, :334) */
        int32_t n_md = 0;
        const char *md = nullptr;
        if (cigar && nm >= 0) { md = (const char *) (cigar + n_cigar); n_md = (int32_t) strlen(md) + 1; }
        int32_t hdr[4] = { score, n_cigar, nm, n_md };
        fwrite(hdr, 4, 4, fo);
        if (n_cigar) fwrite(cigar, 4, (size_t) n_cigar, fo);
        if (n_md) fwrite(md, 1, (size_t) n_md, fo);
        free(cigar);
    }
    fclose(fi); fclose(fo);
    free(opt); bns_destroy(bns);
    return 0;
}

static int ksw_mode(int argc, char *argv[]) {
    if (argc < 4) { fprintf(stderr, "usage: ref_driver ksw <requests.bin> <out.bin>\n"); return 1; }
    mem_opt_t *opt = mem_opt_init();
    FILE *fi = fopen(argv[2], "rb"), *fo = fopen(argv[3], "wb");
    if (!fi || !fo) return 1;
    int64_t n = 0;
    if (fread(&n, 8, 1, fi) != 1) return 1;
    std::vector<uint8_t> q, t;
    for (int64_t i = 0; i < n; ++i) {
        int32_t h[3];
        if (fread(h, 4, 3, fi) != 3) return 1;
        q.resize((size_t) h[0] + 1); t.resize((size_t) h[1] + 1);
        if ((h[0] && fread(q.data(), 1, (size_t) h[0], fi) != (size_t) h[0]) || (h[1] && fread(t.data(), 1, (size_t) h[1], fi) != (size_t) h[1])) return 1;
        kswr_t r = ksw_align2(h[0], q.data(), h[1], t.data(), 5, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, h[2], 0);
        int32_t o[7] = { r.score, r.te, r.qe, r.score2, r.te2, r.tb, r.qb };
        fwrite(o, 4, 7, fo);
    }
    fclose(fi); fclose(fo); free(opt);
    return 0;
}

extern int mem_matesw(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4], const mem_alnreg_t *a, int l_ms,
                      const uint8_t *ms, mem_alnreg_v *ma);

static int matesw_mode(int argc, char *argv[]) {
    if (argc < 5) { fprintf(stderr, "usage: ref_driver matesw <index prefix> <in.bin> <out.bin>\n"); return 1; }
    bntseq_t *bns = bns_restore(argv[2]);
    if (!bns) return 1;
    std::vector<uint8_t> pac((size_t) (bns->l_pac / 4 + 1));
    if (fread(pac.data(), 1, pac.size(), bns->fp_pac) == 0) return 1;
    mem_opt_t *opt = mem_opt_init();
    FILE *fi = fopen(argv[3], "rb"), *fo = fopen(argv[4], "wb");
    if (!fi || !fo) return 1;
    mem_pestat_t pes[4]; memset(pes, 0, sizeof(pes));
    for (int d = 0; d < 4; ++d) { int32_t v[3]; if (fread(v, 4, 3, fi) != 3) return 1; pes[d].low = v[0]; pes[d].high = v[1]; pes[d].failed = v[2]; }
    int64_t np = 0;
    if (fread(&np, 8, 1, fi) != 1) return 1;
    for (int64_t pr = 0; pr < np; ++pr) {
        std::vector<uint8_t> seq[2]; mem_alnreg_v a[2];
        for (int i = 0; i < 2; ++i) {
            int32_t l = 0, n = 0;
            if (fread(&l, 4, 1, fi) != 1) return 1;
            seq[i].resize((size_t) l + 1);
            if (l && fread(seq[i].data(), 1, (size_t) l, fi) != (size_t) l) return 1;
            seq[i].resize((size_t) l);
            if (fread(&n, 4, 1, fi) != 1) return 1;
            a[i].n = (size_t) n; a[i].m = (size_t) n + 8; a[i].a = (mem_alnreg_t *) calloc(a[i].m, sizeof(mem_alnreg_t));
            if (n && fread(a[i].a, sizeof(mem_alnreg_t), (size_t) n, fi) != (size_t) n) return 1;
        }
        /* the rescue block of mem_sam_pe (src/bwamem_pair.cpp:378-412 with MATE_SORT == 0) */
        mem_alnreg_v b[2]; kv_init(b[0]); kv_init(b[1]);
        for (int i = 0; i < 2; ++i)
            for (size_t j = 0; j < a[i].n; ++j)
                if (a[i].a[j].score >= a[i].a[0].score - opt->pen_unpaired) kv_push(mem_alnreg_t, b[i], a[i].a[j]);
        for (int i = 0; i < 2; ++i)
            for (size_t j = 0; j < b[i].n && (int) j < opt->max_matesw; ++j) {
                const int n = mem_matesw(opt, bns, pac.data(), pes, &b[i].a[j], (int) seq[!i].size(), seq[!i].data(), &a[!i]);
                int32_t h[5] = { (int32_t) pr, i, (int32_t) j, n, (int32_t) a[!i].n };
                fwrite(h, 4, 5, fo);
                if (a[!i].n) fwrite(a[!i].a, sizeof(mem_alnreg_t), a[!i].n, fo);
            }
        free(b[0].a); free(b[1].a); free(a[0].a); free(a[1].a);
    }
    fclose(fi); fclose(fo); free(opt); bns_destroy(bns);
    return 0;
}

int main(int argc, char *argv[]) {
    if (argc >= 2 && strcmp(argv[1], "matesw") == 0) return matesw_mode(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "cigar") == 0) return cigar_mode(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "ksw") == 0) return ksw_mode(argc, argv);
    const char *m = getenv("BM2_MODE");
    if (m) {
        if (!strcmp(m, "hotpath")) g_mode = M_HOTPATH;
        else if (!strcmp(m, "gpu")) g_mode = M_GPU;
        else if (!strcmp(m, "gpu_bsw")) g_mode = M_GPU_BSW;
    }
    if (getenv("BM2_STATS")) g_stats = getenv("BM2_STATS");
    if (getenv("BM2_DUMP_PREFIX")) {
        g_dump = getenv("BM2_DUMP_PREFIX");
        f_smem = fopen((g_dump + ".smem.bin").c_str(), "wb");
        f_chain = fopen((g_dump + ".chains.bin").c_str(), "wb");
        f_regs = fopen((g_dump + ".regs.bin").c_str(), "wb");
        f_bsw = fopen((g_dump + ".bsw.bin").c_str(), "wb");
        f_pestat = fopen((g_dump + ".pestat.bin").c_str(), "wb");
    }
    if (getenv("BM2_DUMP_REGS") && !f_regs) f_regs = fopen(getenv("BM2_DUMP_REGS"), "wb");
    if (getenv("BM2_REPEAT")) { g_repeat = atoi(getenv("BM2_REPEAT")); if (g_repeat < 1) g_repeat = 1; }
    /* rdtsc calibration as src/main.cpp:57-59, shortened */
    uint64_t tim = __rdtsc(); usleep(100000); proc_freq = (__rdtsc() - tim) * 10;
    if (argc < 2 || strcmp(argv[1], "mem") != 0) {
        fprintf(stderr, "usage: ref_driver mem <bwa-mem2 mem arguments>\n");
        return 1;
    }
    kstring_t pg = {0, 0, 0};
    ksprintf(&pg, "@PG\tID:bwa-mem2\tPN:bwa-mem2\tVN:2.2.1\tCL:%s", argv[0]);
    for (int i = 1; i < argc; ++i) ksprintf(&pg, " %s", argv[i]);
    ksprintf(&pg, "\n");
    bwa_pg = pg.s;
    tprof[MEM][0] = __rdtsc();
    int ret = main_mem(argc - 1, argv + 1);
    write_stats();
    if (g_ctx) p_destroy(g_ctx);
    for (FILE *f : { f_smem, f_chain, f_regs, f_bsw, f_pestat }) if (f) fclose(f);
    return ret;
}
