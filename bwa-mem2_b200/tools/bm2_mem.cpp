// bm2_mem — FASTQ in, SAM out, on the GPU path: the host side of `bwa-mem2 mem` for the seams of libbm2b200.so (C++, as the reference's host
// code; only the C ABI of include/bm2_b200.h is used).
//
//   bm2_mem [-t threads] [-K chunk_bases] [-p workers] [-o out.sam] <index prefix> <reads_1.fq> [reads_2.fq]
//
// What the reference does in main_mem / process / ktp_worker (src/fastmap.cpp:616-1003, :280-349), with every step of a chunk behind a seam:
//   chunk of the input            bseq_read_orig's rule (src/bwa.cpp:170-216): records until the base count reaches the task size
//                                 (-K, else chunk_size x threads, src/fastmap.cpp:943-949), mates kept together
//   bm2_fastq_encode              parsing + nst_nt4_table encoding on the GPU                         (kseq + src/bwamem.cpp:992-1000)
//   bm2_seed_chain_extend_resident  worker_bwt + worker_aln                                           (src/bwamem.cpp:1359-1363)
//   bm2_pestat                    mem_pestat                                                          (src/bwamem.cpp:1368-1378)
//   bm2_sam_pe / bm2_sam_se       worker_sam's arithmetic                                             (src/bwamem.cpp:1262-1336)
//   bm2_sam_format                mem_aln2sam's text                                                  (src/bwamem.cpp:1592-1730)
// Chunks in flight: the reference's kt_pipeline runs its three steps (read, process, write) on two worker threads so that one chunk's I/O
// overlaps another's computation (src/fastmap.cpp:952-1003, src/kthread.cpp:122-176).  Here -p workers (default 2) each own a context
// (bm2_create_sibling: one index in HBM) and take whole chunks off a queue; the GPU interleaves the kernels of the two chunks, the host side of
// one (pestat, formatting, fwrite) runs under the GPU stages of the other, and the output is written strictly in chunk order.
// The SAM records equal `bwa-mem2 mem` with the same -K (tests/test_zz_fastq_sam_gpu.py); the header carries the same @SQ lines and
// this program's own @PG line.  FASTQ with four-line records, plain or gzip; the files are read whole.
#include "bm2_b200.h"
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <zlib.h>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// the whole file into memory; gzip files (magic 1f 8b) through zlib, as the reference reads its input through zlib (gzdopen + kseq, src/fastmap.cpp:
// 905-907, :933-935) - gzread also passes plain files through, but the plain path below needs no copy loop
static bool read_file(const char *path, std::vector<char> &buf) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    unsigned char magic[2] = {0, 0};
    const size_t got = fread(magic, 1, 2, f);
    if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        fclose(f);
        gzFile g = gzopen(path, "rb");
        if (!g) return false;
        gzbuffer(g, 1 << 20);
        size_t n = 0;
        buf.resize((size_t) 64 << 20);
        for (;;) {
            if (buf.size() - n < ((size_t) 16 << 20)) buf.resize(buf.size() * 2);
            const size_t want = buf.size() - n < ((size_t) 1 << 30) ? buf.size() - n : ((size_t) 1 << 30);
            const int r = gzread(g, buf.data() + n, (unsigned) want);
            if (r < 0) { gzclose(g); return false; }
            if (r == 0) break;
            n += (size_t) r;
        }
        gzclose(g);
        buf.resize(n);
        return true;
    }
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    buf.resize((size_t) n);
    const bool ok = n == 0 || fread(buf.data(), 1, (size_t) n, f) == (size_t) n;
    fclose(f);
    return ok;
}

// end of the record that starts at p (four lines), and the length of its sequence line; nullptr at a truncated record
static const char *next_record(const char *p, const char *end, int64_t *seq_len) {
    const char *l[4]; const char *q = p;
    for (int k = 0; k < 4; ++k) {
        const char *e = (const char *) memchr(q, '\n', (size_t) (end - q));
        if (!e) { if (k == 3 && q < end) { e = end; l[k] = e; q = end; break; } return nullptr; }
        l[k] = e; q = e + 1;
    }
    const char *s0 = l[0] + 1;
    int64_t n = l[1] - s0;
    if (n > 0 && s0[n - 1] == '\r') --n;
    *seq_len = n;
    return q;
}

namespace {

struct Chunk { long long index, first_read; const char *c1, *c2; size_t n1, n2; };

struct Shared {
    // work queue (the chunker fills it, bounded), write order, totals
    std::mutex mu; std::condition_variable cv_work, cv_room, cv_turn;
    std::deque<Chunk> queue; bool done = false;
    long long next_to_write = 0;
    double t_loop = 0;
    double t_enc = 0, t_aln = 0, t_pes = 0, t_sam = 0, t_fmt = 0, t_write = 0, t_turn = 0;
    std::vector<double> chunk_s, chunk_done_s; std::vector<long long> chunk_reads;
    long long n_processed = 0;
    // constants
    const bm2_mem_opt_t *opt = nullptr; const bm2_index_desc *idx = nullptr; const char *const *cnames = nullptr;
    bool paired = false; int threads = 1; FILE *out = nullptr;
};

[[noreturn]] void die(const char *what, const bm2_ctx *ctx) { fprintf(stderr, "bm2_mem: %s%s%s\n", what, ctx ? ": " : "", ctx ? bm2_last_error(ctx) : ""); fflush(stderr); _Exit(3); }

void worker(Shared *sh, bm2_ctx *ctx) {
    for (;;) {
        Chunk ck;
        {
            std::unique_lock<std::mutex> lk(sh->mu);
            sh->cv_work.wait(lk, [&] { return !sh->queue.empty() || sh->done; });
            if (sh->queue.empty()) return;
            ck = sh->queue.front(); sh->queue.pop_front();
            sh->cv_room.notify_one();
        }
        const double t0 = now_s();
        bm2_fastq_batch fq;
        if (bm2_fastq_encode(ctx, ck.c1, (int64_t) ck.n1, sh->paired ? ck.c2 : nullptr, sh->paired ? (int64_t) ck.n2 : 0, &fq)) die("bm2_fastq_encode", ctx);
        const double t1 = now_s();
        bm2_read_batch rb = { fq.n_reads, fq.codes, fq.offsets };
        bm2_reg_result rr;
        if (bm2_seed_chain_extend_resident(ctx, &rb, fq.d_codes, fq.d_offsets, 1, &rr)) die("bm2_seed_chain_extend_resident", ctx);
        const double t2 = now_s();
        double t3 = t2;
        bm2_sam_result sr;
        if (sh->paired) {
            bm2_pestat_t pes[4];
            if (bm2_pestat(sh->opt, sh->idx->l_pac, fq.n_reads, rr.regs, rr.read_off, pes)) die("bm2_pestat failed", nullptr);
            t3 = now_s();
            if (bm2_sam_pe(ctx, &rb, rr.regs, rr.read_off, pes, ck.first_read >> 1, &sr)) die("bm2_sam_pe", ctx);
        } else if (bm2_sam_se(ctx, &rb, rr.regs, rr.read_off, ck.first_read, &sr)) die("bm2_sam_se", ctx);
        const double t4 = now_s();
        bm2_sam_text_in tin; memset(&tin, 0, sizeof tin);
        tin.res = &sr; tin.reads = &rb; tin.quals = fq.quals; tin.contig_names = sh->cnames;
        tin.name_buf[0] = ck.c1; tin.name_buf[1] = sh->paired ? ck.c2 : nullptr; tin.name_beg = fq.name_beg; tin.name_len = fq.name_len;
        char *text = nullptr; int64_t len = 0;
        if (bm2_sam_format(&tin, sh->threads, &text, &len)) die("bm2_sam_format failed", nullptr);
        const double t5 = now_s();
        {   // the output keeps the chunk order
            std::unique_lock<std::mutex> lk(sh->mu);
            sh->cv_turn.wait(lk, [&] { return sh->next_to_write == ck.index; });
        }
        const double t6 = now_s();
        fwrite(text, 1, (size_t) len, sh->out);
        bm2_free(text);
        const double t7 = now_s();
        {
            std::lock_guard<std::mutex> lk(sh->mu);
            sh->t_enc += t1 - t0; sh->t_aln += t2 - t1; sh->t_pes += t3 - t2; sh->t_sam += t4 - t3; sh->t_fmt += t5 - t4; sh->t_turn += t6 - t5; sh->t_write += t7 - t6;
            sh->n_processed += fq.n_reads;
            sh->chunk_s.push_back(t7 - t0); sh->chunk_done_s.push_back(t7 - sh->t_loop); sh->chunk_reads.push_back(fq.n_reads);
            ++sh->next_to_write;
        }
        sh->cv_turn.notify_all();
    }
}

}  // namespace

int main(int argc, char **argv) {
    int threads = 1, workers = 2; long long fixed_k = 0; const char *out_path = nullptr;
    int a = 1;
    for (; a < argc && argv[a][0] == '-' && argv[a][1]; a += 2) {
        if (a + 1 >= argc) break;
        if (!strcmp(argv[a], "-t")) threads = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-K")) fixed_k = atoll(argv[a + 1]);
        else if (!strcmp(argv[a], "-p")) workers = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-o")) out_path = argv[a + 1];
        else { fprintf(stderr, "bm2_mem: unknown option %s\n", argv[a]); return 1; }
    }
    if (argc - a < 2) { fprintf(stderr, "usage: bm2_mem [-t threads] [-K chunk_bases] [-p workers] [-o out.sam] <index prefix> <reads_1.fq> [reads_2.fq]\n"); return 1; }
    if (threads < 1) threads = 1;
    if (workers < 1) workers = 1;
    if (workers > 4) workers = 4;
    const char *prefix = argv[a], *f1 = argv[a + 1], *f2 = argc - a >= 3 ? argv[a + 2] : nullptr;
    const double t_start = now_s();
    bm2_index_desc *idx = nullptr;
    if (bm2_index_load(prefix, &idx)) { fprintf(stderr, "bm2_mem: cannot load the index %s\n", prefix); return 2; }
    // contig names: <prefix>.ann (src/bntseq.cpp:73-110): "l_pac n_seqs seed", then per contig "gi name [anno]" and "offset len n_ambs"
    std::vector<std::string> names; std::vector<long long> lens;
    {
        FILE *f = fopen((std::string(prefix) + ".ann").c_str(), "r");
        if (!f) { fprintf(stderr, "bm2_mem: cannot open %s.ann\n", prefix); return 2; }
        char line[65536];
        if (!fgets(line, sizeof line, f)) return 2;
        for (int i = 0; i < idx->n_seqs; ++i) {
            char nm[4096]; long long gi, off, len; int amb;
            if (!fgets(line, sizeof line, f) || sscanf(line, "%lld %4095s", &gi, nm) != 2) return 2;
            if (!fgets(line, sizeof line, f) || sscanf(line, "%lld %lld %d", &off, &len, &amb) != 3) return 2;
            names.push_back(nm); lens.push_back(len);
        }
        fclose(f);
    }
    std::vector<const char *> cnames; for (auto &s : names) cnames.push_back(s.c_str());
    bm2_mem_opt_t opt; bm2_opt_init(&opt);
    opt.n_threads = threads;
    if (f2) opt.flag |= 0x2;                                   // MEM_F_PE
    std::vector<bm2_ctx *> ctxs((size_t) workers, nullptr);
    if (bm2_create(&ctxs[0], 0, idx, &opt)) { fprintf(stderr, "bm2_mem: %s\n", bm2_last_error(nullptr)); return 3; }
    for (int w = 1; w < workers; ++w)
        if (bm2_create_sibling(&ctxs[w], ctxs[0])) { fprintf(stderr, "bm2_mem: %s\n", bm2_last_error(ctxs[0])); return 3; }
    const double t_index = now_s() - t_start;
    std::vector<char> b1, b2;
    if (!read_file(f1, b1) || (f2 && !read_file(f2, b2))) { fprintf(stderr, "bm2_mem: cannot read the FASTQ files\n"); return 2; }
    FILE *out = out_path ? fopen(out_path, "wb") : stdout;
    if (!out) { fprintf(stderr, "bm2_mem: cannot open %s\n", out_path); return 2; }
    for (size_t i = 0; i < names.size(); ++i) fprintf(out, "@SQ\tSN:%s\tLN:%lld\n", names[i].c_str(), lens[i]);
    fprintf(out, "@PG\tID:bm2_mem\tPN:bm2_mem\tVN:b200-r2\tCL:%s", argv[0]);
    for (int i = 1; i < argc; ++i) fprintf(out, " %s", argv[i]);
    fprintf(out, "\n");
    const long long task = fixed_k > 0 ? fixed_k : (long long) opt.chunk_size * threads;
    Shared sh;
    sh.opt = &opt; sh.idx = idx; sh.cnames = cnames.data(); sh.paired = f2 != nullptr; sh.threads = threads; sh.out = out;
    sh.t_loop = now_s();
    std::vector<std::thread> pool;
    for (int w = 0; w < workers; ++w) pool.emplace_back(worker, &sh, ctxs[w]);
    // the chunker: records until the base count reaches the task size (src/bwa.cpp:204), mates kept together
    const char *p1 = b1.data(), *e1 = p1 + b1.size(), *p2 = b2.data(), *e2 = p2 + b2.size();
    long long n_chunks = 0, first_read = 0;
    while (p1 < e1) {
        const char *c1 = p1, *c2 = p2; long long size = 0, n_rec = 0;
        while (p1 < e1) {
            int64_t sl = 0;
            const char *q = next_record(p1, e1, &sl);
            if (!q) die("truncated record in the 1st file", nullptr);
            p1 = q; size += sl; ++n_rec;
            if (f2) {
                if (p2 >= e2) die("the 2nd file has fewer sequences", nullptr);
                const char *r = next_record(p2, e2, &sl);
                if (!r) die("truncated record in the 2nd file", nullptr);
                p2 = r; size += sl; ++n_rec;
            }
            if (size >= task) break;
        }
        Chunk ck = { n_chunks++, first_read, c1, c2, (size_t) (p1 - c1), (size_t) (p2 - c2) };
        first_read += n_rec;
        std::unique_lock<std::mutex> lk(sh.mu);
        sh.cv_room.wait(lk, [&] { return (int) sh.queue.size() < workers; });
        sh.queue.push_back(ck);
        sh.cv_work.notify_one();
    }
    { std::lock_guard<std::mutex> lk(sh.mu); sh.done = true; }
    sh.cv_work.notify_all();
    for (auto &t : pool) t.join();
    if (f2 && p2 < e2) fprintf(stderr, "[W::bm2_mem] the 1st file has fewer sequences.\n");
    const double loop_s = now_s() - sh.t_loop;
    if (out != stdout) fclose(out);
    fprintf(stderr, "{\"reads\": %lld, \"chunks\": %lld, \"workers\": %d, \"loop_s\": %.6f, \"index_and_context_s\": %.3f, \"fastq_encode_s\": %.6f, \"seed_chain_extend_s\": %.6f, "
                    "\"pestat_s\": %.6f, \"sam_stage_s\": %.6f, \"sam_format_s\": %.6f, \"wait_for_turn_s\": %.6f, \"write_s\": %.6f, \"chunk_s\": [",
            sh.n_processed, n_chunks, workers, loop_s, t_index, sh.t_enc, sh.t_aln, sh.t_pes, sh.t_sam, sh.t_fmt, sh.t_turn, sh.t_write);
    for (size_t i = 0; i < sh.chunk_s.size(); ++i) fprintf(stderr, "%s%.6f", i ? ", " : "", sh.chunk_s[i]);
    fprintf(stderr, "], \"chunk_done_s\": [");
    for (size_t i = 0; i < sh.chunk_done_s.size(); ++i) fprintf(stderr, "%s%.6f", i ? ", " : "", sh.chunk_done_s[i]);
    fprintf(stderr, "], \"chunk_reads\": [");
    for (size_t i = 0; i < sh.chunk_reads.size(); ++i) fprintf(stderr, "%s%lld", i ? ", " : "", sh.chunk_reads[i]);
    fprintf(stderr, "]}\n");
    for (int w = workers - 1; w >= 0; --w) bm2_destroy(ctxs[w]);
    bm2_index_free(idx);
    return 0;
}
