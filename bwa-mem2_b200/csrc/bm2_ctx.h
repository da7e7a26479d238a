// bm2_ctx.h — the device context behind the opaque `bm2_ctx*` of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include <mutex>
#include "bm2_b200.h"

struct DevBuf  { void *p = nullptr; size_t cap = 0; };
struct HostBuf { void *p = nullptr; size_t cap = 0; };

struct DevIndex {                 // FM-index + reference resident in HBM (replicated per GPU)
    int64_t N = 0, l_pac = 0, sentinel = 0;
    int64_t count[5] = {0, 0, 0, 0, 0};
    const bm2_cp_occ *cp_occ = nullptr;
    int occ_layout = 0;                    // FmIndexView::layout of the resident table (1 = half-checkpoint sectors, made at upload)
    const int8_t *sa_ms = nullptr;
    const uint32_t *sa_ls = nullptr;
    const uint8_t *ref = nullptr;          // 2*l_pac codes
    int32_t n_seqs = 0;
    const int64_t *ann_off = nullptr;
    const int32_t *ann_len = nullptr;
    const int32_t *ann_alt = nullptr;
    bool loaded = false;
};

struct bm2_ctx {
    int device = 0, n_sm = 148;
    cudaStream_t stream = nullptr, own_stream = nullptr, side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bm2_mem_opt_t opt;
    std::string err;
    DevIndex idx;
    std::vector<void *> idx_allocs;
    // seam 1
    DevBuf io_pairs, io_ref, io_qer, bsw_jobs, bsw_outs, bsw_scratch;
    // seam 2 (pipeline.cu)
    DevBuf d[128];         // slots: pipeline.cu 0-47, cigar.cu 48-63, sam.cu 64-83 + 90-95, ksw.cu 84-89, fastq.cu 100-111
    HostBuf h[48];         // slots: pipeline.cu 0-7, cigar.cu 8-15, sam.cu 16-23, fastq.cu 24-31
    std::vector<cudaEvent_t> events;
    std::vector<const char *> stage_names;
    std::vector<float> stage_ms;
    unsigned long long last_n_ext = 0, last_n_lf = 0, last_cells = 0, last_n_retry[2] = {0, 0};
    // seam 2 sub-batches in flight (pipeline.cu run_regs): child contexts with their own streams, events and scratch;
    // they share this context's index (their idx_allocs stay empty)
    int n_lanes = 4, lane_min_reads = 16384;
    std::vector<bm2_ctx *> lanes;
    bm2_ctx *parent = nullptr;                 // set in a lane
    // stage tokens (held by one lane at a time): the sub-batches take turns in the DRAM-bound SMEM stage and in the
    // ALU-bound extension stage, so that at any time DIFFERENT kinds of stages overlap instead of four copies of the same
    std::mutex tok_smem, tok_bsw;
    cudaEvent_t ev_entry = nullptr;
    // seam 4 (sam.cu): staged rescue switch (-1: the BM2_SAM_STAGED environment variable decides, default off), events around the stage's
    // kernels, and the last call's times (jobs, window alignments, pairs, gather; ms summed over waves) and counters
    int sam_staged = -1;
    cudaEvent_t sam_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double sam_ms[4] = {0, 0, 0, 0};
    unsigned long long sam_counts[6] = {0, 0, 0, 0, 0, 0};        // staged, jobs, looked up, computed in place, of those: window moved, waves

    int ensure(DevBuf &b, size_t bytes);
    int ensure_host(HostBuf &b, size_t bytes);
    std::vector<DevBuf *> all_dev() {
        std::vector<DevBuf *> v = {&io_pairs, &io_ref, &io_qer, &bsw_jobs, &bsw_outs, &bsw_scratch};
        for (auto &x : d) v.push_back(&x);
        return v;
    }
    std::vector<HostBuf *> all_host() { std::vector<HostBuf *> v; for (auto &x : h) v.push_back(&x); return v; }
};
