// bm2_common.cuh — shared declarations of the B200 seed-and-extend library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include "bm2_b200.h"

#define BM2_CUDA_OK(expr)                                                                       \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            bm2_set_error(ctx_for_error, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
            return 1;                                                                           \
        }                                                                                       \
    } while (0)

struct bm2_ctx;
void bm2_set_error(bm2_ctx *ctx, const std::string &msg);

// Scoring / extension parameters as the BSW kernels consume them (mem_opt_t subset).
struct BswParams {
    int a, b;                 // match, mismatch penalty (positive)
    int o_del, e_del, o_ins, e_ins;
    int zdrop;
    int end_bonus;
    int w;                    // band of this launch
};

// One extension job.  Sequences are addressed as base[off + k*stride], stride = +1 or -1, so that
// left extensions read the read and the reference backwards without materialising reversed copies
// (the reference materialises them: src/bwamem.cpp:2277, :2297).
struct BswJob {
    int64_t toff;             // offset of target[0] in the target buffer
    int64_t qoff;             // offset of query[0] in the query buffer
    int32_t tlen, qlen;
    int32_t h0;
    int8_t  tstride, qstride;
    int16_t _pad;
};

struct BswOut {               // as SeqPair's result fields (src/bandedSWA.h:96-97)
    int32_t score, tle, gtle, qle, gscore, max_off;
};

// Sort-and-launch of the thread-per-job BSW kernel over `n` jobs (device arrays).
// perm/keys are scratch of n elements each; cells (device, may be null) accumulates DP cells.
int bsw_launch(bm2_ctx *ctx_for_error, cudaStream_t stream, const BswJob *d_jobs, BswOut *d_out, int n,
               const uint8_t *d_tbase, const uint8_t *d_qbase, const BswParams &prm,
               unsigned long long *d_cells);
size_t bsw_scratch_bytes(int n);
