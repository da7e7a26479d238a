// bsw_types.h — plain structs of the BSW kernels (no CUDA dependency: also included by the test-only host emulation).
#pragma once
#include <stdint.h>

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

