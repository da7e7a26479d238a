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


// How the reference's SIMD class of a job bends its arithmetic (results must equal the reference's for every mem_opt_t):
// jobs that sortPairsLenExt (src/bwamem.cpp:1944-1952) sends to the 8-bit kernel - len1, len2 < 128 and
// h0 + min(len1, len2) * a < 128 - keep the band operands and the z-drop threshold in 8 bits (src/bandedSWA.cpp:2195-2216,
// :2347), the 16-bit kernel in 16 bits (:2905-2926, :3044); neither kernel guards the z-drop test with `zdrop > 0`
// (ZSCORE8/16 run on every row, :1826-1839, :1868-1880), so -d 0 drops at the first non-improving row and a threshold that
// went negative (-d 128..255 in the 8-bit class) ends the job at its first row.
struct BswQuirk { unsigned band_mask; int zthr; };
#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
BswQuirk bsw_quirk(int qlen, int tlen, int h0, const BswParams &p) {
    const int minlen = qlen < tlen ? qlen : tlen;
    const bool k8 = tlen < 128 && qlen < 128 && h0 + minlen * p.a < 128;
    BswQuirk q;
    q.band_mask = k8 ? 0xFFu : 0xFFFFu;
    q.zthr = k8 ? (int) (int8_t) p.zdrop : (int) (int16_t) p.zdrop;
    return q;
}
