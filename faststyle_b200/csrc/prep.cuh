// Table-driven weight / weight-gradient layout transforms (prep.cu): jobs grouped into dependency phases,
// one kernel launch per phase.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace fs {

enum PrepKind {
    PJ_PAD = 1, PJ_UNPAD, PJ_TRANSPOSE, PJ_FLIP_TRANSPOSE, PJ_UPCONV_COLLAPSE, PJ_UPCONV_COLLAPSE_GRAD,
    PJ_S2_DGRAD_COLLAPSE, PJ_S2_FWD_COLLAPSE, PJ_S2_FWD_COLLAPSE_GRAD, PJ_PAIR, PJ_UNPAIR, PJ_PACK_TAPS,
    PJ_PACK_W3X3, PJ_IN15, PJ_COPY, PJ_X16
};

struct PrepJob {
    int kind;
    int a, b, c, d, e;             // kind-specific dimensions / modes (see run_job in prep.cu)
    const float* src; const float* src2;
    float* dst;
    __nv_bfloat16* hi; __nv_bfloat16* lo;      // split-bf16 destinations of the packing jobs
    long long total;               // elements written (filled by PrepPlan::add)
};

constexpr int PREP_MAX_JOBS = 64;
constexpr int PREP_MAX_PHASES = 8;
struct PrepTable { PrepJob j[PREP_MAX_JOBS]; };      // 64 x 72 B = 4.6 KB: passed by value as a kernel parameter

struct PrepPlan {
    PrepTable t;
    int phase_of[PREP_MAX_JOBS];
    int n = 0;
    void clear() { n = 0; }
    int add(int phase, const PrepJob& job);           // jobs must be added in non-decreasing phase order
    int run(cudaStream_t st) const;                   // one launch per phase
};

long long prep_job_total(const PrepJob& j);

// convenience constructors
static inline PrepJob pj(int kind, const float* src, float* dst, int a = 0, int b = 0, int c = 0, int d = 0, int e = 0) {
    PrepJob j;
    memset(&j, 0, sizeof(j));
    j.kind = kind; j.src = src; j.dst = dst; j.a = a; j.b = b; j.c = c; j.d = d; j.e = e;
    return j;
}
static inline PrepJob pj_pack(int kind, const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, int a, int b, int c = 0, int d = 0) {
    PrepJob j = pj(kind, src, nullptr, a, b, c, d);
    j.hi = hi; j.lo = lo;
    return j;
}

}  // namespace fs
