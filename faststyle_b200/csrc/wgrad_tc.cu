// tcgen05 pixel-reduction GEMMs: weight gradient of the 3x3 64->64 residual convolutions.
//
//   dW[kh,kw,ci,co] = sum_{n,y,x} x[n, y+kh-pad, x+kw-pad, ci] * dy[n,y,x,co]
//
// GEMM view: D[M = (tap pair) x 64 ci, N = 64 co] += A[K = pixels, M]^T * B[K = pixels, N]; both operands
// are "MN-major" (the reduction index - pixels - is the slow one: one 128-byte row of 64 channels per
// pixel), which is exactly what a TMA box of the NHWC split-bf16 planes produces.  Per 8x16-pixel dy tile
// and kw the kernel loads ONE x slab {64 ch, 16 px, 10 rows}; the taps kh=0,1,2 of that kw are the slab
// at +kh*2048 B.  Two M=128 accumulators per kw cover them: group 0 = (kh0 | kh1), group 1 = (kh1 | kh2)
// (the second 64-row half of an MN-major operand sits LBO = 2048 B after the first), so all MMAs are the
// proven M=128 shape; kh1 is computed twice and the duplicate dropped (the whole job is ~8 us per layer).
// Each persistent CTA accumulates its tiles in TMEM (6 accumulators x 64 columns), dumps one partial
// [9,64,64] fp32 and a fixed-order reduction sums the CTAs' partials (deterministic).
// Precision: split-bf16 x3 (hi*hi + hi*lo + lo*hi), fp32 accumulation - as conv3x3_tc.cu.
#include <cuda.h>
#include <stdlib.h>
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace fs {

namespace {

using namespace tcptx;

constexpr int TW = 16, TH = 8;
constexpr int SLAB_BYTES = (TH + 2) * TW * 128;      // one plane of an x slab
constexpr int DT_BYTES = TH * TW * 128;              // one plane of a dy tile
constexpr int X_STAGE = 2 * SLAB_BYTES, D_STAGE = 2 * DT_BYTES;
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * (X_STAGE + D_STAGE) + 1024 + 256;
constexpr int TMEM_COLS = 512;                       // 6 accumulators x 64 columns (power of two >= 384)

// MN-major, 128B-swizzled operand: one 128-byte row (64 channels) per reduction index (pixel);
// 8-row swizzle atoms every SBO = 1024 B; the next 64-channel group of M starts LBO bytes later.
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ uint32_t make_idesc_mn(int n = 64) {
    return (1u << 4) | (1u << 7) | (1u << 10)
         | (1u << 15) | (1u << 16)                  // A and B are MN-major
         | ((uint32_t)(n >> 3) << 17)               // N (multiple of 16, <= 256)
         | ((uint32_t)(128 >> 4) << 24);            // M = 128
}

struct WgParams {
    int N, OH, OW, pad, tilesX, tilesY;
    int total_tiles;
    float* partial;            // [gridDim.x][9*64*64]
};

__global__ void __launch_bounds__(256, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                   const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo,
                   const WgParams p) {
    FS_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smemX = smem;
    uint8_t* smemD = smem + STAGES * X_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemD + STAGES * D_STAGE);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + STAGES;
    uint64_t* d_full = x_empty + STAGES;
    uint64_t* d_empty = d_full + STAGES;
    uint64_t* done = d_empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX_hi); prefetch_tmap(&tmX_lo); prefetch_tmap(&tmD_hi); prefetch_tmap(&tmD_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FS_PDL_WAIT();                        // everything above is CTA-local setup
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int sx = 0, sd = 0; uint32_t px = 0, pd = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int tx = t % p.tilesX;
            int r = t / p.tilesX;
            const int ty = r % p.tilesY;
            const int n = r / p.tilesY;
            const int y0 = ty * TH, x0 = tx * TW;
            mbar_wait(&d_empty[sd], pd ^ 1);
            uint8_t* dd = smemD + sd * D_STAGE;
            mbar_expect_tx(&d_full[sd], D_STAGE);
            tma_load_4d(dd, &tmD_hi, &d_full[sd], 0, x0, y0, n);
            tma_load_4d(dd + DT_BYTES, &tmD_lo, &d_full[sd], 0, x0, y0, n);
            if (++sd == STAGES) { sd = 0; pd ^= 1; }
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&x_empty[sx], px ^ 1);
                uint8_t* xd = smemX + sx * X_STAGE;
                mbar_expect_tx(&x_full[sx], X_STAGE);
                tma_load_4d(xd, &tmX_hi, &x_full[sx], 0, x0 + kw - p.pad, y0 - p.pad, n);
                tma_load_4d(xd + SLAB_BYTES, &tmX_lo, &x_full[sx], 0, x0 + kw - p.pad, y0 - p.pad, n);
                if (++sx == STAGES) { sx = 0; px ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
        const uint32_t idesc = make_idesc_mn();
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int sx = 0, sd = 0; uint32_t px = 0, pd = 0;
        bool first_tile = true;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            mbar_wait(&d_full[sd], pd);
            tc_fence_after();
            const uint64_t dd_hi = make_sdesc_mn(smem_u32(smemD + sd * D_STAGE), 2048);
            const uint64_t dd_lo = make_sdesc_mn(smem_u32(smemD + sd * D_STAGE + DT_BYTES), 2048);
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&x_full[sx], px);
                tc_fence_after();
                const uint64_t xd_hi = make_sdesc_mn(smem_u32(smemX + sx * X_STAGE), 2048);
                const uint64_t xd_lo = make_sdesc_mn(smem_u32(smemX + sx * X_STAGE + SLAB_BYTES), 2048);
                if (elect_one()) {
#pragma unroll
                    for (int grp = 0; grp < 2; ++grp) {
                        const uint32_t acc = tb + (uint32_t)((kw * 2 + grp) * 64);
#pragma unroll
                        for (int prod = 0; prod < 3; ++prod) {
                            const uint64_t ad = (prod == 2 ? xd_lo : xd_hi) + (uint64_t)(grp * 128);   // +2048 B
                            const uint64_t bd = (prod == 1 ? dd_lo : dd_hi);
#pragma unroll
                            for (int ks = 0; ks < TH; ++ks)
                                tc_mma_bf16(acc, ad + (uint64_t)(ks * 128), bd + (uint64_t)(ks * 128), idesc,
                                            (first_tile && prod == 0 && ks == 0) ? 0u : 1u);
                        }
                    }
                    tc_commit(&x_empty[sx]);
                }
                __syncwarp();
                if (++sx == STAGES) { sx = 0; px ^= 1; }
            }
            if (elect_one()) tc_commit(&d_empty[sd]);
            __syncwarp();
            if (++sd == STAGES) { sd = 0; pd ^= 1; }
            first_tile = false;
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== dump (4 warps = 128 TMEM lanes) =====================
        const int ew = warp - 4;
        const int row = ew * 32 + lane;              // accumulator row: 0..63 first tap, 64..127 second tap
        mbar_wait(done, 0);
        tc_fence_after();
        float* part = p.partial + (long long)blockIdx.x * (9 * 64 * 64);
        const bool has_tiles = (int)blockIdx.x < p.total_tiles;
#pragma unroll 1
        for (int a = 0; a < 6; ++a) {
            const int kw = a >> 1, grp = a & 1;
            const int kh = grp + (row >> 6);          // group 0: kh 0|1, group 1: kh 1|2
            const bool keep = !(grp == 1 && row < 64);   // group 1's first half duplicates kh=1
            const int tap = kh * 3 + kw, ci = row & 63;
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * 64 + ch * 32), v);
                if (keep) {
                    float* op = part + ((long long)tap * 64 + ci) * 64 + ch * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(op + i) = has_tiles ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3])
                                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// Batched form: the weight gradients of SEVERAL 3x3 64->64 convolutions (the ten residual convs of the transform
// net) in ONE launch.  Each problem owns a contiguous range of CTAs; a CTA accumulates all of its tiles in TMEM and
// dumps one partial [9,64,64], so a problem produces ~15 partials instead of 148 and the ten launches + ten
// reductions of the per-layer form (about 7 us of MMA work each against ~25 us of dump + reduce + launch latency)
// become one launch + one reduction.  The body is wgrad3x3_tc_kernel's.
struct WgProblem {
    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    int pad, tilesX, tilesY, total_tiles;
    int cta_begin, cta_count;
    float* partial;            // [cta_count][9*64*64]
    float* out;                // [3,3,64,64]
};
struct WgMulti { WgProblem pr[WGRAD_MULTI_MAX]; int count; };

__global__ void __launch_bounds__(256, 1)
wgrad3x3_tc_multi_kernel(const __grid_constant__ WgMulti M) {
    FS_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smemX = smem;
    uint8_t* smemD = smem + STAGES * X_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemD + STAGES * D_STAGE);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + STAGES;
    uint64_t* d_full = x_empty + STAGES;
    uint64_t* d_empty = d_full + STAGES;
    uint64_t* done = d_empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    int j = 0;
    while (j + 1 < M.count && (int)blockIdx.x >= M.pr[j + 1].cta_begin) ++j;
    const WgProblem& P = M.pr[j];
    const int cta = (int)blockIdx.x - P.cta_begin, nctas = P.cta_count;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&P.tmX_hi); prefetch_tmap(&P.tmX_lo); prefetch_tmap(&P.tmD_hi); prefetch_tmap(&P.tmD_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FS_PDL_WAIT();                        // everything above is CTA-local setup
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        int sx = 0, sd = 0; uint32_t px = 0, pd = 0;
        for (int t = cta; t < P.total_tiles; t += nctas) {
            const int tx = t % P.tilesX;
            int r = t / P.tilesX;
            const int ty = r % P.tilesY;
            const int n = r / P.tilesY;
            const int y0 = ty * TH, x0 = tx * TW;
            mbar_wait(&d_empty[sd], pd ^ 1);
            uint8_t* dd = smemD + sd * D_STAGE;
            mbar_expect_tx(&d_full[sd], D_STAGE);
            tma_load_4d(dd, &P.tmD_hi, &d_full[sd], 0, x0, y0, n);
            tma_load_4d(dd + DT_BYTES, &P.tmD_lo, &d_full[sd], 0, x0, y0, n);
            if (++sd == STAGES) { sd = 0; pd ^= 1; }
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&x_empty[sx], px ^ 1);
                uint8_t* xd = smemX + sx * X_STAGE;
                mbar_expect_tx(&x_full[sx], X_STAGE);
                tma_load_4d(xd, &P.tmX_hi, &x_full[sx], 0, x0 + kw - P.pad, y0 - P.pad, n);
                tma_load_4d(xd + SLAB_BYTES, &P.tmX_lo, &x_full[sx], 0, x0 + kw - P.pad, y0 - P.pad, n);
                if (++sx == STAGES) { sx = 0; px ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_mn();
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int sx = 0, sd = 0; uint32_t px = 0, pd = 0;
        bool first_tile = true;
        for (int t = cta; t < P.total_tiles; t += nctas) {
            mbar_wait(&d_full[sd], pd);
            tc_fence_after();
            const uint64_t dd_hi = make_sdesc_mn(smem_u32(smemD + sd * D_STAGE), 2048);
            const uint64_t dd_lo = make_sdesc_mn(smem_u32(smemD + sd * D_STAGE + DT_BYTES), 2048);
            for (int kw = 0; kw < 3; ++kw) {
                mbar_wait(&x_full[sx], px);
                tc_fence_after();
                const uint64_t xd_hi = make_sdesc_mn(smem_u32(smemX + sx * X_STAGE), 2048);
                const uint64_t xd_lo = make_sdesc_mn(smem_u32(smemX + sx * X_STAGE + SLAB_BYTES), 2048);
                if (elect_one()) {
#pragma unroll
                    for (int grp = 0; grp < 2; ++grp) {
                        const uint32_t acc = tb + (uint32_t)((kw * 2 + grp) * 64);
#pragma unroll
                        for (int prod = 0; prod < 3; ++prod) {
                            const uint64_t ad = (prod == 2 ? xd_lo : xd_hi) + (uint64_t)(grp * 128);   // +2048 B
                            const uint64_t bd = (prod == 1 ? dd_lo : dd_hi);
#pragma unroll
                            for (int ks = 0; ks < TH; ++ks)
                                tc_mma_bf16(acc, ad + (uint64_t)(ks * 128), bd + (uint64_t)(ks * 128), idesc,
                                            (first_tile && prod == 0 && ks == 0) ? 0u : 1u);
                        }
                    }
                    tc_commit(&x_empty[sx]);
                }
                __syncwarp();
                if (++sx == STAGES) { sx = 0; px ^= 1; }
            }
            if (elect_one()) tc_commit(&d_empty[sd]);
            __syncwarp();
            if (++sd == STAGES) { sd = 0; pd ^= 1; }
            first_tile = false;
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        float* part = P.partial + (long long)cta * (9 * 64 * 64);
        const bool has_tiles = cta < P.total_tiles;
#pragma unroll 1
        for (int a = 0; a < 6; ++a) {
            const int kw = a >> 1, grp = a & 1;
            const int kh = grp + (row >> 6);
            const bool keep = !(grp == 1 && row < 64);
            const int tap = kh * 3 + kw, ci = row & 63;
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * 64 + ch * 32), v);
                if (keep) {
                    float* op = part + ((long long)tap * 64 + ci) * 64 + ch * 32;
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(op + i) = has_tiles ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3])
                                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

struct WgReduceMulti { const float* partial[WGRAD_MULTI_MAX]; float* out[WGRAD_MULTI_MAX]; int nparts[WGRAD_MULTI_MAX]; };

// out[j][i] = sum_b partial[j][b][i], fixed order; blockIdx.y = problem
__global__ void __launch_bounds__(256) reduce_cta_partials_multi_kernel(const __grid_constant__ WgReduceMulti R, int elems) {
    FS_PDL_ENTER();
    __shared__ float red[8][33];
    const float* __restrict__ partial = R.partial[blockIdx.y];
    const int nparts = R.nparts[blockIdx.y];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lx;
    float s = 0.f;
    if (i < elems)
        for (int b = ly; b < nparts; b += 8) s += partial[(long long)b * elems + i];
    red[ly][lx] = s;
    __syncthreads();
    if (ly == 0 && i < elems) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += red[k][lx];
        R.out[blockIdx.y][i] = r;
    }
}

// ---------------------------------------------------------------------------------------------------
// Weight gradient of the 2x2-tap forms of the stride-2 / resize convolutions (initconv_1/2, upsample_0/1; see
// conv3x3_tc.cu "2x2-tap mode"):  dWc[a,b,cx,cy] = sum_{n,y,x} X[n, y+a, x+b, cx] * dY[n, y, x, cy]
// with (Cx, Cy) = (64, 128) or (128, 64); the 128-channel side is a space-to-depth view read through a 5-D tensor
// map whose 64-channel block index is the row parity.  Same structure as wgrad3x3_tc_kernel: per 8x16-pixel tile
// one dY tile and, per (kw, 64-channel block of X), one slab whose kh = 0 | 1 rows form the two 64-row halves of an
// M = 128 MN-major operand; N = Cy per MMA; 2*CBx accumulators [128, Cy] stay in TMEM (256 columns) over all
// tiles of the CTA; fixed-order reduction of the per-CTA partials [4, Cx, Cy].
struct Wg2Params {
    int tilesX, tilesY, total_tiles;
    int cbx, cby;              // 64-channel blocks of X / dY (one of them is 2)
    int x_s2d, dy_s2d;
    float* partial;            // [gridDim.x][4*Cx*Cy]
};
constexpr int W2_XSTAGE = 2 * SLAB_BYTES;              // hi + lo slab of one (kw, cbx)
constexpr int W2_DSTAGE_MAX = 2 * 2 * DT_BYTES;         // hi + lo of up to two 64-channel boxes
constexpr int W2_SMEM = STAGES * (W2_XSTAGE + W2_DSTAGE_MAX) + 1024 + 256;

__global__ void __launch_bounds__(256, 1)
wgrad2x2_tc_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                   const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo,
                   const Wg2Params p) {
    FS_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smemX = smem;
    uint8_t* smemD = smem + STAGES * W2_XSTAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemD + STAGES * W2_DSTAGE_MAX);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + STAGES;
    uint64_t* d_full = x_empty + STAGES;
    uint64_t* d_empty = d_full + STAGES;
    uint64_t* done = d_empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int Cy = p.cby * 64;
    const int dplane = p.cby * DT_BYTES;              // bytes of the hi (or lo) boxes of one dY stage

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmX_hi); prefetch_tmap(&tmX_lo); prefetch_tmap(&tmD_hi); prefetch_tmap(&tmD_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1);
            mbar_init(&d_full[i], 1); mbar_init(&d_empty[i], 1);
        }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FS_PDL_WAIT();                        // everything above is CTA-local setup
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int sx = 0, sd = 0; uint32_t px = 0, pd = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const int tx = t % p.tilesX;
            int r = t / p.tilesX;
            const int ty = r % p.tilesY;
            const int n = r / p.tilesY;
            const int y0 = ty * TH, x0 = tx * TW;
            mbar_wait(&d_empty[sd], pd ^ 1);
            uint8_t* dd = smemD + sd * W2_DSTAGE_MAX;
            mbar_expect_tx(&d_full[sd], (uint32_t)(2 * dplane));
            for (int b = 0; b < p.cby; ++b) {
                if (p.dy_s2d) {
                    tma_load_5d(dd + b * DT_BYTES, &tmD_hi, &d_full[sd], 0, x0, b, y0, n);
                    tma_load_5d(dd + dplane + b * DT_BYTES, &tmD_lo, &d_full[sd], 0, x0, b, y0, n);
                } else {
                    tma_load_4d(dd + b * DT_BYTES, &tmD_hi, &d_full[sd], b * 64, x0, y0, n);
                    tma_load_4d(dd + dplane + b * DT_BYTES, &tmD_lo, &d_full[sd], b * 64, x0, y0, n);
                }
            }
            if (++sd == STAGES) { sd = 0; pd ^= 1; }
            for (int kw = 0; kw < 2; ++kw) {
                for (int cb = 0; cb < p.cbx; ++cb) {
                    mbar_wait(&x_empty[sx], px ^ 1);
                    uint8_t* xd = smemX + sx * W2_XSTAGE;
                    mbar_expect_tx(&x_full[sx], W2_XSTAGE);
                    if (p.x_s2d) {
                        tma_load_5d(xd, &tmX_hi, &x_full[sx], 0, x0 + kw, cb, y0, n);
                        tma_load_5d(xd + SLAB_BYTES, &tmX_lo, &x_full[sx], 0, x0 + kw, cb, y0, n);
                    } else {
                        tma_load_4d(xd, &tmX_hi, &x_full[sx], cb * 64, x0 + kw, y0, n);
                        tma_load_4d(xd + SLAB_BYTES, &tmX_lo, &x_full[sx], cb * 64, x0 + kw, y0, n);
                    }
                    if (++sx == STAGES) { sx = 0; px ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_mn(Cy);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int sx = 0, sd = 0; uint32_t px = 0, pd = 0;
        bool first_tile = true;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            mbar_wait(&d_full[sd], pd);
            tc_fence_after();
            const uint32_t dbase = smem_u32(smemD + sd * W2_DSTAGE_MAX);
            const uint64_t dd_hi = make_sdesc_mn(dbase, (uint32_t)DT_BYTES);
            const uint64_t dd_lo = make_sdesc_mn(dbase + dplane, (uint32_t)DT_BYTES);
            for (int kw = 0; kw < 2; ++kw) {
                for (int cb = 0; cb < p.cbx; ++cb) {
                    mbar_wait(&x_full[sx], px);
                    tc_fence_after();
                    const uint64_t xd_hi = make_sdesc_mn(smem_u32(smemX + sx * W2_XSTAGE), 2048);
                    const uint64_t xd_lo = make_sdesc_mn(smem_u32(smemX + sx * W2_XSTAGE + SLAB_BYTES), 2048);
                    if (elect_one()) {
                        const uint32_t acc = tb + (uint32_t)((kw * p.cbx + cb) * Cy);
#pragma unroll
                        for (int prod = 0; prod < 3; ++prod) {
                            const uint64_t ad = (prod == 2 ? xd_lo : xd_hi);
                            const uint64_t bd = (prod == 1 ? dd_lo : dd_hi);
#pragma unroll
                            for (int ks = 0; ks < TH; ++ks)
                                tc_mma_bf16(acc, ad + (uint64_t)(ks * 128), bd + (uint64_t)(ks * 128), idesc,
                                            (first_tile && prod == 0 && ks == 0) ? 0u : 1u);
                        }
                        tc_commit(&x_empty[sx]);
                    }
                    __syncwarp();
                    if (++sx == STAGES) { sx = 0; px ^= 1; }
                }
            }
            if (elect_one()) tc_commit(&d_empty[sd]);
            __syncwarp();
            if (++sd == STAGES) { sd = 0; pd ^= 1; }
            first_tile = false;
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== dump (4 warps = 128 TMEM lanes) =====================
        const int ew = warp - 4;
        const int row = ew * 32 + lane;              // accumulator row: 0..63 tap row kh = 0, 64..127 kh = 1
        mbar_wait(done, 0);
        tc_fence_after();
        const int Cx = p.cbx * 64;
        float* part = p.partial + (long long)blockIdx.x * (4 * Cx * Cy);
        const bool has_tiles = (int)blockIdx.x < p.total_tiles;
        const int kh = row >> 6;
#pragma unroll 1
        for (int a = 0; a < 2 * p.cbx; ++a) {
            const int kw = a / p.cbx, cb = a - kw * p.cbx;
            const int tap = kh * 2 + kw, cx = cb * 64 + (row & 63);
            float* op = part + ((long long)tap * Cx + cx) * Cy;
#pragma unroll 1
            for (int ch = 0; ch < Cy / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(a * Cy + ch * 32), v);
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(op + ch * 32 + i) = has_tiles ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3])
                                                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// out[i] = sum_b partial[b][i], fixed order
__global__ void __launch_bounds__(256) reduce_cta_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                  int elems, int nparts) {
    FS_PDL_ENTER();
    __shared__ float red[8][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lx;
    float s = 0.f;
    if (i < elems)
        for (int b = ly; b < nparts; b += 8) s += partial[(long long)b * elems + i];
    red[ly][lx] = s;
    __syncthreads();
    if (ly == 0 && i < elems) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += red[k][lx];
        out[i] = r;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(CUtensorMap* tm, const __nv_bfloat16* base, int N, int H, int W, int box_h) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        FS_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && fp,
                 "cuTensorMapEncodeTiled is not available from the CUDA driver");
        enc = reinterpret_cast<EncodeTiledFn>(fp);
    }
    cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%dx%dx%dx64) failed: %d", N, H, W, (int)r);
    return 0;
}

}  // namespace

long long wgrad3x3_tc_partial_floats() { return 148LL * 9 * 64 * 64; }

// x [N,H,W,64] and dy [N,OH,OW,64] as split planes; out [3,3,64,64] fp32 (HWIO); pad = conv zero padding
int launch_wgrad3x3_tc(SplitPtr x, SplitPtr dy, float* out, float* partial, long long partial_cap, int N, int H,
                       int W, int OH, int OW, int pad, cudaStream_t st) {
    FS_CHECK(x.hi && x.lo && dy.hi && dy.lo && out && partial, "wgrad3x3_tc: NULL argument");
    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    FS_TRY(make_map(&tmX_hi, x.hi, N, H, W, TH + 2));
    FS_TRY(make_map(&tmX_lo, x.lo, N, H, W, TH + 2));
    FS_TRY(make_map(&tmD_hi, dy.hi, N, OH, OW, TH));
    FS_TRY(make_map(&tmD_lo, dy.lo, N, OH, OW, TH));
    WgParams p;
    p.N = N; p.OH = OH; p.OW = OW; p.pad = pad;
    p.tilesX = cdiv(OW, TW); p.tilesY = cdiv(OH, TH);
    p.total_tiles = N * p.tilesX * p.tilesY;
    p.partial = partial;
    int sms = 148;
    {
        int dev = 0; cudaGetDevice(&dev);
        int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        if (sms > 148) sms = 148;
    }
    int grid = p.total_tiles < sms ? p.total_tiles : sms;
    FS_CHECK(grid >= 1, "wgrad3x3_tc: empty problem");
    FS_CHECK((long long)grid * 9 * 64 * 64 <= partial_cap, "wgrad3x3_tc: partial workspace too small");
    FS_DYN_SMEM(wgrad3x3_tc_kernel, SMEM_BYTES);
    launch_k(wgrad3x3_tc_kernel, dim3(grid), dim3(256), SMEM_BYTES, st, tmX_hi, tmX_lo, tmD_hi, tmD_lo, p);
    FS_LAUNCH_CHECK();
    launch_k(reduce_cta_partials_kernel, dim3(cdiv(9 * 64 * 64, 32)), dim3(256), 0, st, partial, out, 9 * 64 * 64, grid);
    FS_LAUNCH_CHECK();
    return 0;
}


namespace {
// 5-D space-to-depth view {64 = (q,c), W, 2 (p), H, N} of a plain [N, 2H, 2W, 32] tensor (cf. conv3x3_tc.cu)
int make_map_s2d(CUtensorMap* tm, const __nv_bfloat16* base, int N, int H, int W, int box_h) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        FS_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && fp,
                 "cuTensorMapEncodeTiled is not available from the CUDA driver");
        enc = reinterpret_cast<EncodeTiledFn>(fp);
    }
    const cuuint64_t C0 = 32, FW = 2 * (cuuint64_t)W, FH = 2 * (cuuint64_t)H;
    cuuint64_t dims[5] = {64, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[4] = {2 * C0 * 2, FW * C0 * 2, 2 * FW * C0 * 2, FH * FW * C0 * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)TW, 1, (cuuint32_t)box_h, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(s2d %dx%dx%d) failed: %d", N, H, W, (int)r);
    return 0;
}
}  // namespace

long long wgrad2x2_tc_partial_floats() { return 148LL * 4 * 64 * 128; }

// x, dy: split planes in GEMM space [N, GH, GW, .]; the side flagged s2d has 128 channels and is the space-to-depth
// view of a plain [N, 2GH, 2GW, 32] tensor, the other side is a plain 64-channel tensor.  out [2,2,Cx,Cy] fp32.
int launch_wgrad2x2_tc(SplitPtr x, int x_s2d, SplitPtr dy, int dy_s2d, float* out, float* partial, long long partial_cap,
                       int N, int GH, int GW, cudaStream_t st) {
    FS_CHECK(x.hi && x.lo && dy.hi && dy.lo && out && partial, "wgrad2x2_tc: NULL argument");
    FS_CHECK((x_s2d != 0) != (dy_s2d != 0), "wgrad2x2_tc: exactly one side is the 128-channel space-to-depth view");
    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    if (x_s2d) {
        FS_TRY(make_map_s2d(&tmX_hi, x.hi, N, GH, GW, TH + 2));
        FS_TRY(make_map_s2d(&tmX_lo, x.lo, N, GH, GW, TH + 2));
        FS_TRY(make_map(&tmD_hi, dy.hi, N, GH, GW, TH));
        FS_TRY(make_map(&tmD_lo, dy.lo, N, GH, GW, TH));
    } else {
        FS_TRY(make_map(&tmX_hi, x.hi, N, GH, GW, TH + 2));
        FS_TRY(make_map(&tmX_lo, x.lo, N, GH, GW, TH + 2));
        FS_TRY(make_map_s2d(&tmD_hi, dy.hi, N, GH, GW, TH));
        FS_TRY(make_map_s2d(&tmD_lo, dy.lo, N, GH, GW, TH));
    }
    Wg2Params p;
    p.tilesX = cdiv(GW, TW); p.tilesY = cdiv(GH, TH);
    p.total_tiles = N * p.tilesX * p.tilesY;
    p.cbx = x_s2d ? 2 : 1; p.cby = dy_s2d ? 2 : 1;
    p.x_s2d = x_s2d; p.dy_s2d = dy_s2d;
    p.partial = partial;
    int sms = 148;
    {
        int dev = 0; cudaGetDevice(&dev);
        int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        if (sms > 148) sms = 148;
    }
    int grid = p.total_tiles < sms ? p.total_tiles : sms;
    FS_CHECK(grid >= 1, "wgrad2x2_tc: empty problem");
    const int elems = 4 * 64 * 128;
    FS_CHECK((long long)grid * elems <= partial_cap, "wgrad2x2_tc: partial workspace too small");
    FS_DYN_SMEM(wgrad2x2_tc_kernel, W2_SMEM);
    launch_k(wgrad2x2_tc_kernel, dim3(grid), dim3(256), W2_SMEM, st, tmX_hi, tmX_lo, tmD_hi, tmD_lo, p);
    FS_LAUNCH_CHECK();
    launch_k(reduce_cta_partials_kernel, dim3(cdiv(elems, 32)), dim3(256), 0, st, partial, out, elems, grid);
    FS_LAUNCH_CHECK();
    return 0;
}

// =====================================================================================
// Gram matrix on tcgen05:  G[n][c1][c2] = scale * sum_p F[n,p,c1] * F[n,p,c2]   (reference utils.py:76-81)
// Work item = (sample, 128-channel row tile, 64-channel column tile, pixel split); operands are MN-major
// boxes {64 ch, 128 px} of the split planes viewed as [N, HW, C].  HBM/L2-bound for C <= 128.
// =====================================================================================
namespace {

constexpr int GP = 128;                               // pixels per chunk
constexpr int GBOX = GP * 128;                        // bytes of one {64 ch, 128 px} box
constexpr int G_STAGE = 6 * GBOX;                     // A hi g0,g1 | A lo g0,g1 | B hi | B lo
constexpr int G_STAGES = 2;
constexpr int G_SMEM = G_STAGES * G_STAGE + 1024 + 256;

struct GramParams {
    int N, HW, C, mtiles, ntiles, ksplit, chunks_per_split;
    float* partial;            // [N][ksplit][C][C]
};

__global__ void __launch_bounds__(256, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const GramParams p) {
    FS_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE);
    uint64_t* full = bars;
    uint64_t* empty = full + G_STAGES;
    uint64_t* done = empty + G_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int item = blockIdx.x;
    const int ks = item % p.ksplit; item /= p.ksplit;
    const int nt = item % p.ntiles; item /= p.ntiles;
    const int mt = item % p.mtiles;
    const int n = item / p.mtiles;
    const bool two_groups = p.C >= 128;               // C == 64: the second 64-row half aliases the first
    const int pix0 = ks * p.chunks_per_split * GP;
    int nchunks = p.chunks_per_split;
    {
        int total = (p.HW + GP - 1) / GP;
        int first = ks * p.chunks_per_split;
        if (first + nchunks > total) nchunks = total - first;
        if (nchunks < 0) nchunks = 0;
    }

    if (warp == 0 && lane == 0) { prefetch_tmap(&tm_hi); prefetch_tmap(&tm_lo); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < G_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FS_PDL_WAIT();                        // everything above is CTA-local setup
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        int s = 0; uint32_t ph = 0;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* d = smem + s * G_STAGE;
            const int pix = pix0 + c * GP;
            mbar_expect_tx(&full[s], (two_groups ? 6 : 4) * GBOX);
            const int ca = mt * 128, cb = nt * 64;
            tma_load_4d(d, &tm_hi, &full[s], ca, pix, n, 0);
            tma_load_4d(d + 2 * GBOX, &tm_lo, &full[s], ca, pix, n, 0);
            if (two_groups) {
                tma_load_4d(d + GBOX, &tm_hi, &full[s], ca + 64, pix, n, 0);
                tma_load_4d(d + 3 * GBOX, &tm_lo, &full[s], ca + 64, pix, n, 0);
            }
            tma_load_4d(d + 4 * GBOX, &tm_hi, &full[s], cb, pix, n, 0);
            tma_load_4d(d + 5 * GBOX, &tm_lo, &full[s], cb, pix, n, 0);
            if (++s == G_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_mn();
        const uint32_t lbo = two_groups ? (uint32_t)GBOX : 0u;
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int s = 0; uint32_t ph = 0;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t base = smem_u32(smem + s * G_STAGE);
            const uint64_t a_hi = make_sdesc_mn(base, lbo), a_lo = make_sdesc_mn(base + 2 * GBOX, lbo);
            const uint64_t b_hi = make_sdesc_mn(base + 4 * GBOX, 2048), b_lo = make_sdesc_mn(base + 5 * GBOX, 2048);
            if (elect_one()) {
#pragma unroll
                for (int prod = 0; prod < 3; ++prod) {
                    const uint64_t ab = prod == 2 ? a_lo : a_hi, bb = prod == 1 ? b_lo : b_hi;
#pragma unroll
                    for (int k = 0; k < GP / 16; ++k)
                        tc_mma_bf16(tb, ab + (uint64_t)(k * 128), bb + (uint64_t)(k * 128), idesc,
                                    (c == 0 && prod == 0 && k == 0) ? 0u : 1u);
                }
                tc_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == G_STAGES) { s = 0; ph ^= 1; }
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        const int c1 = mt * 128 + row;
        float* out = p.partial + (((long long)n * p.ksplit + ks) * p.C + c1) * p.C + nt * 64;
        const bool ok = (two_groups || row < 64) && c1 < p.C;
#pragma unroll 1
        for (int ch = 0; ch < 2; ++ch) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ch * 32), v);
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 4)
                    *reinterpret_cast<float4*>(out + ch * 32 + i) =
                        nchunks > 0 ? make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 64);
}

// ---------------------------------------------------------------------------------------------------
// Version 2: full-width rows.  Work item = (sample, 128-channel row tile, pixel split).  Per pixel chunk the
// CTA loads ALL C/64 channel boxes of F once (hi + lo); the A operand (M = 128 rows = 2 boxes, LBO = box
// stride) is a SUBSET of those boxes and the B operand is all of them (N = min(C, 256) per MMA, two N-halves
// for C = 512), so F crosses L2->SMEM C/128 times instead of 1.5*C/64 times and every MMA is as wide as the
// shape allows (A re-use from shared memory grows with N: the version-1 N = 64 tiles were bound by the
// 128 B/clk shared-memory operand feed).  The accumulator [128, C] fp32 lives in TMEM (C <= 512 columns).
// Chunk size gp is chosen so that one stage is <= 64 KB; 3-6 stages deep.
struct Gram2Params {
    int N, HW, C, mtiles, ksplit, total_chunks;
    int gp, gbox, nbox, stages, stage_bytes, ncols;
    float* partial;            // [N][ksplit][C][C]
};

__global__ void __launch_bounds__(256, 1)
gram_tc2_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const Gram2Params p) {
    FS_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = full + 8;
    uint64_t* done = empty + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int item = blockIdx.x;
    const int ks = item % p.ksplit; item /= p.ksplit;
    const int mt = item % p.mtiles;
    const int n = item / p.mtiles;
    const bool two_groups = p.C >= 128;               // C == 64: the second 64-row half aliases the first
    // balanced pixel split: chunks [c_beg, c_end)
    const int c_beg = (int)((long long)ks * p.total_chunks / p.ksplit);
    const int c_end = (int)((long long)(ks + 1) * p.total_chunks / p.ksplit);
    const int nchunks = c_end - c_beg;
    const int plane = p.nbox * p.gbox;                // bytes of the hi (or lo) boxes of one stage

    if (warp == 0 && lane == 0) { prefetch_tmap(&tm_hi); prefetch_tmap(&tm_lo); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, (uint32_t)p.ncols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FS_PDL_WAIT();                        // everything above is CTA-local setup
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        int s = 0; uint32_t ph = 0;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* d = smem + s * p.stage_bytes;
            const int pix = (c_beg + c) * p.gp;
            mbar_expect_tx(&full[s], (uint32_t)(2 * plane));
            for (int b = 0; b < p.nbox; ++b) {
                tma_load_4d(d + b * p.gbox, &tm_hi, &full[s], b * 64, pix, n, 0);
                tma_load_4d(d + plane + b * p.gbox, &tm_lo, &full[s], b * 64, pix, n, 0);
            }
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        const int nn = p.C < 256 ? p.C : 256;         // MMA N
        const int nhalves = p.C / nn;                 // 1, or 2 for C = 512
        const uint32_t idesc = make_idesc_mn(nn);
        const uint32_t lboA = two_groups ? (uint32_t)p.gbox : 0u;
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        const int ksteps = p.gp / 16;
        int s = 0; uint32_t ph = 0;
        for (int c = 0; c < nchunks; ++c) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint32_t base = smem_u32(smem + s * p.stage_bytes);
            const uint32_t aoff = two_groups ? (uint32_t)(mt * 2 * p.gbox) : 0u;
            const uint64_t a_hi = make_sdesc_mn(base + aoff, lboA), a_lo = make_sdesc_mn(base + plane + aoff, lboA);
            if (elect_one()) {
                for (int h = 0; h < nhalves; ++h) {
                    const uint32_t boff = (uint32_t)(h * (nn / 64) * p.gbox);
                    const uint64_t b_hi = make_sdesc_mn(base + boff, (uint32_t)p.gbox);
                    const uint64_t b_lo = make_sdesc_mn(base + plane + boff, (uint32_t)p.gbox);
                    const uint32_t acc = tb + (uint32_t)(h * nn);
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        const uint64_t ab = prod == 2 ? a_lo : a_hi, bb = prod == 1 ? b_lo : b_hi;
                        for (int k = 0; k < ksteps; ++k)
                            tc_mma_bf16(acc, ab + (uint64_t)(k * 128), bb + (uint64_t)(k * 128), idesc,
                                        (c == 0 && prod == 0 && k == 0) ? 0u : 1u);
                    }
                }
                tc_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        const int c1 = mt * 128 + row;
        const bool ok = (two_groups || row < 64) && c1 < p.C;
        float* out = p.partial + (((long long)n * p.ksplit + ks) * p.C + (ok ? c1 : 0)) * p.C;
#pragma unroll 1
        for (int ch = 0; ch < p.C / 32; ++ch) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ch * 32), v);
            if (ok) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float z[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) z[j] = nchunks > 0 ? v[i + j] : 0.f;
                    stg256(out + ch * 32 + i, z);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, (uint32_t)p.ncols);
}

// G[n][i] = scale * sum_ks partial[n][ks][i]
__global__ void gram_reduce_kernel(const float* __restrict__ partial, float* __restrict__ G, long long cc4, int N,
                                   int ksplit, float scale) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cc4 * N) return;
    long long n = i / cc4, e = i - n * cc4;
    const float4* p = reinterpret_cast<const float4*>(partial) + n * ksplit * cc4 + e;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < ksplit; ++k) {
        float4 v = p[(long long)k * cc4];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(G)[i] = make_float4(s.x * scale, s.y * scale, s.z * scale, s.w * scale);
}

__global__ void pack_gemm_b_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ hi,
                                   __nv_bfloat16* __restrict__ lo, int N, int C) {
    FS_PDL_ENTER();
    // out index: (((n*CB + cb)*C + j)*64 + k)  <-  S[n][cb*64+k][j]
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * C * C;
    if (i >= total) return;
    int k = (int)(i & 63);
    long long r = i >> 6;
    int j = (int)(r % C); r /= C;
    int CB = C / 64;
    int cb = (int)(r % CB);
    int n = (int)(r / CB);
    float v = S[((long long)n * C + cb * 64 + k) * C + j];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

int gram_ksplit(int N, int HW, int C) {
    int items = N * ((C + 127) / 128) * (C / 64);
    int ks = (148 + items - 1) / items;
    int chunks = (HW + GP - 1) / GP;
    if (ks > chunks) ks = chunks;
    if (ks < 1) ks = 1;
    return ks;
}

int make_map3(CUtensorMap* tm, const __nv_bfloat16* base, int N, int HW, int C, int gp = GP) {
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fp = nullptr;
        cudaDriverEntryPointQueryResult q;
        FS_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess && fp,
                 "cuTensorMapEncodeTiled is not available from the CUDA driver");
        enc = reinterpret_cast<EncodeTiledFn>(fp);
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)HW, (cuuint64_t)N, 1};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)HW * C * 2, (cuuint64_t)N * HW * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)gp, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(gram %dx%dx%d) failed: %d", N, HW, C, (int)r);
    return 0;
}

}  // namespace

namespace {
// version-2 plan (see gram_tc2_kernel)
void gram2_plan(int N, int HW, int C, Gram2Params& p) {
    p.N = N; p.HW = HW; p.C = C;
    p.mtiles = C >= 128 ? C / 128 : 1;
    p.nbox = C / 64;
    p.gp = C <= 128 ? 128 : 16384 / C;              // C*gp*4 bytes per stage: 32 KB (C = 64) or 64 KB
    p.gbox = p.gp * 128;
    p.stage_bytes = 2 * p.nbox * p.gbox;
    p.stages = 196608 / p.stage_bytes;
    if (p.stages > 6) p.stages = 6;
    p.total_chunks = (HW + p.gp - 1) / p.gp;
    int items = N * p.mtiles;
    int ks = 148 / items;
    if (ks < 1) ks = 1;
    if (ks > p.total_chunks) ks = p.total_chunks;
    p.ksplit = ks;
    p.ncols = C < 32 ? 32 : C;                      // 64 / 128 / 256 / 512: powers of two
}
bool gram2_supported(int C) { return C == 64 || C == 128 || C == 256 || C == 512; }
int gram_version() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FS_GRAM_V2");
        v = (e && e[0] == '0') ? 1 : 2;
    }
    return v;
}
}  // namespace

long long gram_tc_partial_floats(int N, int HW, int C) {
    long long v1 = (long long)N * gram_ksplit(N, HW, C) * C * C;
    if (!gram2_supported(C)) return v1;
    Gram2Params p;
    gram2_plan(N, HW, C, p);
    long long v2 = (long long)N * p.ksplit * C * C;
    return v1 > v2 ? v1 : v2;
}

static int launch_gram_tc2(SplitPtr f, float* G, float* partial, long long partial_cap, int N, int HW, int C, float scale,
                           cudaStream_t st) {
    Gram2Params p;
    gram2_plan(N, HW, C, p);
    p.partial = partial;
    FS_CHECK((long long)N * p.ksplit * C * C <= partial_cap, "gram_tc2: partial workspace too small");
    CUtensorMap tm_hi, tm_lo;
    FS_TRY(make_map3(&tm_hi, f.hi, N, HW, C, p.gp));
    FS_TRY(make_map3(&tm_lo, f.lo, N, HW, C, p.gp));
    const int smem = p.stages * p.stage_bytes + 1024 + 256;
    FS_DYN_SMEM(gram_tc2_kernel, 196608 + 1024 + 256);
    int grid = N * p.mtiles * p.ksplit;
    launch_k(gram_tc2_kernel, dim3(grid), dim3(256), smem, st, tm_hi, tm_lo, p);
    FS_LAUNCH_CHECK();
    long long cc4 = (long long)C * C / 4;
    launch_k(gram_reduce_kernel, dim3(cdiv(cc4 * N, 256)), dim3(256), 0, st, partial, G, cc4, N, p.ksplit, scale);
    FS_LAUNCH_CHECK();
    return 0;
}

int launch_gram_tc(SplitPtr f, float* G, float* partial, long long partial_cap, int N, int HW, int C, float scale,
                   cudaStream_t st) {
    FS_CHECK(f.hi && f.lo && G && partial, "gram_tc: NULL argument");
    FS_CHECK(C % 64 == 0 && C >= 64, "gram_tc: C must be a multiple of 64");
    if (gram_version() == 2 && gram2_supported(C)) return launch_gram_tc2(f, G, partial, partial_cap, N, HW, C, scale, st);
    GramParams p;
    p.N = N; p.HW = HW; p.C = C;
    p.mtiles = (C + 127) / 128; p.ntiles = C / 64;
    p.ksplit = gram_ksplit(N, HW, C);
    int chunks = (HW + GP - 1) / GP;
    p.chunks_per_split = (chunks + p.ksplit - 1) / p.ksplit;
    p.partial = partial;
    FS_CHECK((long long)N * p.ksplit * C * C <= partial_cap, "gram_tc: partial workspace too small");
    CUtensorMap tm_hi, tm_lo;
    FS_TRY(make_map3(&tm_hi, f.hi, N, HW, C));
    FS_TRY(make_map3(&tm_lo, f.lo, N, HW, C));
    FS_DYN_SMEM(gram_tc_kernel, G_SMEM);
    int grid = N * p.mtiles * p.ntiles * p.ksplit;
    launch_k(gram_tc_kernel, dim3(grid), dim3(256), G_SMEM, st, tm_hi, tm_lo, p);
    FS_LAUNCH_CHECK();
    long long cc4 = (long long)C * C / 4;
    launch_k(gram_reduce_kernel, dim3(cdiv(cc4 * N, 256)), dim3(256), 0, st, partial, G, cc4, N, p.ksplit, scale);
    FS_LAUNCH_CHECK();
    return 0;
}

int pack_gemm_b_tc(const float* S, SplitPtr out, int N, int C, cudaStream_t st) {
    FS_CHECK(C % 64 == 0, "pack_gemm_b_tc: C must be a multiple of 64");
    long long total = (long long)N * C * C;
    launch_k(pack_gemm_b_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, S, out.hi, out.lo, N, C);
    FS_LAUNCH_CHECK();
    return 0;
}

// Batched weight gradients of `count` 3x3 64->64 convolutions (see wgrad3x3_tc_multi_kernel).  partial: workspace of
// wgrad3x3_tc_partial_floats() floats (148 partials in total, shared out over the problems).
int launch_wgrad3x3_tc_multi(const WgradMultiItem* items, int count, float* partial, long long partial_cap, int N,
                             cudaStream_t st) {
    FS_CHECK(count >= 1 && count <= WGRAD_MULTI_MAX, "wgrad3x3_tc_multi: 1..%d problems", WGRAD_MULTI_MAX);
    int sms = 148;
    {
        int dev = 0; cudaGetDevice(&dev);
        int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
        if (sms > 148) sms = 148;
    }
    FS_CHECK(148LL * 9 * 64 * 64 <= partial_cap, "wgrad3x3_tc_multi: partial workspace too small");
    WgMulti M;
    memset(&M, 0, sizeof(M));
    M.count = count;
    long long tiles[WGRAD_MULTI_MAX], total = 0;
    for (int j = 0; j < count; ++j) {
        const WgradMultiItem& it = items[j];
        FS_CHECK(it.x.hi && it.x.lo && it.dy.hi && it.dy.lo && it.out, "wgrad3x3_tc_multi: NULL argument");
        WgProblem& P = M.pr[j];
        FS_TRY(make_map(&P.tmX_hi, it.x.hi, N, it.H, it.W, TH + 2));
        FS_TRY(make_map(&P.tmX_lo, it.x.lo, N, it.H, it.W, TH + 2));
        FS_TRY(make_map(&P.tmD_hi, it.dy.hi, N, it.OH, it.OW, TH));
        FS_TRY(make_map(&P.tmD_lo, it.dy.lo, N, it.OH, it.OW, TH));
        P.pad = it.pad; P.tilesX = cdiv(it.OW, TW); P.tilesY = cdiv(it.OH, TH);
        P.total_tiles = N * P.tilesX * P.tilesY;
        P.out = it.out;
        tiles[j] = P.total_tiles; total += tiles[j];
    }
    // CTAs in proportion to the tile counts (at least one each), all SMs used
    int begin = 0, left = sms;
    for (int j = 0; j < count; ++j) {
        int c = j == count - 1 ? left : (int)((tiles[j] * sms + total / 2) / total);
        if (c < 1) c = 1;
        const int others = count - 1 - j;
        if (c > left - others) c = left - others;
        if (c > M.pr[j].total_tiles) c = M.pr[j].total_tiles;
        if (c < 1) c = 1;
        M.pr[j].cta_begin = begin; M.pr[j].cta_count = c;
        M.pr[j].partial = partial + (long long)begin * (9 * 64 * 64);
        begin += c; left -= c;
    }
    FS_CHECK(begin <= 148 && left >= 0, "wgrad3x3_tc_multi: CTA assignment overflow");
    FS_DYN_SMEM(wgrad3x3_tc_multi_kernel, SMEM_BYTES);
    launch_k(wgrad3x3_tc_multi_kernel, dim3(begin), dim3(256), SMEM_BYTES, st, M);
    FS_LAUNCH_CHECK();
    WgReduceMulti R;
    memset(&R, 0, sizeof(R));
    for (int j = 0; j < count; ++j) { R.partial[j] = M.pr[j].partial; R.out[j] = M.pr[j].out; R.nparts[j] = M.pr[j].cta_count; }
    launch_k(reduce_cta_partials_multi_kernel, dim3(cdiv(9 * 64 * 64, 32), count), dim3(256), 0, st, R, 9 * 64 * 64);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
