// tcgen05 weight gradient of the two 9x9 stride-1 SAME convolutions of the transform net (3(4) <-> 16 channels;
// reference im_transf_net.py:38,70 under tf.gradients) in their x8 forms.
//
// Both layers have a 4-channel side.  As in the forward x8 form (Engine::tc9) that side is held in WINDOWED planes
// Xw[n, y, X, 16 px x 4 ch] (group X = pixels 8X-4 .. 8X+11) and the 16-channel side in GROUPED planes
// Dg[n, y, X, 8 px x 16 ch] (= the plain NHWC tensor).  The weight gradient in Toeplitz space is one pixel-reduction GEMM
//
//     Wt[kh][k = (dxi, c4)][n = (dxo, c16)] = sum_{n,y,X} Xw[n, y + kh - 4, X, k] * Dg[n, y, X, n]
//
// (9 vertical taps, NO horizontal shift), folded back afterwards: dW[kh][kw] = sum_dxo Wt[kh][(dxo + kw, .)][(dxo, .)].
// Both operands are MN-major (one 128-byte row of 64 channels per GEMM pixel), exactly what TMA boxes of the planes
// produce.  Per tile (8 rows x 16 groups) a CTA loads ONE slab of Xw {64 ch, 16 groups, 16 rows} - tap kh is the slab at
// +kh * 2048 B - and one 64-channel half of the Dg tile; five M = 128 accumulators cover the nine taps in pairs
// ((0|1), (2|3), (4|5), (6|7), (7|8): the second 64 rows of an MN-major operand sit LBO = 2048 B = one slab row later;
// kh = 7 is computed twice and the duplicate dropped).  CTAs of even / odd index take the two halves of the 128 GEMM
// columns and keep their accumulators in TMEM over all their tiles; one partial per CTA, then one fold + reduction.
// MMA work: 2 * 576 * 128 per GEMM pixel = 2.4x the algorithmic FLOPs (x3 split-bf16 passes), and the layers' weight
// gradients take 40 / 65 us instead of 130 / 186 us on the CUDA-core kernel (which ran at half the FMA peak).
#include <cuda.h>
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace fs {

namespace {

using namespace tcptx;

constexpr int TW = 16, TH = 8;
constexpr int X_PLANE = (TH + 8) * TW * 128;         // 32 KB: 16 slab rows of 16 groups
constexpr int D_PLANE = TH * TW * 128;               // 16 KB
constexpr int X_STAGE = 2 * X_PLANE, D_STAGE = 2 * D_PLANE;
constexpr int STAGES = 2;
constexpr int W9_SMEM = STAGES * (X_STAGE + D_STAGE) + 1024 + 256;
constexpr int W9_TMEM = 512;                         // 5 accumulators x 64 columns (power of two >= 320)
constexpr int W9_PART = 9 * 64 * 64;                 // floats per CTA partial: [kh][k][n half]

__device__ __forceinline__ uint64_t sdesc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct W9Params { int tilesX, tilesY, total_tiles; float* partial; };

__global__ void __launch_bounds__(256, 1)
wgrad9_x8_tc_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                    const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo,
                    const W9Params p) {
    FS_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smemX = smem;
    uint8_t* smemD = smem + STAGES * X_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemD + STAGES * D_STAGE);
    uint64_t* full = bars;                // [STAGES] X slab + D tile of one tile
    uint64_t* empty = full + STAGES;      // [STAGES]
    uint64_t* done = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nb = (int)blockIdx.x & 1;                    // which 64 of the 128 GEMM columns
    const int cta = (int)blockIdx.x >> 1, nctas = (int)gridDim.x >> 1;
    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX_hi); prefetch_tmap(&tmX_lo); prefetch_tmap(&tmD_hi); prefetch_tmap(&tmD_lo); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, W9_TMEM);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    FS_PDL_WAIT();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int s = 0; uint32_t ph = 0;
        for (int t = cta; t < p.total_tiles; t += nctas) {
            const int tx = t % p.tilesX;
            const int r = t / p.tilesX;
            const int ty = r % p.tilesY, n = r / p.tilesY;
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* xd = smemX + s * X_STAGE;
            uint8_t* dd = smemD + s * D_STAGE;
            mbar_expect_tx(&full[s], X_STAGE + D_STAGE);
            tma_load_4d(xd, &tmX_hi, &full[s], 0, tx * TW, ty * TH - 4, n);
            tma_load_4d(xd + X_PLANE, &tmX_lo, &full[s], 0, tx * TW, ty * TH - 4, n);
            tma_load_4d(dd, &tmD_hi, &full[s], nb * 64, tx * TW, ty * TH, n);
            tma_load_4d(dd + D_PLANE, &tmD_lo, &full[s], nb * 64, tx * TW, ty * TH, n);
            if (++s == STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16)      // fp32 acc, bf16 x bf16, both MN-major
                             | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);       // N = 64, M = 128
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int s = 0; uint32_t ph = 0;
        bool first_tile = true;
        for (int t = cta; t < p.total_tiles; t += nctas) {
            mbar_wait(&full[s], ph);
            tc_fence_after();
            const uint64_t xd_hi = sdesc_mn(smem_u32(smemX + s * X_STAGE), 2048);
            const uint64_t xd_lo = sdesc_mn(smem_u32(smemX + s * X_STAGE + X_PLANE), 2048);
            const uint64_t dd_hi = sdesc_mn(smem_u32(smemD + s * D_STAGE), 2048);
            const uint64_t dd_lo = sdesc_mn(smem_u32(smemD + s * D_STAGE + D_PLANE), 2048);
            if (elect_one()) {
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    const int kh0 = g < 4 ? 2 * g : 7;                       // accumulator rows 0-63: kh0, 64-127: kh0 + 1
                    const uint32_t acc = tb + (uint32_t)(g * 64);
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        const uint64_t ad = (prod == 2 ? xd_lo : xd_hi) + (uint64_t)(kh0 * 128);   // + kh0 slab rows
                        const uint64_t bd = (prod == 1 ? dd_lo : dd_hi);
#pragma unroll
                        for (int ks = 0; ks < TH; ++ks)                      // one tile row (16 GEMM pixels) per MMA
                            tc_mma_bf16(acc, ad + (uint64_t)(ks * 128), bd + (uint64_t)(ks * 128), idesc,
                                        (first_tile && prod == 0 && ks == 0) ? 0u : 1u);
                    }
                }
                tc_commit(&empty[s]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ph ^= 1; }
            first_tile = false;
        }
        if (elect_one()) tc_commit(done);
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== dump (4 warps = 128 TMEM lanes) =====================
        const int ew = warp - 4;
        const int row = ew * 32 + lane;
        mbar_wait(done, 0);
        tc_fence_after();
        float* part = p.partial + (long long)blockIdx.x * W9_PART;
        const bool has_tiles = cta < p.total_tiles;
#pragma unroll 1
        for (int g = 0; g < 5; ++g) {
            const int kh = (g < 4 ? 2 * g : 7) + (row >> 6);
            const bool keep = !(g == 4 && row < 64);                         // (7|8): the first half repeats kh = 7
            const int k = row & 63;
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(g * 64 + ch * 32), v);
                if (keep) {
                    float* op = part + ((long long)kh * 64 + k) * 64 + ch * 32;
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
    if (warp == 2) tmem_dealloc(tmem_base, W9_TMEM);
}

// Fold the Toeplitz-space partials back and sum them over the CTAs.
//   mode 0 (windowed side = the conv INPUT, c4 = ci; grouped side = dY, c16 = co):  out[kh][kw][ci < A][co < 16]
//       = sum_dxo Wt[kh][(dxo + kw) * 4 + ci][dxo * 16 + co]
//   mode 1 (windowed side = dY, c4 = co; grouped side = the conv input, c16 = ci):  out[kh][kw][ci < 16][co < A]
//       = sum_dxo Wt[8 - kh][(dxo + 8 - kw) * 4 + co][dxo * 16 + ci]      (correlation in the other direction)
// Partials: [cta][9][64][64] with the column half nb = cta & 1.  Fixed summation order -> deterministic.
__global__ void __launch_bounds__(256) wgrad9_x8_fold_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                             int nparts, int A, int mode) {
    FS_PDL_ENTER();
    // one CTA per (kh, kw, c4); lane = (half of the dxo range, c16) reads 16 contiguous floats of a partial row; the
    // eight warps split the CTA pairs (fixed assignment and order -> deterministic)
    __shared__ float red[8][16];
    const int wid = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c4 = wid % A, kw = (wid / A) % 9, kh = wid / (9 * A);
    const int tkh = mode == 0 ? kh : 8 - kh, tkw = mode == 0 ? kw : 8 - kw;
    const int c16 = lane & 15, half = lane >> 4;
    const int pairs = nparts >> 1;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int dxo = 4 * half + d;
        const int n = dxo * 16 + c16, nbh = n >> 6;          // column half = dxo / 4 = `half`
        const int k = (dxo + tkw) * 4 + c4;
        const float* pp = partial + ((long long)nbh * 9 + tkh) * 4096 + k * 64 + (n & 63);
        for (int cp = warp; cp < pairs; cp += 8) s += pp[(long long)cp * (2 * 9 * 4096)];
    }
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    if (half == 0) red[warp][c16] = s;
    __syncthreads();
    if (warp == 0 && half == 0) {
        float r = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) r += red[w][c16];
        const int ci = mode == 0 ? c4 : c16, co = mode == 0 ? c16 : c4;
        const int Ci = mode == 0 ? A : 16, Co = mode == 0 ? 16 : A;
        out[((kh * 9 + kw) * Ci + ci) * Co + co] = r;
    }
}

}  // namespace

long long wgrad9_x8_partial_floats() { return 148LL * W9_PART; }

// xw: windowed x8 planes [N, H, G, 64] of the 4-channel side; dg: plain planes [N, H, 8 G, 16] of the 16-channel side
// (viewed as [N, H, G, 128]); out: the weight-gradient slot, [9,9,A,16] (mode 0) or [9,9,16,A] (mode 1), A <= 4
int launch_wgrad9_x8_tc(SplitPtr xw, SplitPtr dg, float* out, float* partial, long long partial_cap, int N, int H, int G,
                        int A, int mode, cudaStream_t st) {
    FS_CHECK(xw.hi && xw.lo && dg.hi && dg.lo && out && partial, "wgrad9_x8_tc: NULL argument");
    FS_CHECK(A >= 1 && A <= 4 && (mode == 0 || mode == 1) && N >= 1 && H >= 1 && G >= 1, "wgrad9_x8_tc: bad dims");
    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    FS_TRY(tc_make_map_nhwc(&tmX_hi, xw.hi, 2, N, H, G, 64, 64, TW, TH + 8, 128));
    FS_TRY(tc_make_map_nhwc(&tmX_lo, xw.lo, 2, N, H, G, 64, 64, TW, TH + 8, 128));
    FS_TRY(tc_make_map_nhwc(&tmD_hi, dg.hi, 2, N, H, G, 128, 64, TW, TH, 128));
    FS_TRY(tc_make_map_nhwc(&tmD_lo, dg.lo, 2, N, H, G, 128, 64, TW, TH, 128));
    W9Params p;
    p.tilesX = cdiv(G, TW); p.tilesY = cdiv(H, TH);
    p.total_tiles = N * p.tilesX * p.tilesY;
    p.partial = partial;
    int pairs = num_sms() / 2;
    if (pairs > 74) pairs = 74;
    if (pairs > p.total_tiles) pairs = p.total_tiles;
    FS_CHECK(pairs >= 1, "wgrad9_x8_tc: empty problem");
    const int grid = 2 * pairs;
    FS_CHECK((long long)grid * W9_PART <= partial_cap, "wgrad9_x8_tc: partial workspace too small");
    FS_DYN_SMEM(wgrad9_x8_tc_kernel, W9_SMEM);
    launch_k(wgrad9_x8_tc_kernel, dim3(grid), dim3(256), W9_SMEM, st, tmX_hi, tmX_lo, tmD_hi, tmD_lo, p);
    FS_LAUNCH_CHECK();
    launch_k(wgrad9_x8_fold_kernel, dim3(81 * A), dim3(256), 0, st, (const float*)partial, out, grid, A, mode);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
