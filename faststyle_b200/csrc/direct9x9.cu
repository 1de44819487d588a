// Direct (shared-memory tiled) kernels for the two 9x9 stride-1 SAME convolutions of the
// transform net (initconv_0: 3(4)->16, upsample_2: 16->3(4); reference im_transf_net.py:38,70).
// Their GEMM view has N or K-per-tap of 4..16, which starves both the tensor pipe and the generic
// implicit-GEMM tiles (the im2col gather dominates).  Here a CTA stages a 32x32 pixel tile (+4 px
// halo) once in shared memory and every thread owns one (tap, 4x4 channel block) of the
// 81 x CI x CO weight-gradient, so each staged element is reused 81 times from SMEM.
// A 4-channel side is always RGB padded with a zero 4th channel here (the engine is the only caller): the
// kernels skip that channel's FMAs (a quarter of the work) - its inputs are zero and its outputs stay zero.
#include "common.cuh"
#include <cuda_bf16.h>
#include "tc_ptx.cuh"

namespace fs {

namespace {

constexpr int TS = 32;            // tile side (output pixels)
constexpr int HS = TS + 8;        // with the 4-px halo of a 9x9 window
constexpr int TSY = 16, HSY = TSY + 8;   // weight-gradient tiles are 16 rows x 32 columns: 2-3 CTAs per SM
constexpr int NT9 = 256;          // 2 row-groups x 108 workers (9 kh x 3 kw-triples x 4 channel blocks)
constexpr int WG9 = 108;

// dW[kh,kw,ci,co] partial sums of one tile: partial[2*block + rowgroup][(tap*CI + ci)*CO + co].
// A worker owns THREE horizontally adjacent taps (kw = 3g..3g+2) of one 4x4 channel block and walks a row of the
// tile with a 3-deep register window over the input: one new input float4 + one dy float4 per pixel feed 36-48
// FMAs (the first version - one tap per thread, 2 loads per 16 FMAs - ran at 92 % of the shared-memory pipe and
// 46 % of the FMA pipe, profiles/r01t).  The two row groups take rows 0-7 / 8-15 of the tile.
template <int CI, int CO>
__global__ void __launch_bounds__(NT9) wgrad9x9_kernel(const float* __restrict__ in, const float* __restrict__ dy,
                                                       float* __restrict__ partial, int H, int W, int tilesX,
                                                       int tilesY, int total_tiles) {
    FS_PDL_ENTER();
    static_assert(CI * CO == 64 && CI % 4 == 0 && CO % 4 == 0, "channel block must be 4x16 or 16x4");
    constexpr int CIQ = CI / 4, COQ = CO / 4;
    constexpr int IR = CI == 4 ? 3 : 4, JR = CO == 4 ? 3 : 4;      // a 4-channel side is zero-padded RGB
    extern __shared__ float4 sm4[];
    float4* in_s = sm4;                         // [HSY][HS][CIQ]
    float4* dy_s = sm4 + HSY * HS * CIQ;        // [TSY][TS][COQ]
    const int t = threadIdx.x;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool worker = t < 2 * WG9;
    const int rg = t / WG9, r = t - rg * WG9;           // row group, worker
    const int q = r & 3, tg = r >> 2;                    // channel block, tap triple 0..26
    const int kh = tg / 3, kw0 = (tg - kh * 3) * 3;
    const int ciq = q / COQ, coq = q - ciq * COQ;
    float acc[3][4][4];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][i][j] = 0.f;
    // persistent over tiles: the accumulators stay in registers, so a CTA writes ONE partial per row group
    // (148*3 CTAs x 2 x 20 KB = 18 MB instead of one per tile = 77 MB at batch 8, 336^2)
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tx = tile % tilesX;
        const int rr = tile / tilesX;
        const int ty = rr % tilesY, n = rr / tilesY;
        const int x0 = tx * TS, y0 = ty * TSY;
        const float4* in4 = reinterpret_cast<const float4*>(in + (long long)n * H * W * CI);
        const float4* dy4 = reinterpret_cast<const float4*>(dy + (long long)n * H * W * CO);
        __syncthreads();                                 // the previous tile's readers are done
        for (int i = t; i < HSY * HS * CIQ; i += NT9) {
            int pix = i / CIQ, c4 = i - pix * CIQ;
            int yy = y0 - 4 + pix / HS, xx = x0 - 4 + pix % HS;
            in_s[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(in4 + ((long long)yy * W + xx) * CIQ + c4) : z;
        }
        for (int i = t; i < TSY * TS * COQ; i += NT9) {
            int pix = i / COQ, c4 = i - pix * COQ;
            int yy = y0 + pix / TS, xx = x0 + pix % TS;
            dy_s[i] = (yy < H && xx < W) ? __ldg(dy4 + ((long long)yy * W + xx) * COQ + c4) : z;
        }
        __syncthreads();
        if (!worker) continue;
        for (int py = rg * (TSY / 2); py < (rg + 1) * (TSY / 2); ++py) {
            const float4* ip = in_s + ((py + kh) * HS + kw0) * CIQ + ciq;
            const float4* dp = dy_s + (py * TS) * COQ + coq;
            float4 a0 = ip[0], a1 = ip[CIQ];
#pragma unroll 8
            for (int px = 0; px < TS; ++px) {
                const float4 a2 = ip[(px + 2) * CIQ];
                const float4 b = dp[px * COQ];
                const float av[3][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}, {a2.x, a2.y, a2.z, a2.w}};
                const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int i = 0; i < IR; ++i)
#pragma unroll
                        for (int j = 0; j < JR; ++j) acc[k][i][j] = fmaf(av[k][i], bv[j], acc[k][i][j]);
                a0 = a1; a1 = a2;
            }
        }
    }
    if (!worker) return;
    float* outp = partial + ((long long)blockIdx.x * 2 + rg) * (81 * 64);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int tap = kh * 9 + kw0 + k;
        float* out = outp + (tap * CI + ciq * 4) * CO + coq * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<float4*>(out + i * CO) = make_float4(acc[k][i][0], acc[k][i][1], acc[k][i][2], acc[k][i][3]);
    }
}

// Direct forward convolution (also the data gradient, with flipped/transposed weights).
// CTA = 32x32 output pixels; warp w owns columns 4w..4w+3, lane = row: every thread keeps a
// 4-pixel x CO register tile, slides a 12-pixel input window along x per (kh, 4-channel group) and
// reads the weights as warp-wide broadcasts: >= 12 FMAs per shared-memory load.
constexpr int PITCH = HS + 1;     // row pitch (float4) that spreads the 32 row-lanes over all banks

template <int CI, int CO>
__global__ void __launch_bounds__(256, 2) conv9x9_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                         float* __restrict__ out, int H, int W) {
    FS_PDL_ENTER();
    constexpr int CIQ = CI / 4, COQ = CO / 4;
    extern __shared__ float4 sm4[];
    float4* in_s = sm4;                              // [CIQ][HS][PITCH]
    float4* w_s = sm4 + CIQ * HS * PITCH;            // [81][CI][COQ]
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS, n = blockIdx.z;
    const float4* in4 = reinterpret_cast<const float4*>(in + (long long)n * H * W * CI);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = t; i < HS * HS * CIQ; i += 256) {
        int pix = i / CIQ, c4 = i - pix * CIQ;
        int py = pix / HS, px = pix - py * HS;
        int yy = y0 - 4 + py, xx = x0 - 4 + px;
        in_s[(c4 * HS + py) * PITCH + px] =
            (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(in4 + ((long long)yy * W + xx) * CIQ + c4) : z;
    }
    const float4* w4 = reinterpret_cast<const float4*>(w);
    for (int i = t; i < 81 * CI * COQ; i += 256) w_s[i] = __ldg(w4 + i);
    __syncthreads();

    const int wq = t >> 5, lane = t & 31;
    float acc[4][CO];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < CO; ++j) acc[i][j] = 0.f;
#pragma unroll 1
    for (int kh = 0; kh < 9; ++kh) {
#pragma unroll 1
        for (int cq = 0; cq < CIQ; ++cq) {
            float iv[12][4];
            const float4* ip = in_s + (cq * HS + lane + kh) * PITCH + 4 * wq;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                float4 v = ip[j];
                iv[j][0] = v.x; iv[j][1] = v.y; iv[j][2] = v.z; iv[j][3] = v.w;
            }
            const float4* wp = w_s + ((kh * 9) * CI + cq * 4) * COQ;
#pragma unroll
            for (int kw = 0; kw < 9; ++kw) {
#pragma unroll
                for (int c = 0; c < (CI == 4 ? 3 : 4); ++c) {
#pragma unroll
                    for (int q = 0; q < COQ; ++q) {
                        const float4 wv = wp[(kw * CI + c) * COQ + q];
#pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            const float xv = iv[px + kw][c];
                            acc[px][q * 4 + 0] = fmaf(xv, wv.x, acc[px][q * 4 + 0]);
                            acc[px][q * 4 + 1] = fmaf(xv, wv.y, acc[px][q * 4 + 1]);
                            acc[px][q * 4 + 2] = fmaf(xv, wv.z, acc[px][q * 4 + 2]);
                            if (CO != 4) acc[px][q * 4 + 3] = fmaf(xv, wv.w, acc[px][q * 4 + 3]);
                        }
                    }
                }
            }
        }
    }
    const int oy = y0 + lane;
    if (oy >= H) return;
    float* orow = out + (((long long)n * H + oy) * W) * CO;
    if (CO == 4 && (W & 1) == 0 && x0 + 4 * wq + 3 < W) {        // 4 pixels x 4 channels = 64 contiguous bytes: two 256-bit stores
        float r[16];
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int j = 0; j < 4; ++j) r[px * 4 + j] = acc[px][j];
        tcptx::stg256(orow + (long long)(x0 + 4 * wq) * CO, r);
        tcptx::stg256(orow + (long long)(x0 + 4 * wq) * CO + 8, r + 8);
        return;
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
        const int ox = x0 + 4 * wq + px;
        if (ox >= W) continue;
        if (CO == 16) {
            tcptx::stg256(orow + (long long)ox * CO, &acc[px][0]);
            tcptx::stg256(orow + (long long)ox * CO + 8, &acc[px][8]);
        } else {
#pragma unroll
            for (int q = 0; q < COQ; ++q)
                *reinterpret_cast<float4*>(orow + (long long)ox * CO + q * 4) =
                    make_float4(acc[px][q * 4], acc[px][q * 4 + 1], acc[px][q * 4 + 2], acc[px][q * 4 + 3]);
        }
    }
}

// =====================================================================================
// VGG conv1_1: 3x3 SAME, 3(4) -> 64, + bias + ReLU (reference libs/vgg16.py:45-55) and its data gradient
// (64 -> 4).  Tile = 32 rows x 8 columns; lane = row, warp = (4-pixel column group, 16-channel group).
// =====================================================================================
constexpr int C1_ROWS = 32, C1_COLS = 8, C1_PITCH = 11;   // pitch 11 float4: the 32 row-lanes hit distinct banks

__global__ void __launch_bounds__(256) conv3x3_c4_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             __nv_bfloat16* __restrict__ shi, __nv_bfloat16* __restrict__ slo,
                                                             unsigned char* __restrict__ code, int H, int W) {
    FS_PDL_ENTER();
    __shared__ float4 in_s[(C1_ROWS + 2) * C1_PITCH];
    __shared__ float4 w_s[9 * 4 * 16];
    __shared__ float4 b_s[16];
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * C1_COLS, y0 = blockIdx.y * C1_ROWS, n = blockIdx.z;
    const float4* in4 = reinterpret_cast<const float4*>(in + (long long)n * H * W * 4);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = t; i < (C1_ROWS + 2) * (C1_COLS + 2); i += 256) {
        int py = i / (C1_COLS + 2), px = i - py * (C1_COLS + 2);
        int yy = y0 - 1 + py, xx = x0 - 1 + px;
        in_s[py * C1_PITCH + px] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(in4 + (long long)yy * W + xx) : z;
    }
    for (int i = t; i < 9 * 4 * 16; i += 256) w_s[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
    if (t < 16) b_s[t] = bias ? __ldg(reinterpret_cast<const float4*>(bias) + t) : z;
    __syncthreads();
    const int lane = t & 31, wq = t >> 5;
    const int g = wq & 1, cog = wq >> 1;            // column group (4 px), output-channel group (16 ch)
    float acc[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        float iv[6][4];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            float4 v = in_s[(lane + kh) * C1_PITCH + 4 * g + j];
            iv[j][0] = v.x; iv[j][1] = v.y; iv[j][2] = v.z; iv[j][3] = v.w;
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int c = 0; c < 3; ++c)          // the 4th input channel is the zero pad of RGB
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = w_s[((kh * 3 + kw) * 4 + c) * 16 + cog * 4 + q];
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float xv = iv[px + kw][c];
                        acc[px][q * 4 + 0] = fmaf(xv, wv.x, acc[px][q * 4 + 0]);
                        acc[px][q * 4 + 1] = fmaf(xv, wv.y, acc[px][q * 4 + 1]);
                        acc[px][q * 4 + 2] = fmaf(xv, wv.z, acc[px][q * 4 + 2]);
                        acc[px][q * 4 + 3] = fmaf(xv, wv.w, acc[px][q * 4 + 3]);
                    }
                }
    }
    const int oy = y0 + lane;
    if (oy >= H) return;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
        const int ox = x0 + 4 * g + px;
        if (ox >= W) continue;
        const long long o = (((long long)n * H + oy) * W + ox) * 64 + cog * 16;
        float r[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 b = b_s[cog * 4 + q];
            r[q * 4 + 0] = fmaxf(acc[px][q * 4 + 0] + b.x, 0.f);
            r[q * 4 + 1] = fmaxf(acc[px][q * 4 + 1] + b.y, 0.f);
            r[q * 4 + 2] = fmaxf(acc[px][q * 4 + 2] + b.z, 0.f);
            r[q * 4 + 3] = fmaxf(acc[px][q * 4 + 3] + b.w, 0.f);
        }
        if (out) {
            tcptx::stg256(out + o, r);
            tcptx::stg256(out + o + 8, r + 8);
        }
        if (shi) {
            uint32_t hw[8], lw[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                __nv_bfloat16 h0 = __float2bfloat16_rn(r[2 * j]), h1 = __float2bfloat16_rn(r[2 * j + 1]);
                __nv_bfloat16 l0 = __float2bfloat16_rn(r[2 * j] - __bfloat162float(h0));
                __nv_bfloat16 l1 = __float2bfloat16_rn(r[2 * j + 1] - __bfloat162float(h1));
                hw[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lw[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            tcptx::stg256_b32(shi + o, hw);
            tcptx::stg256_b32(slo + o, lw);
        }
        if (code) {          // ReLU codes for the backward pass (Conv3x3TcArgs::ref_code): bit 0 = value > 0
            uint32_t cw[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                cw[j] = (r[4 * j] > 0.f ? 1u : 0u) | (r[4 * j + 1] > 0.f ? 0x100u : 0u) |
                        (r[4 * j + 2] > 0.f ? 0x10000u : 0u) | (r[4 * j + 3] > 0.f ? 0x1000000u : 0u);
            *reinterpret_cast<uint4*>(code + o) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
        }
    }
}

// dX[p][0..3] = sum_{tap,co} P[p + tap - 1][co] * Wf[tap][co][0..3]   (Wf = flipped/transposed conv1_1 weights)
// warp = (column group, 16-channel slice of the reduction); the 4 slices are summed through shared memory.
constexpr int C1_QP = 4;     // channel quads staged per pass (16 of the 64 channels): 45 KB of shared memory, 5 CTAs per SM

__global__ void __launch_bounds__(256) dgrad3x3_c4_kernel(const float* __restrict__ P, const float* __restrict__ wf,
                                                          float* __restrict__ dx, int H, int W) {
    FS_PDL_ENTER();
    extern __shared__ float4 sm4[];
    float4* in_s = sm4;                                         // [C1_QP channel quads][ROWS+2][PITCH]
    float4* w_s = sm4 + C1_QP * (C1_ROWS + 2) * C1_PITCH;       // [9][64] float4
    float4* red = w_s + 9 * 64;                                 // [3 slices][64 threads][4 px]
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * C1_COLS, y0 = blockIdx.y * C1_ROWS, n = blockIdx.z;
    const float4* p4 = reinterpret_cast<const float4*>(P + (long long)n * H * W * 64);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = t; i < 9 * 64; i += 256) w_s[i] = __ldg(reinterpret_cast<const float4*>(wf) + i);
    const int lane = t & 31, wq = t >> 5;
    const int g = wq & 1, kq = wq >> 1;               // column group, channel quad of the pass owned by this warp
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // The 64 gradient channels are staged 16 at a time (the first version staged all 64: 117 KB of shared memory,
    // ONE resident CTA per SM, 12 % warp occupancy and a latency-bound load phase - profiles/r01t).
#pragma unroll 1
    for (int pass = 0; pass < 16 / C1_QP; ++pass) {
        if (pass) __syncthreads();
        for (int i = t; i < (C1_ROWS + 2) * (C1_COLS + 2) * C1_QP; i += 256) {
            int cq = i % C1_QP, pix = i / C1_QP;
            int py = pix / (C1_COLS + 2), px = pix - py * (C1_COLS + 2);
            int yy = y0 - 1 + py, xx = x0 - 1 + px;
            in_s[(cq * (C1_ROWS + 2) + py) * C1_PITCH + px] =
                (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(p4 + ((long long)yy * W + xx) * 16 + pass * C1_QP + cq) : z;
        }
        __syncthreads();
        const int cqg = pass * C1_QP + kq;            // global channel quad
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            float iv[6][4];
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                float4 v = in_s[(kq * (C1_ROWS + 2) + lane + kh) * C1_PITCH + 4 * g + j];
                iv[j][0] = v.x; iv[j][1] = v.y; iv[j][2] = v.z; iv[j][3] = v.w;
            }
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float4 wv = w_s[(kh * 3 + kw) * 64 + cqg * 4 + c];
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float xv = iv[px + kw][c];
                        acc[px][0] = fmaf(xv, wv.x, acc[px][0]);
                        acc[px][1] = fmaf(xv, wv.y, acc[px][1]);
                        acc[px][2] = fmaf(xv, wv.z, acc[px][2]);   // channel 3 = zero pad of RGB: stays 0
                    }
                }
        }
    }
    const int slot = g * 32 + lane;
    if (kq > 0) {
#pragma unroll
        for (int px = 0; px < 4; ++px)
            red[((kq - 1) * 64 + slot) * 4 + px] = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
    }
    __syncthreads();
    if (kq == 0) {
        const int oy = y0 + lane;
        if (oy >= H) return;
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            const int ox = x0 + 4 * g + px;
            if (ox >= W) continue;
            float4 s = make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float4 v = red[(k * 64 + slot) * 4 + px];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            *reinterpret_cast<float4*>(dx + (((long long)n * H + oy) * W + ox) * 4) = s;
        }
    }
}

// Wf[80-tap][co][ci] = W[tap][ci][co]: weights that turn the data gradient into a forward convolution
__global__ void flip_transpose_taps_kernel(const float* __restrict__ W, float* __restrict__ Wf, int T, int Ci, int Co) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)T * Ci * Co) return;
    int ci = (int)(i % Ci);
    long long r = i / Ci;
    int co = (int)(r % Co);
    int tf = (int)(r / Co);
    Wf[i] = W[((long long)(T - 1 - tf) * Ci + ci) * Co + co];
}

// out[i] = sum_b partial[b][i]; 32 elements x 8 block-slices per CTA, fixed summation order
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                              int elems, int nblocks) {
    FS_PDL_ENTER();
    __shared__ float red[8][33];
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + lx;
    float s = 0.f;
    if (i < elems)
        for (int b = ly; b < nblocks; b += 8) s += partial[(long long)b * elems + i];
    red[ly][lx] = s;
    __syncthreads();
    if (ly == 0 && i < elems) {
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += red[k][lx];
        out[i] = r;
    }
}

}  // namespace

// 9x9 stride-1 SAME convolution, in [N,H,W,CI] -> out [N,H,W,CO], w [81,CI,CO]; (CI,CO) in {(4,16),(16,4)}
int launch_conv9x9(const float* in, const float* w, float* out, int N, int H, int W, int CI, int CO, cudaStream_t st) {
    FS_CHECK((CI == 16 && CO == 4) || (CI == 4 && CO == 16), "conv9x9: unsupported channel block %dx%d", CI, CO);
    dim3 grid(cdiv(W, TS), cdiv(H, TS), N);
    const size_t smem = (size_t)((CI / 4) * HS * PITCH + 81 * CI * (CO / 4)) * sizeof(float4);
    if (CI == 4) {
        FS_DYN_SMEM((conv9x9_kernel<4, 16>), smem);
        launch_k((conv9x9_kernel<4, 16>), dim3(grid), dim3(256), smem, st, in, w, out, H, W);
    } else {
        FS_DYN_SMEM((conv9x9_kernel<16, 4>), smem);
        launch_k((conv9x9_kernel<16, 4>), dim3(grid), dim3(256), smem, st, in, w, out, H, W);
    }
    FS_LAUNCH_CHECK();
    return 0;
}

// conv1_1 forward: in [N,H,W,4], w [9,4,64], bias [64] -> out [N,H,W,64] fp32 (+ optional split planes)
int launch_conv3x3_c4_fwd(const float* in, const float* w, const float* bias, float* out, void* split_hi, void* split_lo,
                          int N, int H, int W, cudaStream_t st, unsigned char* code) {
    FS_CHECK(out || split_hi || code, "conv3x3_c4_fwd: no output requested");
    dim3 grid(cdiv(W, C1_COLS), cdiv(H, C1_ROWS), N);
    launch_k(conv3x3_c4_fwd_kernel, dim3(grid), dim3(256), 0, st, in, w, bias, out, (__nv_bfloat16*)split_hi, (__nv_bfloat16*)split_lo, code, H, W);
    FS_LAUNCH_CHECK();
    return 0;
}

// conv1_1 data gradient: P [N,H,W,64], wf [9,64,4] (flip_transpose_taps of the forward weights) -> dx [N,H,W,4]
int launch_dgrad3x3_c4(const float* P, const float* wf, float* dx, int N, int H, int W, cudaStream_t st) {
    const size_t smem = (size_t)(C1_QP * (C1_ROWS + 2) * C1_PITCH + 9 * 64 + 3 * 64 * 4) * sizeof(float4);
    FS_DYN_SMEM((dgrad3x3_c4_kernel), smem);
    dim3 grid(cdiv(W, C1_COLS), cdiv(H, C1_ROWS), N);
    launch_k(dgrad3x3_c4_kernel, dim3(grid), dim3(256), smem, st, P, wf, dx, H, W);
    FS_LAUNCH_CHECK();
    return 0;
}

int flip_transpose_taps(const float* W, float* Wf, int T, int Ci, int Co, cudaStream_t st) {
    long long n = (long long)T * Ci * Co;
    launch_k(flip_transpose_taps_kernel, dim3(cdiv(n, 256)), dim3(256), 0, st, W, Wf, T, Ci, Co);
    FS_LAUNCH_CHECK();
    return 0;
}

constexpr int WG9_MAX_CTAS = 148 * 3;

long long wgrad9x9_partial_floats(int N, int H, int W) {
    long long tiles = (long long)N * cdiv(H, TSY) * cdiv(W, TS);
    if (tiles > WG9_MAX_CTAS) tiles = WG9_MAX_CTAS;
    return 2LL * tiles * 81 * 64;      // two row groups per CTA
}

// in [N,H,W,CI], dy [N,H,W,CO] (9x9 stride-1 SAME conv: equal spatial dims); out [81,CI,CO]
int launch_wgrad9x9(const float* in, const float* dy, float* out, float* partial, long long partial_cap, int N,
                    int H, int W, int CI, int CO, cudaStream_t st) {
    FS_CHECK((CI == 16 && CO == 4) || (CI == 4 && CO == 16), "wgrad9x9: unsupported channel block %dx%d", CI, CO);
    long long need = wgrad9x9_partial_floats(N, H, W);
    FS_CHECK(need <= partial_cap, "wgrad9x9: partial workspace too small (%lld > %lld floats)", need, partial_cap);
    const int tilesX = cdiv(W, TS), tilesY = cdiv(H, TSY);
    const int total = N * tilesX * tilesY;
    const int nblocks = total < WG9_MAX_CTAS ? total : WG9_MAX_CTAS;
    if (CI == 16) {
        size_t smem = (size_t)(HSY * HS * 4 + TSY * TS * 1) * sizeof(float4);
        FS_DYN_SMEM((wgrad9x9_kernel<16, 4>), smem);
        launch_k((wgrad9x9_kernel<16, 4>), dim3(nblocks), dim3(NT9), smem, st, in, dy, partial, H, W, tilesX, tilesY, total);
    } else {
        size_t smem = (size_t)(HSY * HS * 1 + TSY * TS * 4) * sizeof(float4);
        FS_DYN_SMEM((wgrad9x9_kernel<4, 16>), smem);
        launch_k((wgrad9x9_kernel<4, 16>), dim3(nblocks), dim3(NT9), smem, st, in, dy, partial, H, W, tilesX, tilesY, total);
    }
    FS_LAUNCH_CHECK();
    launch_k(reduce_partials_kernel, dim3(cdiv(81 * 64, 32)), dim3(256), 0, st, partial, out, 81 * 64, 2 * nblocks);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
