// HBM-bound kernels of the hot path: padding, InstanceNorm (fwd/bwd), pooling,
// loss reductions + their gradients, TF-Adam, weight-layout helpers.
// All tensors are fp32 NHWC with channel counts that are multiples of 4 (RGB
// tensors are carried as 4 channels with a zero 4th channel inside the engine).
#include "ops.cuh"
#include <cuda_bf16.h>

namespace fs {

namespace {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of a double, result valid in thread 0
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double red[32];
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = l < (int)((blockDim.x + 31) >> 5) ? red[l] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// optional split-bf16 companion output (x ~= hi + lo) for the tensor-core convolutions
__device__ __forceinline__ void store_split4(__nv_bfloat16* hi, __nv_bfloat16* lo, long long idx, const float* r) {
    uint32_t h[2], l[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {           // packed conversions: one cvt.rn.bf16x2.f32 per two values
        const __nv_bfloat162 hh = __floats2bfloat162_rn(r[2 * j], r[2 * j + 1]);
        const float2 hf = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(r[2 * j] - hf.x, r[2 * j + 1] - hf.y);
        h[j] = *reinterpret_cast<const uint32_t*>(&hh);
        l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(lo + idx) = make_uint2(l[0], l[1]);
}

__device__ __forceinline__ float in_affine(float x, float mean, float rstd, float g, float b) {
    return __fmaf_rn((x - mean) * rstd, g, b);
}

// ------------------------------------------------------------------ padding
// x16hi / x16lo (optional): the same padded image as split-bf16 planes [N, OH, OW + 16, 4] at column offset 4 - the
// zero-margined input of the x16 space-to-depth 9x9 form of initconv_0 (Engine::tc9)
// x8 != 0: the planes of the x8 form instead, [N, OH, OW / 8, 16 px x 4]: group X holds the padded-row columns
// 8X .. 8X+15 (pixels 8X-4 .. 8X+11), so every pixel is stored into the two groups that overlap it
__device__ __forceinline__ void store_x8(__nv_bfloat16* hi, __nv_bfloat16* lo, long long row, int col, int G, const float* r4) {
    const int X = col >> 3;
    if (X < G) store_split4(hi, lo, ((row * G + X) * 16 + (col - 8 * X)) * 4, r4);
    if (X >= 1 && X - 1 < G) store_split4(hi, lo, ((row * G + X - 1) * 16 + (col - 8 * (X - 1))) * 4, r4);
}

__global__ void reflect_pad_c4_kernel(const float* __restrict__ x, float* __restrict__ out, int N,
                                      int H, int W, int pad, __nv_bfloat16* __restrict__ x16hi,
                                      __nv_bfloat16* __restrict__ x16lo, int x8) {
    FS_PDL_ENTER();
    int OH = H + 2 * pad, OW = W + 2 * pad;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * OH * OW) return;
    int ox = (int)(i % OW);
    long long r = i / OW;
    int oy = (int)(r % OH);
    int n = (int)(r / OH);
    int sy = oy - pad, sx = ox - pad;
    if (sy < 0) sy = -sy;
    if (sy >= H) sy = 2 * (H - 1) - sy;
    if (sx < 0) sx = -sx;
    if (sx >= W) sx = 2 * (W - 1) - sx;
    const float* s = x + (((long long)n * H + sy) * W + sx) * 3;
    st4(out + i * 4, make_float4(s[0], s[1], s[2], 0.f));
    if (x16hi) {
        const float r4[4] = {s[0], s[1], s[2], 0.f};
        if (x8) store_x8(x16hi, x16lo, r, ox + 4, OW >> 3, r4);
        else store_split4(x16hi, x16lo, ((r * (OW + 16)) + ox + 4) * 4, r4);
    }
}

__global__ void vgg_preprocess_kernel(const float* __restrict__ x, float* __restrict__ out,
                                      long long npix) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float* s = x + i * 3;
    st4(out + i * 4, make_float4(s[0] - 123.68f, s[1] - 116.779f, s[2] - 103.939f, 0.f));
}

// ------------------------------------------------------------------ uint8 frames (streaming path)
__global__ void u8_to_f32_kernel(const uchar4* __restrict__ in, float4* __restrict__ out, long long n4,
                                 const unsigned char* __restrict__ in1, float* __restrict__ out1, long long n) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        uchar4 v = in[i];
        out[i] = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
    }
    long long tail = n4 * 4 + i;                      // < 4 leftover bytes, handled by the first threads
    if (i < 4 && tail < n) out1[tail] = (float)in1[tail];
}

// numpy astype(uint8) of a value in [0,255] truncates (stylize_webcam.py:89); swap_rb = cv2.COLOR_BGR2RGB (:90).
// mode bit 0: swap channels 0/2; bit 1: round half to even first (cv2.imwrite's saturate_cast of a float image,
// utils.py:51-52 / stylize_image.py:79-80) instead of truncating.
__global__ void f32_to_u8_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, long long npix,
                                 int mode) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const int swap_rb = mode & 1;
    const float* s = in + i * 3;
    float a = s[0], b = s[1], c = s[2];
    if (mode & 2) { a = rintf(a); b = rintf(b); c = rintf(c); }
    unsigned char* o = out + i * 3;
    unsigned char ua = (unsigned char)fminf(fmaxf(a, 0.f), 255.f);
    unsigned char ub = (unsigned char)fminf(fmaxf(b, 0.f), 255.f);
    unsigned char uc = (unsigned char)fminf(fmaxf(c, 0.f), 255.f);
    o[0] = swap_rb ? uc : ua; o[1] = ub; o[2] = swap_rb ? ua : uc;
}

// ------------------------------------------------------------------ input pipeline: TF-1.0 bicubic resize
// tf.image.resize_images(method=2) of TF 1.0 (reference datapipe.py:25): align_corners = False, no half-pixel
// centres (pos = out * in/out), A = -0.75, coefficients from a 1024-entry table indexed by lrintf(delta * 1024)
// (table entries evaluated in double and stored as float), border taps clamped.  Arithmetic order and roundings
// follow TF's kernel: per 4x4 patch four horizontal interpolations, then one vertical; every operation
// individually rounded (no FMA contraction).
__device__ __forceinline__ void bicubic_taps(int o, float scale, int in_size, int* idx, float* w) {
    const float pos = __fmul_rn(scale, (float)o);
    const float fl = floorf(pos);
    const int loc = (int)fl;
    const float delta = __fsub_rn(pos, fl);
    const int off = (int)rintf(__fmul_rn(delta, 1024.f));
    const double A = -0.75;
    auto t0 = [&](int k) {          // table[2k]:   ((A+2)x - (A+3)) x x + 1
        const double x = (double)((float)k / 1024.f);
        return (float)(((A + 2.0) * x - (A + 3.0)) * x * x + 1.0);
    };
    auto t1 = [&](int k) {          // table[2k+1]: ((A x1 - 5A) x1 + 8A) x1 - 4A,  x1 = x + 1
        const double x1 = (double)((float)k / 1024.f) + 1.0;
        return (float)(((A * x1 - 5.0 * A) * x1 + 8.0 * A) * x1 - 4.0 * A);
    };
    w[0] = t1(off); w[1] = t0(off); w[2] = t0(1024 - off); w[3] = t1(1024 - off);
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = min(max(loc - 1 + k, 0), in_size - 1);
}

__global__ void resize_bicubic_tf1_kernel(const unsigned char* __restrict__ in, float* __restrict__ out, int H, int W,
                                          int OH, int OW, float sy, float sx) {
    FS_PDL_ENTER();
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    if (ox >= OW) return;
    int iy[4], ix[4];
    float wy[4], wx[4];
    bicubic_taps(oy, sy, H, iy, wy);
    bicubic_taps(ox, sx, W, ix, wx);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r = 0.f;           // coeff[i] = sum_k img[iy[i], ix[k], c] * wx[k], k in order
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float v = (float)in[((long long)iy[i] * W + ix[k]) * 3 + c];
                const float pr = __fmul_rn(v, wx[k]);
                r = k == 0 ? pr : __fadd_rn(r, pr);
            }
            const float pr = __fmul_rn(r, wy[i]);
            acc = i == 0 ? pr : __fadd_rn(acc, pr);
        }
        out[((long long)oy * OW + ox) * 3 + c] = acc;
    }
}

// ------------------------------------------------------------------ InstanceNorm
// MODE 0: sums of (x, x^2).   MODE 1: sums of (dz, dz*xhat) for the backward pass.
// Reduces pixels [beg, end) of sample n into the CTA's shared accumulators sm[C][2] (valid after the call).
template <int MODE>
__device__ __forceinline__ void in_reduce_chunk(const float* __restrict__ x, const float* __restrict__ dY,
                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                double* sm, int n, int beg, int end, int HW, int C, int act) {
    const int t = threadIdx.x;
    const int CL = C >> 2, rows = 256 / CL;
    const int lane = t % CL, row = t / CL;
    for (int i = t; i < 2 * C; i += 256) sm[i] = 0.0;
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    float mu[4], rs[4], g[4], b[4];
    if (MODE == 1) {
        float4 v;
        v = ld4(mean + (long long)n * C + lane * 4); mu[0] = v.x; mu[1] = v.y; mu[2] = v.z; mu[3] = v.w;
        v = ld4(rstd + (long long)n * C + lane * 4); rs[0] = v.x; rs[1] = v.y; rs[2] = v.z; rs[3] = v.w;
        v = ld4(scale + lane * 4); g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
        v = ld4(shift + lane * 4); b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
    }
    const float* xn = x + (long long)n * HW * C + lane * 4;
    const float* dn = MODE == 1 ? dY + (long long)n * HW * C + lane * 4 : nullptr;
    if (row < rows) {
        // 4 pixels per trip, all loads issued before the (fp64) accumulation: the planes are L2-resident and the
        // loop is latency-bound otherwise.  Out-of-range slots contribute exact zeros.
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = beg + row; p < end; p += 4 * rows) {
            float4 xq[4], dq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int pu = p + u * rows;
                xq[u] = pu < end ? ld4(xn + (long long)pu * C) : z4;
                if (MODE == 1) dq[u] = pu < end ? ld4(dn + (long long)pu * C) : z4;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float xv[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w};
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        double d = (double)xv[j];
                        s1[j] += d;
                        s2[j] += d * d;
                    }
                } else {
                    float dy[4] = {dq[u].x, dq[u].y, dq[u].z, dq[u].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float xh = (xv[j] - mu[j]) * rs[j];
                        float dz = dy[j];
                        if (act == ACT_RELU) {
                            dz = __fmaf_rn(xh, g[j], b[j]) > 0.f ? dz : 0.f;
                        } else if (act == ACT_TANH255) {
                            float th = tanhf(__fmaf_rn(xh, g[j], b[j]));
                            dz = dz * 127.5f * (1.f - th * th);
                        }
                        s1[j] += (double)dz;
                        s2[j] += (double)dz * (double)xh;
                    }
                }
            }
        }
    }
    __syncthreads();
    // lanes t, t+CL, t+2CL, ... of a warp hold the same channels: fold them with shuffles first so that only
    // CL lanes per warp touch the shared accumulators (CL <= 32; for CL = 64 every lane is distinct)
    if (CL < 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            for (int o = 16; o >= CL; o >>= 1) {
                s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
                s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
            }
        }
    }
    if (CL >= 32 || (t & 31) < CL) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(&sm[(lane * 4 + j) * 2 + 0], s1[j]);
            atomicAdd(&sm[(lane * 4 + j) * 2 + 1], s2[j]);
        }
    }
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(256)
in_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dY,
                 const float* __restrict__ mean, const float* __restrict__ rstd,
                 const float* __restrict__ scale, const float* __restrict__ shift,
                 double* __restrict__ partial, int HW, int C, int chunks, int act) {
    FS_PDL_ENTER();
    extern __shared__ double sm[];            // [C][2]
    const int t = threadIdx.x, n = blockIdx.y, chunk = blockIdx.x;
    const int per = (HW + chunks - 1) / chunks;
    const int beg = chunk * per, end = min(beg + per, HW);
    in_reduce_chunk<MODE>(x, dY, mean, rstd, scale, shift, sm, n, beg, end, HW, C, act);
    double* dst = partial + ((long long)n * chunks + chunk) * 2 * C;
    for (int i = t; i < 2 * C; i += 256) dst[i] = sm[i];
}

// one warp per (n, c): lanes split the chunks, fixed shuffle-tree order -> deterministic
__global__ void __launch_bounds__(256)
in_stats_finalize_kernel(const double* __restrict__ partial, float* __restrict__ mean, float* __restrict__ rstd,
                         int N, int C, int chunks, int HW, float eps) {
    FS_PDL_ENTER();
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N * C) return;
    const int n = i / C, c = i - n * C;
    double s1 = 0, s2 = 0;
    for (int k = lane; k < chunks; k += 32) {
        const double* p = partial + ((long long)n * chunks + k) * 2 * C + c * 2;
        s1 += p[0];
        s2 += p[1];
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
        double m = s1 / HW;
        double var = s2 / HW - m * m;
        if (var < 0) var = 0;
        mean[i] = (float)m;
        rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// mean / rstd from the (sum, sum of squares) a conv epilogue accumulated in `reps` replicas [reps][NC][2];
// zeroes the sums for the next launch
__global__ void in_stats_from_sums_kernel(double* __restrict__ sums, float* __restrict__ mean, float* __restrict__ rstd,
                                          int NC, int HW, float eps, int reps) {
    FS_PDL_ENTER();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NC) return;
    double s1 = 0.0, s2 = 0.0;
    for (int r = 0; r < reps; ++r) {
        double* p = sums + ((long long)r * NC + i) * 2;
        s1 += p[0]; s2 += p[1];
        p[0] = 0.0; p[1] = 0.0;
    }
    const double m = s1 / HW;
    double var = s2 / HW - m * m;
    if (var < 0) var = 0;
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void in_bwd_finalize_kernel(const double* __restrict__ partial, float* __restrict__ m12,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int N,
                                       int C, int chunks, int HW) {
    FS_PDL_ENTER();
    // one block per channel; warp w handles samples w, w+8, ...; lanes split the chunks.
    // Fixed reduction order (lane tree, then warps 0..7 in order) -> deterministic.
    __shared__ double wsum[8][2];
    const int c = blockIdx.x;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double g = 0, b = 0;
    for (int n = w; n < N; n += 8) {
        double s1 = 0, s2 = 0;
        for (int k = lane; k < chunks; k += 32) {
            const double* p = partial + ((long long)n * chunks + k) * 2 * C + c * 2;
            s1 += p[0];
            s2 += p[1];
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
            m12[((long long)n * C + c) * 2 + 0] = (float)(s1 / HW);
            m12[((long long)n * C + c) * 2 + 1] = (float)(s2 / HW);
        }
        b += s1;
        g += s2;
    }
    if (lane == 0) { wsum[w][0] = b; wsum[w][1] = g; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double bb = 0, gg = 0;
        for (int i = 0; i < 8; ++i) { bb += wsum[i][0]; gg += wsum[i][1]; }
        if (dgamma) dgamma[c] = (float)gg;
        if (dbeta) dbeta[c] = (float)bb;
    }
}

// one float4 (4 channels of one pixel) of the InstanceNorm forward: affine + activation (+ cropped skip) + stores
__device__ __forceinline__ const float* in_skip_ptr(const float* skip, int n, int p, int c, int H, int W, int C) {
    int y = p / W, xx = p - y * W;
    return skip + (((long long)n * (H + 4) + y + 2) * (W + 4) + xx + 2) * C + c;
}
// v = the raw value, s = the (cropped) skip value (ignored unless has_skip)
__device__ __forceinline__ void in_apply_elem(const float4 v, const float4 s, bool has_skip, const float4 mu, const float4 rs,
                                              const float4 g, const float4 b, float* __restrict__ out,
                                              long long i, int n, int p, int H, int W, int act, int out3,
                                              __nv_bfloat16* __restrict__ shi, __nv_bfloat16* __restrict__ slo) {
    float r[4] = {in_affine(v.x, mu.x, rs.x, g.x, b.x), in_affine(v.y, mu.y, rs.y, g.y, b.y),
                  in_affine(v.z, mu.z, rs.z, g.z, b.z), in_affine(v.w, mu.w, rs.w, g.w, b.w)};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (act == ACT_RELU) r[j] = fmaxf(r[j], 0.f);
        else if (act == ACT_TANH255) r[j] = __fmaf_rn(127.5f, tanhf(r[j]), 127.5f);
    }
    if (has_skip) { r[0] += s.x; r[1] += s.y; r[2] += s.z; r[3] += s.w; }
    if (out3) {
        float* o = out + ((long long)n * H * W + p) * 3;
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
    } else {
        if (out) st4(out + i * 4, make_float4(r[0], r[1], r[2], r[3]));      // null: only the split planes are consumed
        if (shi) store_split4(shi, slo, i * 4, r);
    }
}

__global__ void in_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float* __restrict__ skip,
                                float* __restrict__ out, int N, int H, int W, int C, int act, int out3,
                                __nv_bfloat16* __restrict__ shi, __nv_bfloat16* __restrict__ slo) {
    FS_PDL_ENTER();
    const int C4 = C >> 2;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * H * W * C4;
    if (i >= total) return;
    int c = (int)(i % C4) * 4;
    long long pix = i / C4;
    int HW = H * W;
    int n = (int)(pix / HW);
    int p = (int)(pix - (long long)n * HW);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    in_apply_elem(ld4(x + i * 4), skip ? ld4(in_skip_ptr(skip, n, p, c, H, W, C)) : z4, skip != nullptr,
                  ld4(mean + (long long)n * C + c), ld4(rstd + (long long)n * C + c), ld4(scale + c), ld4(shift + c),
                  out, i, n, p, H, W, act, out3, shi, slo);
}

// (sum0, sum1) of channel t of sample n over the `reps` replicas [reps][N*C][2], for threads t < C of a 256-thread CTA.
// The 256 threads split the replicas (256 / C' groups, C' = C rounded up to a power of two <= 128) so that every
// thread's loads are independent and issued together: one L2 round trip instead of `reps` dependent ones.
__device__ __forceinline__ double2 sum_replicas(const double* __restrict__ sums, int reps, int N, int n, int C,
                                               double2* s_part) {
    const int t = threadIdx.x;
    int Cp = 4;
    while (Cp < C) Cp <<= 1;
    const int groups = 256 / Cp, c = t % Cp, grp = t / Cp;
    double2 acc = make_double2(0.0, 0.0);
    if (c < C) {
#pragma unroll 4
        for (int r = grp; r < reps; r += groups) {
            const double2 v = *reinterpret_cast<const double2*>(sums + (((long long)r * N + n) * C + c) * 2);
            acc.x += v.x; acc.y += v.y;
        }
    }
    s_part[t] = acc;
    __syncthreads();
    double2 tot = make_double2(0.0, 0.0);
    if (t < C)
        for (int g = 0; g < groups; ++g) { tot.x += s_part[g * Cp + t].x; tot.y += s_part[g * Cp + t].y; }
    __syncthreads();
    return tot;
}

// Fused form of stats-finalize + apply for the layers whose conv epilogue accumulated (sum, sum of squares) in
// `reps` replicas [reps][N*C][2]: every CTA (chunk, n) derives its sample's mean / rstd itself (C x reps fp64 loads,
// the same arithmetic as in_stats_from_sums_kernel), chunk 0 publishes them for the backward pass, and the CTAs
// jointly zero `zero_buf` - the OTHER ping-pong set of accumulators, last read one layer ago - for the next conv.
__global__ void __launch_bounds__(256)
in_apply_sums_kernel(const float* __restrict__ x, const double* __restrict__ sums, int reps, float eps,
                     float* __restrict__ mean, float* __restrict__ rstd, double* __restrict__ zero_buf, long long zero_n,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ skip,
                     float* __restrict__ out, int N, int H, int W, int C, int act, int out3,
                     __nv_bfloat16* __restrict__ shi, __nv_bfloat16* __restrict__ slo, int chunks) {
    FS_PDL_ENTER();
    __shared__ __align__(16) float s_mu[128], s_rs[128];
    __shared__ double2 s_part[256];
    const int n = blockIdx.y, chunk = blockIdx.x, t = threadIdx.x, HW = H * W;
    {
        const double2 s = sum_replicas(sums, reps, N, n, C, s_part);      // valid for t < C
        if (t < C) {
            const double m = s.x / HW;
            double var = s.y / HW - m * m;
            if (var < 0) var = 0;
            const float mf = (float)m, rf = (float)(1.0 / sqrt(var + (double)eps));
            s_mu[t] = mf; s_rs[t] = rf;
            if (chunk == 0) { mean[(long long)n * C + t] = mf; rstd[(long long)n * C + t] = rf; }
        }
    }
    if (zero_buf) {
        const long long cta = (long long)blockIdx.y * gridDim.x + blockIdx.x, nct = (long long)gridDim.x * gridDim.y;
        const long long per = (zero_n + nct - 1) / nct, zb = cta * per, ze = zb + per < zero_n ? zb + per : zero_n;
        for (long long i = zb + t; i < ze; i += 256) zero_buf[i] = 0.0;
    }
    __syncthreads();
    const int C4 = C >> 2, lane = t % C4, rows = 256 / C4, row = t / C4, c = lane * 4;
    const float4 mu = ld4(s_mu + c), rs = ld4(s_rs + c), g = ld4(scale + c), b = ld4(shift + c);
    const int per = (HW + chunks - 1) / chunks, beg = chunk * per, end = min(beg + per, HW);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = beg + row; p < end; p += rows) {
        const long long i = ((long long)n * HW + p) * C4 + lane;
        in_apply_elem(ld4(x + i * 4), skip ? ld4(in_skip_ptr(skip, n, p, c, H, W, C)) : z4, skip != nullptr,
                      mu, rs, g, b, out, i, n, p, H, W, act, out3, shi, slo);
    }
}

// one float4 of the InstanceNorm backward: dx = g*rstd*(dz - mean(dz) - xhat*mean(dz*xhat))
__device__ __forceinline__ void in_bwd_elem(const float4 xv4, const float4 dv4, const float* mu,
                                            const float* rs, const float* g, const float* b, const float* m1,
                                            const float* m2, float* __restrict__ dx, long long i, int act,
                                            __nv_bfloat16* __restrict__ shi, __nv_bfloat16* __restrict__ slo) {
    float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w}, dy[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
    float r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float xh = (xv[j] - mu[j]) * rs[j];
        float dz = dy[j];
        if (act == ACT_RELU) {
            dz = __fmaf_rn(xh, g[j], b[j]) > 0.f ? dz : 0.f;
        } else if (act == ACT_TANH255) {
            float th = tanhf(__fmaf_rn(xh, g[j], b[j]));
            dz = dz * 127.5f * (1.f - th * th);
        }
        r[j] = g[j] * rs[j] * (dz - m1[j] - xh * m2[j]);
    }
    if (dx) st4(dx + i * 4, make_float4(r[0], r[1], r[2], r[3]));            // null: only the split planes are consumed
    if (shi) store_split4(shi, slo, i * 4, r);
}

__global__ void in_bwd_apply_kernel(const float* __restrict__ dY, const float* __restrict__ x,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ m12, float* __restrict__ dx, int N,
                                    int HW, int C, int act, __nv_bfloat16* __restrict__ shi,
                                    __nv_bfloat16* __restrict__ slo) {
    FS_PDL_ENTER();
    const int C4 = C >> 2;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * HW * C4;
    if (i >= total) return;
    int c = (int)(i % C4) * 4;
    long long pix = i / C4;
    int n = (int)(pix / HW);
    float4 mu4 = ld4(mean + (long long)n * C + c), rs4 = ld4(rstd + (long long)n * C + c);
    float4 g4 = ld4(scale + c), b4 = ld4(shift + c);
    const float* mm = m12 + ((long long)n * C + c) * 2;
    float4 ma = ld4(mm), mb = ld4(mm + 4);
    float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
    float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    float m1[4] = {ma.x, ma.z, mb.x, mb.z}, m2[4] = {ma.y, ma.w, mb.y, mb.w};
    in_bwd_elem(ld4(x + i * 4), ld4(dY + i * 4), mu, rs, g, b, m1, m2, dx, i, act, shi, slo);
}

// Two-launch InstanceNorm backward.  in_bwd_reduce_kernel: per-CTA (sum dz, sum dz*xhat) added to the replicated fp64
// accumulators bs[rep][N*C][2] (rep = chunk % reps; one atomic per CTA, channel and quantity).  in_bwd_apply_sums_kernel:
// every CTA (chunk, n) derives m1 / m2 of its sample from the replicas, applies dx, zeroes the other ping-pong set for the
// next layer; the first ceil(C/8) CTAs also reduce dgamma / dbeta over (replica, sample) - a warp per channel.
__global__ void __launch_bounds__(256)
in_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dY, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                     double* __restrict__ bs, int reps, int N, int HW, int C, int chunks, int act) {
    FS_PDL_ENTER();
    extern __shared__ double sm[];            // [C][2]
    const int t = threadIdx.x, n = blockIdx.y, chunk = blockIdx.x;
    const int per = (HW + chunks - 1) / chunks;
    const int beg = chunk * per, end = min(beg + per, HW);
    in_reduce_chunk<1>(x, dY, mean, rstd, scale, shift, sm, n, beg, end, HW, C, act);
    double* dst = bs + (((long long)(chunk % reps) * N + n) * C) * 2;
    for (int i = t; i < 2 * C; i += 256) atomicAdd(dst + i, sm[i]);
}

__global__ void __launch_bounds__(256)
in_bwd_apply_sums_kernel(const float* __restrict__ dY, const float* __restrict__ x, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ shift,
                         const double* __restrict__ bs, int reps, double* __restrict__ zero_buf, long long zero_n,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx, int N, int HW,
                         int C, int act, __nv_bfloat16* __restrict__ shi, __nv_bfloat16* __restrict__ slo, int chunks) {
    FS_PDL_ENTER();
    __shared__ __align__(16) float s_m1[128], s_m2[128];
    __shared__ double2 s_part[256];
    const int n = blockIdx.y, chunk = blockIdx.x, t = threadIdx.x;
    {
        const double2 s = sum_replicas(bs, reps, N, n, C, s_part);        // valid for t < C
        if (t < C) { s_m1[t] = (float)(s.x / HW); s_m2[t] = (float)(s.y / HW); }
    }
    const long long cta = (long long)blockIdx.y * gridDim.x + blockIdx.x, nct = (long long)gridDim.x * gridDim.y;
    if (zero_buf) {
        const long long per = (zero_n + nct - 1) / nct, zb = cta * per, ze = zb + per < zero_n ? zb + per : zero_n;
        for (long long i = zb + t; i < ze; i += 256) zero_buf[i] = 0.0;
    }
    // dgamma[c] = sum over (replica, sample) of sum dz*xhat, dbeta[c] likewise of sum dz: CTA j takes channels 8j..8j+7
    for (long long j = cta; j * 8 < C; j += nct) {
        const int cc = (int)j * 8 + (t >> 5), lane = t & 31;
        double gb = 0.0, gg = 0.0;
        if (cc < C)
            for (int k = lane; k < reps * N; k += 32) {
                const double* p = bs + ((long long)k * C + cc) * 2;
                gb += p[0]; gg += p[1];
            }
        gb = warp_sum(gb); gg = warp_sum(gg);
        if (lane == 0 && cc < C) {
            if (dgamma) dgamma[cc] = (float)gg;
            if (dbeta) dbeta[cc] = (float)gb;
        }
    }
    __syncthreads();
    const int C4 = C >> 2, lane = t % C4, rows = 256 / C4, row = t / C4, c = lane * 4;
    const float4 mu4 = ld4(mean + (long long)n * C + c), rs4 = ld4(rstd + (long long)n * C + c);
    const float4 g4 = ld4(scale + c), b4 = ld4(shift + c), ma = ld4(s_m1 + c), mb = ld4(s_m2 + c);
    const float mu[4] = {mu4.x, mu4.y, mu4.z, mu4.w}, rs[4] = {rs4.x, rs4.y, rs4.z, rs4.w};
    const float g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
    const float m1[4] = {ma.x, ma.y, ma.z, ma.w}, m2[4] = {mb.x, mb.y, mb.z, mb.w};
    const int per = (HW + chunks - 1) / chunks, beg = chunk * per, end = min(beg + per, HW);
#pragma unroll 4
    for (int p = beg + row; p < end; p += rows) {
        const long long i = ((long long)n * HW + p) * C4 + lane;
        in_bwd_elem(ld4(x + i * 4), ld4(dY + i * 4), mu, rs, g, b, m1, m2, dx, i, act, shi, slo);
    }
}

// ------------------------------------------------------------------ pooling
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int H,
                                   int W, int C, __nv_bfloat16* __restrict__ shi,
                                   __nv_bfloat16* __restrict__ slo) {
    FS_PDL_ENTER();
    const int C4 = C >> 2, PH = (H + 1) >> 1, PW = (W + 1) >> 1;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * PH * PW * C4;
    if (i >= total) return;
    int c = (int)(i % C4) * 4;
    long long pix = i / C4;
    int px = (int)(pix % PW);
    long long r = pix / PW;
    int py = (int)(r % PH);
    int n = (int)(r / PH);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            int y = 2 * py + dy, xx = 2 * px + dx;
            if (y < H && xx < W) {
                float4 v = ld4(x + (((long long)n * H + y) * W + xx) * C + c);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y);
                m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
    st4(out + i * 4, m);
    if (shi) {
        float r[4] = {m.x, m.y, m.z, m.w};
        store_split4(shi, slo, i * 4, r);
    }
}

// One thread per 2x2 pooling window and 4 channels: the four activations are read once (the previous
// per-output-pixel form re-read every window four times).
__global__ void pool_bwd_combine_kernel(const float* __restrict__ act, const float* __restrict__ gpool,
                                        const float* __restrict__ ctarget, float cw2, int apply_mask,
                                        float* __restrict__ out, int N, int H, int W, int C) {
    FS_PDL_ENTER();
    const int C4 = C >> 2, PH = (H + 1) >> 1, PW = (W + 1) >> 1;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * PH * PW * C4;
    if (i >= total) return;
    int c = (int)(i % C4) * 4;
    long long pp = i / C4;
    int px = (int)(pp % PW);
    long long r = pp / PW;
    int py = (int)(r % PH);
    int n = (int)(r / PH);
    float v[4][4];
    bool in[4];
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    int arg[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int yy = 2 * py + (k >> 1), xx = 2 * px + (k & 1);
        in[k] = yy < H && xx < W;
        if (in[k]) {
            float4 v4 = ld4(act + (((long long)n * H + yy) * W + xx) * C + c);
            v[k][0] = v4.x; v[k][1] = v4.y; v[k][2] = v4.z; v[k][3] = v4.w;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (v[k][j] > best[j]) { best[j] = v[k][j]; arg[j] = k; }       // first maximum wins (scan order)
        }
    }
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (gpool) {
        float4 g4 = ld4(gpool + i * 4);
        g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!in[k]) continue;
        int yy = 2 * py + (k >> 1), xx = 2 * px + (k & 1);
        long long o = (((long long)n * H + yy) * W + xx) * C + c;
        float res[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) res[j] = (gpool && arg[j] == k) ? g[j] : 0.f;
        if (ctarget) {
            float4 t4 = ld4(ctarget + o);
            res[0] += cw2 * (v[k][0] - t4.x); res[1] += cw2 * (v[k][1] - t4.y);
            res[2] += cw2 * (v[k][2] - t4.z); res[3] += cw2 * (v[k][3] - t4.w);
        }
        if (apply_mask) {
#pragma unroll
            for (int j = 0; j < 4; ++j) res[j] = v[k][j] > 0.f ? res[j] : 0.f;
        }
        st4(out + o, make_float4(res[0], res[1], res[2], res[3]));
    }
}

// ------------------------------------------------------------------ losses
__global__ void sqdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n4,
                                  double scale, double* acc) {
    FS_PDL_ENTER();
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 x = ld4(a + i * 4), y = ld4(b + i * 4);
        float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
        s += (double)d0 * d0 + (double)d1 * d1 + (double)d2 * d2 + (double)d3 * d3;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) atomicAdd(acc, s * scale);
}

__global__ void style_loss_grad_kernel(const float* __restrict__ G, const float* __restrict__ T,
                                       float* __restrict__ S, long long total, int CC, float coef,
                                       double lscale, double* acc) {
    FS_PDL_ENTER();
    double s = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        float d = G[i] - T[i % CC];
        if (S) S[i] = coef * d;
        s += (double)d * d;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) atomicAdd(acc, s * lscale);
}

__global__ void tv_kernel(const float* __restrict__ Y, float* __restrict__ dY4, int N, int H, int W,
                          float beta, double* acc) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)N * H * W;
    double s = 0;
    if (i < total) {
        int x = (int)(i % W);
        long long r = i / W;
        int y = (int)(r % H);
        const float* p = Y + i * 3;
        float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = p[c];
            if (y + 1 < H) { float d = v - p[c + (long long)W * 3]; s += (double)d * d; g[c] += 2.f * d; }
            if (y > 0) { float d = p[c - (long long)W * 3] - v; g[c] -= 2.f * d; }
            if (x + 1 < W) { float d = v - p[c + 3]; s += (double)d * d; g[c] += 2.f * d; }
            if (x > 0) { float d = p[c - 3] - v; g[c] -= 2.f * d; }
        }
        if (dY4) {
            float4 o = ld4(dY4 + i * 4);
            o.x += beta * g[0]; o.y += beta * g[1]; o.z += beta * g[2];
            st4(dY4 + i * 4, o);
        }
    }
    s = block_sum(s);
    if (threadIdx.x == 0 && s != 0.0) atomicAdd(acc, s * (double)beta);
}

__global__ void finalize_losses_kernel(const double* acc, float* out4) {
    FS_PDL_ENTER();
    out4[0] = (float)acc[0];
    out4[1] = (float)acc[1];
    out4[2] = (float)acc[2];
    out4[3] = (float)(acc[0] + acc[1] + acc[2]);
}

// ------------------------------------------------------------------ Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                            const int* __restrict__ step) {
    FS_PDL_ENTER();
    __shared__ float lr_t_s;
    if (threadIdx.x == 0) {
        int t = *step + 1;
        double c = (double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t));
        lr_t_s = (float)c;
    }
    __syncthreads();
    const float lr_t = lr_t_s;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i];
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}
__global__ void incr_kernel(int* c) {
    FS_PDL_ENTER(); *c += 1; }

// ------------------------------------------------------------------ weight layouts
__global__ void pad_taps_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int Ci,
                                int Co, int Cip, int Cop, int unpad) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!unpad) {
        if (i >= (long long)T * Cip * Cop) return;
        int o = (int)(i % Cop);
        long long r = i / Cop;
        int ci = (int)(r % Cip);
        int t = (int)(r / Cip);
        dst[i] = (ci < Ci && o < Co) ? src[((long long)t * Ci + ci) * Co + o] : 0.f;
    } else {
        if (i >= (long long)T * Ci * Co) return;
        int o = (int)(i % Co);
        long long r = i / Co;
        int ci = (int)(r % Ci);
        int t = (int)(r / Ci);
        dst[i] = src[((long long)t * Cip + ci) * Cop + o];
    }
}

__global__ void transpose_taps_kernel(const float* __restrict__ src, float* __restrict__ dst, int T,
                                      int Ci, int Co) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)T * Ci * Co) return;
    int ci = (int)(i % Ci);
    long long r = i / Ci;
    int o = (int)(r % Co);
    int t = (int)(r / Co);
    dst[i] = src[((long long)t * Ci + ci) * Co + o];     // dst[t][o][ci]
}

// R_0 = [[1,1,1],[0,0,0]],  R_1 = [[1,1,0],[0,0,1]]   (rows a, cols kh)
__device__ __forceinline__ bool rsel(int p, int a, int k) {
    return p == 0 ? (a == 0) : (a == 0 ? k < 2 : k == 2);
}

__global__ void upconv_collapse_kernel(const float* __restrict__ W, float* __restrict__ Wc, int Ci,
                                       int Co) {
    FS_PDL_ENTER();
    // Wc[a][b][ci][(p*2+q)*Co + co]
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 4LL * Ci * 4 * Co;
    if (i >= total) return;
    int co = (int)(i % Co);
    long long r = i / Co;
    int pq = (int)(r % 4); r /= 4;
    int ci = (int)(r % Ci); r /= Ci;
    int b = (int)(r % 2), a = (int)(r / 2);
    int p = pq >> 1, q = pq & 1;
    float s = 0.f;
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
            if (rsel(p, a, kh) && rsel(q, b, kw)) s += W[(((long long)kh * 3 + kw) * Ci + ci) * Co + co];
    Wc[i] = s;
}

__global__ void upconv_collapse_grad_kernel(const float* __restrict__ dWc, float* __restrict__ dW,
                                            int Ci, int Co) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 9LL * Ci * Co;
    if (i >= total) return;
    int co = (int)(i % Co);
    long long r = i / Co;
    int ci = (int)(r % Ci); r /= Ci;
    int kw = (int)(r % 3), kh = (int)(r / 3);
    float s = 0.f;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int p = 0; p < 2; ++p)
                for (int q = 0; q < 2; ++q)
                    if (rsel(p, a, kh) && rsel(q, b, kw))
                        s += dWc[((((long long)a * 2 + b) * Ci + ci) * 4 + (p * 2 + q)) * Co + co];
    dW[i] = s;
}

// Stride-2 3x3 SAME(pad 0/1) data gradient as a 4-phase 2x2 gather over dy (sub-pixel form):
//   dx[2m+p, 2n+q, ci] = sum_{a,b,co} dy[m-a, n-b, co] * Wd[a][b][co][(p*2+q)*Ci + ci]
// with kh(p,a): p=0 -> {0,2}, p=1 -> {1,-};   2.25x fewer MACs than gathering all 9 taps.
__global__ void s2_dgrad_collapse_kernel(const float* __restrict__ W, float* __restrict__ Wd, int Ci, int Co) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 16LL * Ci * Co;
    if (i >= total) return;
    int ci = (int)(i % Ci);
    long long r = i / Ci;
    int pq = (int)(r % 4); r /= 4;
    int co = (int)(r % Co); r /= Co;
    int b = (int)(r % 2), a = (int)(r / 2);
    int p = pq >> 1, q = pq & 1;
    int kh = p == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : -1);
    int kw = q == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : -1);
    Wd[i] = (kh >= 0 && kw >= 0) ? W[(((long long)kh * 3 + kw) * Ci + ci) * Co + co] : 0.f;
}

// Stride-2 3x3 SAME(pad 0/1) FORWARD conv as a 2x2 stride-1 conv over the space-to-depth view of the input:
//   y[m, n, co] = sum_{a,b} sum_{p,q,ci} S[m+a, n+b, (p*2+q)*Ci + ci] * Wf[a][b][(p*2+q)*Ci + ci][co],
//   S[Y, X, (p,q,ci)] = x[2Y+p, 2X+q, ci],   Wf = W[2a+p][2b+q] (zero where 2a+p or 2b+q == 3).
__global__ void s2_fwd_collapse_kernel(const float* __restrict__ W, float* __restrict__ Wf, int Ci, int Co) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 16LL * Ci * Co;
    if (i >= total) return;
    int co = (int)(i % Co);
    long long r = i / Co;
    int ci = (int)(r % Ci); r /= Ci;
    int pq = (int)(r % 4); r /= 4;
    int b = (int)(r % 2), a = (int)(r / 2);
    int kh = 2 * a + (pq >> 1), kw = 2 * b + (pq & 1);
    Wf[i] = (kh < 3 && kw < 3) ? W[(((long long)kh * 3 + kw) * Ci + ci) * Co + co] : 0.f;
}

// Pixel pairing of a 2x2-tap GEMM whose channel sides are too narrow for a 64-channel K block: two horizontally
// adjacent pixels become one "pixel" with twice the channels on BOTH sides, W [2][2][K0][N0] -> Wp [2][2][2K0][2N0].
// Output pixel X = 2j+e reads input pixel X+b (correlation) or X-b (gather) = 2(j +/- kwp) + h, hence
//   Wp[a][kwp][kidx(h,k)][e*N0 + n] = W[a][b][k][n],  b = 2kwp + h - e (correlation) | e - h + 2kwp (gather), 0 if b not in {0,1}.
// kmode 0: plain input, kidx = h*K0 + k.  kmode 1: the input is a space-to-depth view, k = (p*2+q)*c0 + c, and the
// paired view keeps the row parity p outermost: kidx = p*K0 + h*(K0/2) + q*c0 + c (c0 = K0/4).
__global__ void pair_taps_kernel(const float* __restrict__ W, float* __restrict__ Wp, int K0, int N0, int kmode, int gather) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 16LL * K0 * N0;
    if (i >= total) return;
    int np = (int)(i % (2 * N0));
    long long r = i / (2 * N0);
    int kp = (int)(r % (2 * K0)); r /= 2 * K0;
    int kwp = (int)(r % 2), a = (int)(r / 2);
    int e = np / N0, n = np - e * N0;
    int h, k;
    if (kmode == 0) {
        h = kp / K0; k = kp - h * K0;
    } else {
        int c0 = K0 / 4;
        int p = kp / K0, rem = kp - p * K0;
        h = rem / (K0 / 2); rem -= h * (K0 / 2);
        int q = rem / c0, c = rem - q * c0;
        k = (p * 2 + q) * c0 + c;
    }
    int b = gather ? e - h + 2 * kwp : 2 * kwp + h - e;
    Wp[i] = (b == 0 || b == 1) ? W[(((long long)a * 2 + b) * K0 + k) * N0 + n] : 0.f;
}

// adjoint of pair_taps in its correlation form (gradients of paired weights back to the unpaired 2x2 weights):
//   dW[a][b][k][n] = sum over (kwp, h, e) with 2kwp + h - e == b of dWp[a][kwp][kidx(h,k)][e*N0 + n]
// n_s2d: the N side of dWp came from a paired space-to-depth view, whose channel order is (p; e, q, c) instead of
// the (e; p, q, c) order of pair_taps: column e*N0 + n with n = p*(N0/2) + r sits at p*N0 + e*(N0/2) + r.
__global__ void unpair_taps_kernel(const float* __restrict__ dWp, float* __restrict__ dW, int K0, int N0, int kmode,
                                   int n_s2d) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 4LL * K0 * N0;
    if (i >= total) return;
    int n = (int)(i % N0);
    long long r = i / N0;
    int k = (int)(r % K0); r /= K0;
    int b = (int)(r % 2), a = (int)(r / 2);
    float s = 0.f;
#pragma unroll
    for (int kwp = 0; kwp < 2; ++kwp)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int e = 2 * kwp + h - b;
            if (e < 0 || e > 1) continue;
            int kp;
            if (kmode == 0) kp = h * K0 + k;
            else {
                int c0 = K0 / 4, pq = k / c0, c = k - pq * c0;
                kp = (pq >> 1) * K0 + h * (K0 / 2) + (pq & 1) * c0 + c;
            }
            int col = e * N0 + n;
            if (n_s2d) { int p = n / (N0 / 2), r2 = n - p * (N0 / 2); col = p * N0 + e * (N0 / 2) + r2; }
            s += dWp[(((long long)a * 2 + kwp) * (2 * K0) + kp) * (2 * N0) + col];
        }
    dW[i] = s;
}

// adjoint of s2_fwd_collapse: dW[kh][kw][ci][co] = dWf[kh>>1][kw>>1][((kh&1)*2 + (kw&1))*Ci + ci][co]
__global__ void s2_fwd_collapse_grad_kernel(const float* __restrict__ dWf, float* __restrict__ dW, int Ci, int Co) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 9LL * Ci * Co;
    if (i >= total) return;
    int co = (int)(i % Co);
    long long r = i / Co;
    int ci = (int)(r % Ci); r /= Ci;
    int kw = (int)(r % 3), kh = (int)(r / 3);
    int a = kh >> 1, b = kw >> 1, pq = (kh & 1) * 2 + (kw & 1);
    dW[i] = dWf[((((long long)a * 2 + b) * 4 + pq) * Ci + ci) * Co + co];
}

// fp32 [N*H, W, C] -> split-bf16 planes [N*H, W + 16, C] with the row at column offset 4: the zero-margined input of
// the x16 space-to-depth 9x9 forms (the margins are zeroed once when the workspace is bound)
__global__ void split_pad_x16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo, long long n4, int W, int C, int x8) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long e = i * 4;
    const int c = (int)(e % C);
    const long long pix = e / C;
    const int xc = (int)(pix % W);
    const long long row = pix / W;
    const float4 v = ld4(x + e);
    float r[4] = {v.x, v.y, v.z, v.w};
    if (x8) store_x8(hi, lo, row, xc + 4, W >> 3, r);          // (C == 4: one pixel per thread)
    else store_split4(hi, lo, (row * (W + 16) + xc + 4) * C + c, r);
}

inline int grid1(long long n, int bs = 256) { return (int)((n + bs - 1) / bs); }

}  // namespace

// ==================================================================== launchers
int reflect_pad_c4(const float* x, float* out, int N, int H, int W, int pad, cudaStream_t st, void* x16hi, void* x16lo, int x8) {
    FS_CHECK(pad < H && pad < W, "reflect_pad: pad %d must be smaller than the image (%dx%d)", pad, H, W);
    long long n = (long long)N * (H + 2 * pad) * (W + 2 * pad);
    FS_CHECK(!x8 || (W + 2 * pad) % 8 == 0, "reflect_pad: the x8 planes need a padded width in whole 8-pixel groups");
    launch_k(reflect_pad_c4_kernel, dim3(grid1(n)), dim3(256), 0, st, x, out, N, H, W, pad, (__nv_bfloat16*)x16hi,
             (__nv_bfloat16*)x16lo, x8);
    FS_LAUNCH_CHECK();
    return 0;
}

int vgg_preprocess_c4(const float* x, float* out, long long npix, cudaStream_t st) {
    launch_k(vgg_preprocess_kernel, dim3(grid1(npix)), dim3(256), 0, st, x, out, npix);
    FS_LAUNCH_CHECK();
    return 0;
}

int frame_u8_to_f32(const unsigned char* in, float* out, long long n, cudaStream_t st) {
    FS_CHECK(((uintptr_t)in & 3) == 0 && ((uintptr_t)out & 15) == 0, "frame_u8_to_f32: misaligned buffer");
    long long n4 = n / 4;
    launch_k(u8_to_f32_kernel, dim3(grid1(n4 > 4 ? n4 : 4)), dim3(256), 0, st, reinterpret_cast<const uchar4*>(in),
             reinterpret_cast<float4*>(out), n4, in, out, n);
    FS_LAUNCH_CHECK();
    return 0;
}

int frame_f32_to_u8(const float* in, unsigned char* out, long long npix, int swap_rb, cudaStream_t st) {
    launch_k(f32_to_u8_kernel, dim3(grid1(npix)), dim3(256), 0, st, in, out, npix, swap_rb);
    FS_LAUNCH_CHECK();
    return 0;
}

int resize_bicubic_tf1_u8(const unsigned char* in, float* out, int H, int W, int OH, int OW, cudaStream_t st) {
    FS_CHECK(H >= 1 && W >= 1 && OH >= 1 && OW >= 1 && OH <= 65535, "resize_bicubic: bad dims %dx%d -> %dx%d", H, W, OH, OW);
    const float sy = (float)H / (float)OH, sx = (float)W / (float)OW;
    launch_k(resize_bicubic_tf1_kernel, dim3(grid1(OW, 128), OH), dim3(128), 0, st, in, out, H, W, OH, OW, sy, sx);
    FS_LAUNCH_CHECK();
    return 0;
}

int in_chunks(int N, int HW) {
    int c = (148 * 4 + N - 1) / N;
    int maxc = HW / 128;             // >= 128 pixels per chunk: enough CTAs to cover the SMs on small planes
    if (maxc < 1) maxc = 1;
    if (c > maxc) c = maxc;
    if (c > 64) c = 64;
    if (c < 1) c = 1;
    return c;
}

static int check_in_c(int C) {
    int CL = C / 4;
    FS_CHECK(C % 4 == 0 && CL >= 1 && CL <= 64 && 256 % CL == 0,
             "instnorm: unsupported channel count %d (need C/4 in {1,2,4,...,64})", C);
    return 0;
}

int instnorm_stats(const float* x, float* mean, float* rstd, int N, int HW, int C, float eps,
                   double* partial, cudaStream_t st) {
    FS_TRY(check_in_c(C));
    int chunks = in_chunks(N, HW);
    launch_k((in_reduce_kernel<0>), dim3(dim3(chunks, N)), dim3(256), 2 * C * sizeof(double), st, 
        x, nullptr, nullptr, nullptr, nullptr, nullptr, partial, HW, C, chunks, 0);
    FS_LAUNCH_CHECK();
    launch_k(in_stats_finalize_kernel, dim3(cdiv((long long)N * C, 8)), dim3(256), 0, st, partial, mean, rstd, N, C, chunks, HW, eps);
    FS_LAUNCH_CHECK();
    return 0;
}

int instnorm_stats_from_sums(double* sums, float* mean, float* rstd, int N, int HW, int C, float eps, int reps, cudaStream_t st) {
    launch_k(in_stats_from_sums_kernel, dim3(cdiv((long long)N * C, 128)), dim3(128), 0, st, sums, mean, rstd, N * C, HW, eps, reps);
    FS_LAUNCH_CHECK();
    return 0;
}

int instnorm_apply(const float* x, const float* mean, const float* rstd, const float* scale,
                   const float* shift, const float* skip, float* out, int N, int H, int W, int C,
                   int act, int out3, cudaStream_t st, void* split_hi, void* split_lo) {
    FS_CHECK(C % 4 == 0, "instnorm_apply: C%%4 != 0");
    FS_CHECK(!(split_hi && out3), "instnorm_apply: split output not available with out3");
    FS_CHECK(!out3 || C == 4, "instnorm_apply: out3 needs C==4");
    long long n = (long long)N * H * W * (C / 4);
    launch_k(in_apply_kernel, dim3(grid1(n)), dim3(256), 0, st, x, mean, rstd, scale, shift, skip, out, N, H, W, C, act, out3,
                                              (__nv_bfloat16*)split_hi, (__nv_bfloat16*)split_lo);
    FS_LAUNCH_CHECK();
    return 0;
}

int instnorm_bwd(const float* dY, const float* x, const float* mean, const float* rstd,
                 const float* scale, const float* shift, float* dx, float* dgamma, float* dbeta,
                 int N, int HW, int C, int act, double* partial, float* m12, cudaStream_t st,
                 void* split_hi, void* split_lo) {
    FS_TRY(check_in_c(C));
    int chunks = in_chunks(N, HW);
    launch_k((in_reduce_kernel<1>), dim3(dim3(chunks, N)), dim3(256), 2 * C * sizeof(double), st, 
        x, dY, mean, rstd, scale, shift, partial, HW, C, chunks, act);
    FS_LAUNCH_CHECK();
    launch_k(in_bwd_finalize_kernel, dim3(C), dim3(256), 0, st, partial, m12, dgamma, dbeta, N, C, chunks, HW);
    FS_LAUNCH_CHECK();
    long long n = (long long)N * HW * (C / 4);
    launch_k(in_bwd_apply_kernel, dim3(grid1(n)), dim3(256), 0, st, dY, x, mean, rstd, scale, shift, m12, dx, N, HW, C, act,
                                                  (__nv_bfloat16*)split_hi, (__nv_bfloat16*)split_lo);
    FS_LAUNCH_CHECK();
    return 0;
}

// grid of the fused apply kernels: about eight CTAs per SM (one resident wave), at least 64 pixels each
static int apply_chunks(int N, int HW) {
    int c = (num_sms() * 8 + N - 1) / N;
    int maxc = HW / 64;
    if (maxc < 1) maxc = 1;
    if (c > maxc) c = maxc;
    if (c < 1) c = 1;
    return c;
}

int instnorm_apply_from_sums(const float* x, const double* sums, int reps, float eps, float* mean, float* rstd,
                             double* zero_buf, long long zero_n, const float* scale, const float* shift,
                             const float* skip, float* out, int N, int H, int W, int C, int act, int out3,
                             cudaStream_t st, void* split_hi, void* split_lo) {
    FS_TRY(check_in_c(C));
    FS_CHECK(C <= 128, "instnorm_apply_from_sums: C <= 128");
    FS_CHECK(!(split_hi && out3), "instnorm_apply_from_sums: split output not available with out3");
    FS_CHECK(!out3 || C == 4, "instnorm_apply_from_sums: out3 needs C==4");
    const int chunks = apply_chunks(N, H * W);
    launch_k(in_apply_sums_kernel, dim3(chunks, N), dim3(256), 0, st, x, sums, reps, eps, mean, rstd, zero_buf, zero_n,
             scale, shift, skip, out, N, H, W, C, act, out3, (__nv_bfloat16*)split_hi, (__nv_bfloat16*)split_lo, chunks);
    FS_LAUNCH_CHECK();
    return 0;
}

int instnorm_bwd_sums(const float* dY, const float* x, const float* mean, const float* rstd, const float* scale,
                      const float* shift, float* dx, float* dgamma, float* dbeta, int N, int HW, int C, int act,
                      double* bs, int reps, double* zero_buf, long long zero_n, cudaStream_t st, void* split_hi,
                      void* split_lo) {
    FS_TRY(check_in_c(C));
    FS_CHECK(C <= 128, "instnorm_bwd_sums: C <= 128");
    const int rchunks = in_chunks(N, HW);          // (more, smaller CTAs measured slower: 0.506 -> 0.559 ms per step)
    launch_k(in_bwd_reduce_kernel, dim3(rchunks, N), dim3(256), 2 * C * sizeof(double), st, x, dY, mean, rstd, scale, shift,
             bs, reps, N, HW, C, rchunks, act);
    FS_LAUNCH_CHECK();
    const int chunks = apply_chunks(N, HW);
    launch_k(in_bwd_apply_sums_kernel, dim3(chunks, N), dim3(256), 0, st, dY, x, mean, rstd, scale, shift,
             (const double*)bs, reps, zero_buf, zero_n, dgamma, dbeta, dx, N, HW, C, act, (__nv_bfloat16*)split_hi,
             (__nv_bfloat16*)split_lo, chunks);
    FS_LAUNCH_CHECK();
    return 0;
}

int maxpool2x2_fwd(const float* x, float* out, int N, int H, int W, int C, cudaStream_t st,
                   void* split_hi, void* split_lo) {
    FS_CHECK(C % 4 == 0, "maxpool: C%%4 != 0");
    long long n = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    launch_k(maxpool_fwd_kernel, dim3(grid1(n)), dim3(256), 0, st, x, out, N, H, W, C, (__nv_bfloat16*)split_hi,
                                                 (__nv_bfloat16*)split_lo);
    FS_LAUNCH_CHECK();
    return 0;
}

int pool_bwd_combine(const float* act, const float* gpool, const float* ctarget, float cw2,
                     int apply_mask, float* out, int N, int H, int W, int C, cudaStream_t st) {
    FS_CHECK(C % 4 == 0, "pool_bwd: C%%4 != 0");
    long long n = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    launch_k(pool_bwd_combine_kernel, dim3(grid1(n)), dim3(256), 0, st, act, gpool, ctarget, cw2, apply_mask, out, N, H, W, C);
    FS_LAUNCH_CHECK();
    return 0;
}

int sqdiff_sum(const float* a, const float* b, long long n, double scale, double* acc, cudaStream_t st) {
    FS_CHECK(n % 4 == 0, "sqdiff_sum: n%%4 != 0");
    int blocks = grid1(n / 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(sqdiff_sum_kernel, dim3(blocks), dim3(256), 0, st, a, b, n / 4, scale, acc);
    FS_LAUNCH_CHECK();
    return 0;
}

int style_loss_grad(const float* G, const float* T, float* S, int N, int CC, float coef, double lscale,
                    double* acc, cudaStream_t st) {
    long long total = (long long)N * CC;
    int blocks = grid1(total);
    if (blocks > 148 * 8) blocks = 148 * 8;
    launch_k(style_loss_grad_kernel, dim3(blocks), dim3(256), 0, st, G, T, S, total, CC, coef, lscale, acc);
    FS_LAUNCH_CHECK();
    return 0;
}

int tv_loss_grad(const float* Y, float* dY4, int N, int H, int W, float beta, double* acc, cudaStream_t st) {
    long long n = (long long)N * H * W;
    launch_k(tv_kernel, dim3(grid1(n)), dim3(256), 0, st, Y, dY4, N, H, W, beta, acc);
    FS_LAUNCH_CHECK();
    return 0;
}

int finalize_losses(const double* acc, float* out4, cudaStream_t st) {
    launch_k(finalize_losses_kernel, dim3(1), dim3(1), 0, st, acc, out4);
    FS_LAUNCH_CHECK();
    return 0;
}

int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
              float eps, int* step_counter, cudaStream_t st) {
    launch_k(adam_kernel, dim3(grid1(n)), dim3(256), 0, st, p, g, m, v, n, lr, b1, b2, eps, step_counter);
    FS_LAUNCH_CHECK();
    launch_k(incr_kernel, dim3(1), dim3(1), 0, st, step_counter);
    FS_LAUNCH_CHECK();
    return 0;
}

int pad_taps(const float* src, float* dst, int T, int Ci, int Co, int Cip, int Cop, cudaStream_t st) {
    launch_k(pad_taps_kernel, dim3(grid1((long long)T * Cip * Cop)), dim3(256), 0, st, src, dst, T, Ci, Co, Cip, Cop, 0);
    FS_LAUNCH_CHECK();
    return 0;
}
int unpad_taps(const float* src, float* dst, int T, int Ci, int Co, int Cip, int Cop, cudaStream_t st) {
    launch_k(pad_taps_kernel, dim3(grid1((long long)T * Ci * Co)), dim3(256), 0, st, src, dst, T, Ci, Co, Cip, Cop, 1);
    FS_LAUNCH_CHECK();
    return 0;
}
int transpose_taps(const float* src, float* dst, int T, int Ci, int Co, cudaStream_t st) {
    launch_k(transpose_taps_kernel, dim3(grid1((long long)T * Ci * Co)), dim3(256), 0, st, src, dst, T, Ci, Co);
    FS_LAUNCH_CHECK();
    return 0;
}
int upconv_collapse(const float* W, float* Wc, int Ci, int Co, cudaStream_t st) {
    launch_k(upconv_collapse_kernel, dim3(grid1(16LL * Ci * Co)), dim3(256), 0, st, W, Wc, Ci, Co);
    FS_LAUNCH_CHECK();
    return 0;
}
int upconv_collapse_grad(const float* dWc, float* dW, int Ci, int Co, cudaStream_t st) {
    launch_k(upconv_collapse_grad_kernel, dim3(grid1(9LL * Ci * Co)), dim3(256), 0, st, dWc, dW, Ci, Co);
    FS_LAUNCH_CHECK();
    return 0;
}

int s2_dgrad_collapse(const float* W, float* Wd, int Ci, int Co, cudaStream_t st) {
    launch_k(s2_dgrad_collapse_kernel, dim3(grid1(16LL * Ci * Co)), dim3(256), 0, st, W, Wd, Ci, Co);
    FS_LAUNCH_CHECK();
    return 0;
}

int s2_fwd_collapse(const float* W, float* Wf, int Ci, int Co, cudaStream_t st) {
    launch_k(s2_fwd_collapse_kernel, dim3(grid1(16LL * Ci * Co)), dim3(256), 0, st, W, Wf, Ci, Co);
    FS_LAUNCH_CHECK();
    return 0;
}

int pair_taps(const float* W, float* Wp, int K0, int N0, int kmode, int gather, cudaStream_t st) {
    FS_CHECK(K0 % 4 == 0 && N0 >= 1, "pair_taps: bad channel counts");
    launch_k(pair_taps_kernel, dim3(grid1(16LL * K0 * N0)), dim3(256), 0, st, W, Wp, K0, N0, kmode, gather);
    FS_LAUNCH_CHECK();
    return 0;
}

int unpair_taps(const float* dWp, float* dW, int K0, int N0, int kmode, int n_s2d, cudaStream_t st) {
    launch_k(unpair_taps_kernel, dim3(grid1(4LL * K0 * N0)), dim3(256), 0, st, dWp, dW, K0, N0, kmode, n_s2d);
    FS_LAUNCH_CHECK();
    return 0;
}
int s2_fwd_collapse_grad(const float* dWf, float* dW, int Ci, int Co, cudaStream_t st) {
    launch_k(s2_fwd_collapse_grad_kernel, dim3(grid1(9LL * Ci * Co)), dim3(256), 0, st, dWf, dW, Ci, Co);
    FS_LAUNCH_CHECK();
    return 0;
}

int split_pad_x16(const float* x, void* hi, void* lo, long long rows, int W, int C, cudaStream_t st, int x8) {
    FS_CHECK(C % 4 == 0 && W >= 1 && rows >= 1, "split_pad_x16: bad dims");
    FS_CHECK(!x8 || (C == 4 && W % 8 == 0), "split_pad_x16: the x8 form needs 4 channels and W %% 8 == 0");
    const long long n4 = rows * W * C / 4;
    launch_k(split_pad_x16_kernel, dim3(grid1(n4)), dim3(256), 0, st, x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n4, W, C, x8);
    FS_LAUNCH_CHECK();
    return 0;
}

int fill_zero(void* p, size_t bytes, cudaStream_t st) {
    FS_CUDA(cudaMemsetAsync(p, 0, bytes, st));
    return 0;
}

}  // namespace fs
