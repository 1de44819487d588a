// Direct (shared-memory tiled) kernels for the two 9x9 stride-1 SAME convolutions of the
// transform net (initconv_0: 3(4)->16, upsample_2: 16->3(4); reference im_transf_net.py:38,70).
// Their GEMM view has N or K-per-tap of 4..16, which starves both the tensor pipe and the generic
// implicit-GEMM tiles (the im2col gather dominates).  Here a CTA stages a 32x32 pixel tile (+4 px
// halo) once in shared memory and every thread owns one (tap, 4x4 channel block) of the
// 81 x CI x CO weight-gradient, so each staged element is reused 81 times from SMEM.
#include "common.cuh"

namespace fs {

namespace {

constexpr int TS = 32;            // tile side (output pixels)
constexpr int HS = TS + 8;        // with the 4-px halo of a 9x9 window
constexpr int NT9 = 352;          // 11 warps; 324 = 81 taps x 4 channel blocks are active in the main loop

// dW[kh,kw,ci,co] partial sums of one tile: partial[block][(tap*CI + ci)*CO + co]
template <int CI, int CO>
__global__ void __launch_bounds__(NT9) wgrad9x9_kernel(const float* __restrict__ in, const float* __restrict__ dy,
                                                       float* __restrict__ partial, int H, int W) {
    static_assert(CI * CO == 64 && CI % 4 == 0 && CO % 4 == 0, "channel block must be 4x16 or 16x4");
    constexpr int CIQ = CI / 4, COQ = CO / 4;
    extern __shared__ float4 sm4[];
    float4* in_s = sm4;                         // [HS][HS][CIQ]
    float4* dy_s = sm4 + HS * HS * CIQ;         // [TS][TS][COQ]
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * TS, y0 = blockIdx.y * TS, n = blockIdx.z;
    const float4* in4 = reinterpret_cast<const float4*>(in + (long long)n * H * W * CI);
    const float4* dy4 = reinterpret_cast<const float4*>(dy + (long long)n * H * W * CO);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = t; i < HS * HS * CIQ; i += NT9) {
        int pix = i / CIQ, c4 = i - pix * CIQ;
        int yy = y0 - 4 + pix / HS, xx = x0 - 4 + pix % HS;
        in_s[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(in4 + ((long long)yy * W + xx) * CIQ + c4) : z;
    }
    for (int i = t; i < TS * TS * COQ; i += NT9) {
        int pix = i / COQ, c4 = i - pix * COQ;
        int yy = y0 + pix / TS, xx = x0 + pix % TS;
        dy_s[i] = (yy < H && xx < W) ? __ldg(dy4 + ((long long)yy * W + xx) * COQ + c4) : z;
    }
    __syncthreads();
    if (t >= 324) return;
    const int tap = t >> 2, q = t & 3;
    const int ciq = q / COQ, coq = q - ciq * COQ;
    const int kh = tap / 9, kw = tap - kh * 9;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int py = 0; py < TS; ++py) {
        const float4* ip = in_s + ((py + kh) * HS + kw) * CIQ + ciq;
        const float4* dp = dy_s + (py * TS) * COQ + coq;
#pragma unroll 8
        for (int px = 0; px < TS; ++px) {
            const float4 a = ip[px * CIQ];
            const float4 b = dp[px * COQ];
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
    const long long blk = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    float* out = partial + blk * (81 * 64) + (tap * CI + ciq * 4) * CO + coq * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(out + i * CO) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
}

// out[i] = sum_b partial[b][i]; 32 elements x 8 block-slices per CTA, fixed summation order
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                              int elems, int nblocks) {
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

long long wgrad9x9_partial_floats(int N, int H, int W) {
    return (long long)N * cdiv(H, TS) * cdiv(W, TS) * 81 * 64;
}

// in [N,H,W,CI], dy [N,H,W,CO] (9x9 stride-1 SAME conv: equal spatial dims); out [81,CI,CO]
int launch_wgrad9x9(const float* in, const float* dy, float* out, float* partial, long long partial_cap, int N,
                    int H, int W, int CI, int CO, cudaStream_t st) {
    FS_CHECK((CI == 16 && CO == 4) || (CI == 4 && CO == 16), "wgrad9x9: unsupported channel block %dx%d", CI, CO);
    long long need = wgrad9x9_partial_floats(N, H, W);
    FS_CHECK(need <= partial_cap, "wgrad9x9: partial workspace too small (%lld > %lld floats)", need, partial_cap);
    dim3 grid(cdiv(W, TS), cdiv(H, TS), N);
    const int nblocks = grid.x * grid.y * grid.z;
    if (CI == 16) {
        size_t smem = (size_t)(HS * HS * 4 + TS * TS * 1) * sizeof(float4);
        static bool set16 = false;
        if (!set16) { FS_CUDA(cudaFuncSetAttribute(wgrad9x9_kernel<16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set16 = true; }
        wgrad9x9_kernel<16, 4><<<grid, NT9, smem, st>>>(in, dy, partial, H, W);
    } else {
        size_t smem = (size_t)(HS * HS * 1 + TS * TS * 4) * sizeof(float4);
        static bool set4 = false;
        if (!set4) { FS_CUDA(cudaFuncSetAttribute(wgrad9x9_kernel<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set4 = true; }
        wgrad9x9_kernel<4, 16><<<grid, NT9, smem, st>>>(in, dy, partial, H, W);
    }
    FS_LAUNCH_CHECK();
    reduce_partials_kernel<<<cdiv(81 * 64, 32), 256, 0, st>>>(partial, out, 81 * 64, nblocks);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
