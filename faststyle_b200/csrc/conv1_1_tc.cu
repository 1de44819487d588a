// VGG conv1_1 (3 -> 64 channels, 3x3 SAME, bias + ReLU; reference libs/vgg16.py:45-53) on tcgen05.
//
// A 3-channel input cannot feed the TMA-slab kernel (its K blocks are 64 channels), and as a direct FFMA kernel the
// layer ran at a third of the FMA peak while its real cost is the 64-channel output stream.  Here the im2col tile is
// BUILT in shared memory instead of fetched: four builder warps write, for 128 pixels (8 rows x 16 columns), the
// 9 taps x 3 channels = 27 reduction values per pixel (K padded to 32 = one 64-byte row) straight into the
// 64-byte-swizzled K-major layout the MMA reads; one thread issues the tcgen05.mma of a tile into a double-buffered
// TMEM accumulator; four epilogue warps drain it (thread = pixel), add the bias, apply ReLU and store the split-bf16
// planes conv1_2 reads, the ReLU code bytes of the training pass and / or the fp32 activation.  The weight matrix
// B[64 out][32] is packed from the fp32 HWIO weights by the CTA itself.
//
// Precision: this layer's ReLU mask gates the gradient w.r.t. the image directly, so it keeps fp32-class operands:
// every value is split THREE ways, x = h + m + l with 8 mantissa bits each (exact for fp32), and the six cross
// products of weight >= 2^-16 (hh, hm, mh, hl, lh, mm) are accumulated in fp32 - 12 MMAs (K = 32) per 128 pixels,
// nothing next to the output stream (5 - 9 bytes per output element) that bounds the kernel.
#include <cuda.h>
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace fs {

namespace {

using namespace tcptx;

constexpr int C11_THREADS = 288;          // warps 0-3 epilogue (TMEM lane quadrant = warp), 4-7 builders, 8 MMA issue
constexpr int C11_PLANE = 128 * 64;       // one operand plane of the A tile: 128 rows x 64 B
constexpr int C11_A_STAGE = 3 * C11_PLANE;
constexpr int C11_B_PLANE = 64 * 64;
constexpr int C11_STG_WARP = 4096 + 4096 + 2048;      // per epilogue warp: hi / lo planes (32 px x 128 B), codes (32 px x 64 B)
constexpr int C11_SMEM = 2 * C11_A_STAGE + 3 * C11_B_PLANE + 4 * C11_STG_WARP + 1024 /*align*/ + 256 /*bias*/ + 128 /*barriers*/;

__device__ __forceinline__ uint64_t sdesc64(uint32_t saddr) {      // K-major, SWIZZLE_64B: 64-byte rows, 8-row atoms of 512 B
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}

// byte offset of 16-byte chunk `c` (0..3) of row `r` in a 64B-swizzled tile of 64-byte rows (address bits 4-5 ^= bits 7-8)
__device__ __forceinline__ uint32_t swz64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void split2(float a, float b, uint32_t& h, uint32_t& l) { split_pair(a, b, h, l); }
// three-way split of a pair: a = h + m + l exactly (fp32 has 24 mantissa bits, each part carries 8)
__device__ __forceinline__ void split3(float a, float b, uint32_t& h, uint32_t& m, uint32_t& l) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(hh);
    const float ra = a - hf.x, rb = b - hf.y;
    const __nv_bfloat162 mm = __floats2bfloat162_rn(ra, rb);
    const float2 mf = __bfloat1622float2(mm);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(ra - mf.x, rb - mf.y);
    h = *reinterpret_cast<const uint32_t*>(&hh);
    m = *reinterpret_cast<const uint32_t*>(&mm);
    l = *reinterpret_cast<const uint32_t*>(&ll);
}

// Output path: every epilogue warp owns two image rows x 16 pixels of the tile.  Its lanes write their pixel's 64
// channels into a 128B-swizzled (codes: 64B-swizzled) staging box in shared memory - conflict-free 16-byte stores -
// and one lane hands the box to the TMA engine (cp.async.bulk.tensor store, clipped at the image border).  Scattered
// 32-byte global stores (one 128-byte line per lane and instruction) kept the LSU data pipe at 67 % of its peak.
__global__ void __launch_bounds__(C11_THREADS, 2)
conv1_1_tc_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                  const __grid_constant__ CUtensorMap tm_code, const float* __restrict__ in, const float* __restrict__ w,
                  const float* __restrict__ bias, float* __restrict__ out, int has_split, int has_code, int N, int H, int W,
                  int tilesX, int tilesY) {
    FS_PDL_TRIGGER();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smA = smem;                                   // [2 stages][h | m | l][128 rows x 64 B]
    uint8_t* smB = smem + 2 * C11_A_STAGE;                 // [h | m | l][64 rows x 64 B]
    uint8_t* smS = smB + 3 * C11_B_PLANE;                  // [4 epilogue warps][hi 4 KB | lo 4 KB | codes 2 KB]
    float* s_bias = reinterpret_cast<float*>(smS + 4 * C11_STG_WARP);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 64);
    uint64_t* a_full = bars;            // [2] 128 builder arrivals
    uint64_t* a_empty = bars + 2;       // [2] tcgen05.commit
    uint64_t* t_full = bars + 4;        // [2] tcgen05.commit
    uint64_t* t_empty = bars + 6;       // [2] 4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = threadIdx.x;
    if (t == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_slot, 128);             // 2 accumulators x 64 fp32 columns
    if (warp == 0 && lane == 0) { prefetch_tmap(&tm_hi); prefetch_tmap(&tm_lo); prefetch_tmap(&tm_code); }
    FS_PDL_WAIT();
    // weights: B[n][k] = w[tap][c][n] for k = tap * 3 + c < 27 (fp32 [9][4][64]), zero up to k = 31; bias
    for (int i = t; i < 64 * 4; i += C11_THREADS) {        // one 16-byte chunk (8 k values) of one output-channel row
        const int n = i >> 2, c = i & 3;
        uint32_t h[4], m[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k0 = c * 8 + 2 * j, k1 = k0 + 1;
            const float a = k0 < 27 ? __ldg(w + ((k0 / 3) * 4 + k0 % 3) * 64 + n) : 0.f;
            const float b = k1 < 27 ? __ldg(w + ((k1 / 3) * 4 + k1 % 3) * 64 + n) : 0.f;
            split3(a, b, h[j], m[j], l[j]);
        }
        *reinterpret_cast<uint4*>(smB + swz64(n, c)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(smB + C11_B_PLANE + swz64(n, c)) = make_uint4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<uint4*>(smB + 2 * C11_B_PLANE + swz64(n, c)) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    if (t < 64) s_bias[t] = bias ? __ldg(bias + t) : 0.f;
    fence_proxy_async();                                   // generic-proxy writes of B -> visible to the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long total = (long long)N * tilesX * tilesY;

    if (warp >= 4 && warp < 8) {
        // ===================== builders: im2col rows of one tile =====================
        const int r = t - 128;                             // row of the A tile = pixel of the 8 x 16 tile
        const int prow = r >> 4, pcol = r & 15;
        int s = 0; uint32_t ph = 0;
        for (long long u = blockIdx.x; u < total; u += gridDim.x) {
            const int tx = (int)(u % tilesX); const long long q = u / tilesX;
            const int ty = (int)(q % tilesY), n = (int)(q / tilesY);
            const int y = ty * 8 + prow, x = tx * 16 + pcol;
            float v[28];                                   // k = tap * 3 + c
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
                float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (yy >= 0 && yy < H && xx >= 0 && xx < W)
                    q4 = __ldg(reinterpret_cast<const float4*>(in) + ((long long)n * H + yy) * W + xx);
                v[3 * tap] = q4.x; v[3 * tap + 1] = q4.y; v[3 * tap + 2] = q4.z;
            }
            v[27] = 0.f;
            mbar_wait(&a_empty[s], ph ^ 1);
            uint8_t* a0 = smA + s * C11_A_STAGE;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t h[4], m[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k0 = c * 8 + 2 * j;
                    if (k0 < 28) split3(v[k0], v[k0 + 1], h[j], m[j], l[j]);
                    else { h[j] = 0u; m[j] = 0u; l[j] = 0u; }
                }
                const uint32_t o = swz64(r, c);
                *reinterpret_cast<uint4*>(a0 + o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(a0 + C11_PLANE + o) = make_uint4(m[0], m[1], m[2], m[3]);
                *reinterpret_cast<uint4*>(a0 + 2 * C11_PLANE + o) = make_uint4(l[0], l[1], l[2], l[3]);
            }
            fence_proxy_async();
            mbar_arrive(&a_full[s]);
            if (++s == 2) { s = 0; ph ^= 1; }
        }
    } else if (warp == 8) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint64_t bd[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) bd[i] = sdesc64(smem_u32(smB + i * C11_B_PLANE));
        int s = 0; uint32_t ph = 0;
        for (long long u = blockIdx.x; u < total; u += gridDim.x) {
            mbar_wait(&t_empty[s], ph ^ 1);
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            uint64_t ad[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) ad[i] = sdesc64(smem_u32(smA + s * C11_A_STAGE + i * C11_PLANE));
            if (elect_one()) {
                const uint32_t d = tmem_base + (uint32_t)(s * 64);
                // (A part, B part) pairs with combined weight >= 2^-16: h*h, h*m, m*h, h*l, l*h, m*m
                constexpr int PA[6] = {0, 0, 1, 0, 2, 1}, PB[6] = {0, 1, 0, 2, 0, 1};
#pragma unroll
                for (int prod = 0; prod < 6; ++prod) {
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_bf16(d, ad[PA[prod]] + (uint64_t)(k * 2), bd[PB[prod]] + (uint64_t)(k * 2), idesc,
                                    (prod == 0 && k == 0) ? 0u : 1u);
                }
                tc_commit(&a_empty[s]);
                tc_commit(&t_full[s]);
            }
            __syncwarp();
            if (++s == 2) { s = 0; ph ^= 1; }
        }
    } else if (warp < 4) {
        // ===================== epilogue: thread = pixel, 2 x 32 channels =====================
        const int r = warp * 32 + lane, prow = r >> 4, pcol = r & 15;
        uint8_t* stg_hi = smS + warp * C11_STG_WARP;       // row = lane (y-major over the warp's 2 x 16 pixels)
        uint8_t* stg_lo = stg_hi + 4096;
        uint8_t* stg_cd = stg_hi + 8192;
        int s = 0; uint32_t ph = 0;
        for (long long u = blockIdx.x; u < total; u += gridDim.x) {
            const int tx = (int)(u % tilesX); const long long q = u / tilesX;
            const int ty = (int)(q % tilesY), n = (int)(q / tilesY);
            const int y = ty * 8 + prow, x = tx * 16 + pcol;
            const bool ok = y < H && x < W;
            const long long o = (((long long)n * H + y) * W + x) * 64;
            mbar_wait(&t_full[s], ph);
            tc_fence_after();
            if (lane == 0) bulk_wait_read0();              // the previous tile's boxes have left the staging buffers
            __syncwarp();
#pragma unroll 1
            for (int ch = 0; ch < 2; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 64 + ch * 32), v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + s_bias[ch * 32 + i], 0.f);
                if (out && ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) stg256(out + o + ch * 32 + i, v + i);
                }
                if (has_split) {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {       // 16-byte chunk ch * 4 + jj of this pixel's 128-byte row
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) split2(v[jj * 8 + 2 * j], v[jj * 8 + 2 * j + 1], hw[j], lw[j]);
                        const uint32_t off = (uint32_t)(lane * 128 + (((ch * 4 + jj) ^ (lane & 7)) << 4));
                        *reinterpret_cast<uint4*>(stg_hi + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        *reinterpret_cast<uint4*>(stg_lo + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                }
                if (has_code) {                            // ReLU codes (Conv3x3TcArgs::ref_code): bit 0 = value > 0
#pragma unroll
                    for (int jj = 0; jj < 2; ++jj) {       // 16-byte chunk ch * 2 + jj of this pixel's 64-byte row
                        uint32_t cw[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int b = jj * 16 + 4 * j;
                            cw[j] = (v[b] > 0.f ? 1u : 0u) | (v[b + 1] > 0.f ? 0x100u : 0u) |
                                    (v[b + 2] > 0.f ? 0x10000u : 0u) | (v[b + 3] > 0.f ? 0x1000000u : 0u);
                        }
                        const uint32_t off = (uint32_t)(lane * 64 + (((ch * 2 + jj) ^ ((lane >> 1) & 3)) << 4));
                        *reinterpret_cast<uint4*>(stg_cd + off) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();                           // staging writes -> visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&t_empty[s]);
                const int y0 = ty * 8 + warp * 2, x0 = tx * 16;
                if (y0 < H) {
                    if (has_split) { tma_store_4d(&tm_hi, stg_hi, 0, x0, y0, n); tma_store_4d(&tm_lo, stg_lo, 0, x0, y0, n); }
                    if (has_code) tma_store_4d(&tm_code, stg_cd, 0, x0, y0, n);
                }
                bulk_commit();
            }
            if (++s == 2) { s = 0; ph ^= 1; }
        }
        if (lane == 0) bulk_wait0();                       // all boxes written before the CTA retires
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 128);
}

}  // namespace

// in [N,H,W,4] fp32 (4th channel zero), w [9,4,64] fp32, bias [64] -> any of: out fp32 [N,H,W,64], split planes, codes
int launch_conv1_1_tc(const float* in, const float* w, const float* bias, float* out, void* split_hi, void* split_lo,
                      unsigned char* code, int N, int H, int W, cudaStream_t st) {
    FS_CHECK(in && w, "conv1_1_tc: NULL argument");
    FS_CHECK(out || split_hi || code, "conv1_1_tc: no output requested");
    FS_CHECK((split_hi == nullptr) == (split_lo == nullptr), "conv1_1_tc: split output needs both planes");
    const int tilesX = cdiv(W, 16), tilesY = cdiv(H, 8);
    const long long total = (long long)N * tilesX * tilesY;
    // store maps: box = 2 rows x 16 pixels x 64 channels (an absent output gets a map over a present one; never used)
    CUtensorMap tm_hi, tm_lo, tm_code;
    const void* any2 = split_hi ? split_hi : (const void*)in;
    FS_TRY(tc_make_map_nhwc(&tm_hi, any2, 2, N, H, W, 64, 64, 16, 2, 128));
    FS_TRY(tc_make_map_nhwc(&tm_lo, split_lo ? split_lo : any2, 2, N, H, W, 64, 64, 16, 2, 128));
    FS_TRY(tc_make_map_nhwc(&tm_code, code ? (const void*)code : any2, 1, N, H, W, 64, 64, 16, 2, 64));
    FS_DYN_SMEM(conv1_1_tc_kernel, C11_SMEM);
    const long long cap = 2LL * num_sms();
    const int grid = (int)(total < cap ? total : cap);
    launch_k(conv1_1_tc_kernel, dim3(grid), dim3(C11_THREADS), C11_SMEM, st, tm_hi, tm_lo, tm_code, in, w, bias, out,
             split_hi ? 1 : 0, code ? 1 : 0, N, H, W, tilesX, tilesY);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
