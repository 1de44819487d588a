// tcgen05 implicit-GEMM 3x3 stride-1 convolution (forward and data gradient) for NHWC
// activations whose channel counts are multiples of 64.
//
//   GEMM view per output tile: D[128*NACC pixels, BN out-channels] += A[pixels, 64 ch] * B[64 ch, BN]
//   summed over 9 taps x (C/64) channel blocks.
//
// Precision: operands are "split bf16" (x ~= hi + lo, 16 mantissa bits); each logical product is
// three kind::f16 MMAs (hi*hi + hi*lo + lo*hi) accumulated in fp32 in TMEM.  That restores
// fp32-class accuracy (~1e-5 on the unit pixel scale) at 3 MMA issues.
//
// Data movement: a persistent CTA per SM; a TMA producer thread streams
//   * A "slabs": for every (channel block, kw) one box {64 ch, 16 px, TH+2 rows} of the hi and lo
//     planes (128B-swizzled, zero-filled out of bounds = conv zero padding for free).  The three kh
//     taps of that kw are the SAME slab at +kh*2048 B (16 px * 128 B), so the slab is read once from
//     L2 and used by 3 taps: 3.4-3.75 slab loads per tile instead of 9.
//   * B tiles: per tap one [BN x 64] box of the packed hi/lo weights.
// One elected thread issues tcgen05.mma into double-buffered TMEM accumulators; four epilogue warps
// drain TMEM (tcgen05.ld), apply bias/ReLU/addend/mask and store fp32 and/or split-bf16 NHWC while the
// next tile's MMAs run.
#include <cuda.h>
#include <stdlib.h>
#include "tc.cuh"
#include "tc_ptx.cuh"

namespace fs {

namespace {

using namespace tcptx;

constexpr int TW = 16;            // tile width in pixels (16 px * 128 B = 2048 B per slab row)
constexpr int KB = 64;            // channels per K block (= one 128-byte swizzle row of bf16)

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row (1024 B) swizzle atoms.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);     // start address, 16-byte units
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                       // layout: SWIZZLE_128B
    return d;
}

template <int BN>
__device__ __forceinline__ uint32_t make_idesc() {
    return (1u << 4)                // D format: fp32
         | (1u << 7)                // A format: bf16
         | (1u << 10)               // B format: bf16
         | ((uint32_t)(BN >> 3) << 17)      // N
         | ((uint32_t)(128 >> 4) << 24);    // M = 128
}

struct TcParams {
    int N, H, W, C, OH, OW, OC, pad;
    int taps;                  // vertical taps: 3: 3x3 convolution, 2: 2x2 (collapsed stride-2 forms), 1: 1x1 (per-sample GEMM),
                               // 9: the 9-row forms of the 9x9 layers
    int taps_w, pad_x;         // horizontal taps / left padding (== taps / pad for the square forms)
    int in_s2d, out_d2s;       // space-to-depth input view (5-D tensor map) / depth-to-space fp32 store
    int w_sample_rows;         // rows of the packed weight matrix per sample (0 = shared weights)
    int tilesX, tilesY, tilesN;
    long long total_tiles;     // work units: spatial tiles (per-sample pairs of them for the CTA-pair kernel) x tilesN
    const float* bias; const float* addend; const float* ref;
    const float* pool_grad; const float* ctarget; float cw2;
    const unsigned char* ref8; unsigned char* out_code;
    __nv_bfloat16* pool_hi; __nv_bfloat16* pool_lo;
    int relu, add_crop, addH, addW;
    float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
    double* stats; int stats_c;       // [N][stats_c][2] running (sum, sum of squares) of the raw output per real channel
};

// Sum of s[0..31] over the 32 lanes of a warp for 32 quantities at once: lane L returns the total of quantity L.
// Butterfly transposition: 16+8+4+2+1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum(float* s, int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const float send = up ? s[i] : s[i + step];
            const float keep = up ? s[i + step] : s[i];
            s[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return s[0];
}

// HALO = slab rows beyond the tile height = (vertical taps - 1): 2 for the 3x3 / 2x2 / 1x1 forms, 8 for the 9-row
// forms of the 9x9 layers
template <int TH, int BN, bool PAIR = false, int HALO = 2>
struct Cfg {
    static constexpr int NACC = TH / 8;
    static constexpr int A_STAGES = (TH == 8 && HALO == 2) ? 3 : 2;   // small tiles: deeper slab ring (latency-bound)
    static constexpr int B_ROWS = PAIR ? BN / 2 : BN;              // weight rows held by one CTA
    static constexpr int B_STAGES = (B_ROWS <= 64) ? 4 : 2;
    static constexpr int SLAB_BYTES = (TH + HALO) * TW * 128;     // one plane
    static constexpr int A_STAGE_BYTES = 2 * SLAB_BYTES;          // hi + lo
    static constexpr int BTILE_BYTES = B_ROWS * 128;              // one plane
    static constexpr int B_STAGE_BYTES = 2 * BTILE_BYTES;
    static constexpr int TMEM_COLS = 2 * NACC * BN;               // double-buffered accumulators
    static constexpr int SMEM_BYTES = A_STAGES * A_STAGE_BYTES + B_STAGES * B_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4.. epilogue (EW / 4 warps per TMEM lane quadrant).
// The epilogues of the memory-heavy launches (64-channel VGG layers, Gram backward) are chains of dependent TMEM loads,
// shuffles, conversions and scattered stores: with 8 warps they are instruction-latency-bound (ncu: tensor pipe 45 % and
// no memory pipe above 55 % on conv1_2).  Launches without fused statistics therefore run SIXTEEN epilogue warps; the
// CTA's registers (640 threads x 96 at launch: setmaxnreg moves registers only WITHIN the CTA's own allocation) are
// re-divided: producer / MMA warpgroup down to 64, the four epilogue warpgroups up to 104 (128 x 64 + 512 x 104 =
// 640 x 96).  The statistics variant keeps 64 running sums per thread and stays at 8 warps x 168 registers.
// EPI: 0 = lean epilogue (16 warps), 1 = fused InstanceNorm statistics (8 warps), 2 = fp32 reference tensor (ReLU mask /
// pooling / content term from the fp32 activation: 32 more live registers; 8 warps)
__host__ __device__ constexpr int tc_epi_warps(int epi) { return epi == 0 ? 16 : 8; }
__host__ __device__ constexpr int tc_threads(int epi) { return 128 + 32 * tc_epi_warps(epi); }

template <int BN, bool PAIR>
__device__ __forceinline__ uint32_t make_idesc_p() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
}

// PAIR = true: two CTAs of a cluster (one TPC) work as a CTA pair on TWO spatial tiles at once.  The leader
// (cluster rank 0) issues `tcgen05.mma.cta_group::2` with M = 256: rows 0-127 are the leader's pixels, rows
// 128-255 the peer's, each read from the owning CTA's shared memory; the B (weight) tile of BN rows is SPLIT -
// each CTA loads and holds BN/2 rows - so per MMA a CTA's shared memory feeds 4 KB of A + half of B instead of
// all of it, and the weights cross L2->SMEM once per pair instead of once per CTA.  (At BN = 128 the single-CTA
// form reads 128 B/clk of operands, the whole shared-memory bandwidth, before the TMA writes; the pair reads 96.)
// Each CTA keeps its own 128 accumulator rows in its own TMEM and runs its own epilogue.
// Barrier protocol: `*_full` barriers live in the leader only (both CTAs' TMA loads count their bytes there);
// `*_empty` and `t_full` exist in both CTAs and are signalled by the leader's multicast tcgen05.commit; the
// leader's `t_empty` collects the arrivals of both CTAs' epilogue warps (the peer's arrive remotely).
template <int TH, int BN, int EPI, bool PAIR, int HALO>
__global__ void __launch_bounds__(tc_threads(EPI), 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const TcParams p) {
    FS_PDL_TRIGGER();
    using K = Cfg<TH, BN, PAIR, HALO>;
    constexpr int A_STAGES = K::A_STAGES;
    constexpr int NCTA = PAIR ? 2 : 1;
    constexpr bool STATS = EPI == 1, F32REF = EPI == 2;
    constexpr int EW = tc_epi_warps(EPI);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + A_STAGES * K::A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smemB + K::B_STAGES * K::B_STAGE_BYTES);
    uint64_t* a_full = bars;                       // [A_STAGES]
    uint64_t* a_empty = a_full + A_STAGES;         // [A_STAGES]
    uint64_t* b_full = a_empty + A_STAGES;         // [B_STAGES]
    uint64_t* b_empty = b_full + K::B_STAGES;      // [B_STAGES]
    uint64_t* t_full = b_empty + K::B_STAGES;      // [2]
    uint64_t* t_empty = t_full + 2;                // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int CB = p.C / KB;
    const uint32_t cta = PAIR ? cluster_ctarank() : 0u;           // 0 = leader
    // work distribution over the CTAs (CTA pairs): strided - at any time the chip works on a band of consecutive tiles,
    // so neighbouring tiles' halos meet in L2 (the large VGG layers) - or, for launches with fused statistics,
    // contiguous blocks of tiles per CTA so that a CTA stays within one or two samples (one statistics flush each).
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    long long u_begin = worker, u_end = p.total_tiles, u_stride = nworkers;
    if (STATS) {
        const long long per = (p.total_tiles + nworkers - 1) / nworkers;
        u_begin = worker * per;
        u_end = u_begin + per < p.total_tiles ? u_begin + per : p.total_tiles;
        u_stride = 1;
    }

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < K::B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], EW * NCTA); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc2(tmem_slot, K::TMEM_COLS);
        else tmem_alloc(tmem_slot, K::TMEM_COLS);
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all();         // the peer's barriers are initialised before anything remote touches them
    else __syncthreads();
    tc_fence_after();
    FS_PDL_WAIT();                        // everything above is CTA-local setup
    const uint32_t tmem_base = *tmem_slot;

    // work unit u -> (nt, spatial tile of THIS CTA).  PAIR: unit = (nt, sample, pair of consecutive tiles of that sample):
    // both CTAs always work on the SAME sample (per-sample weight matrices stay consistent across the split B tile);
    // the peer of an odd tile count gets a dummy tile below the image (its TMA boxes are out of bounds = zeros, it
    // still loads its half of the weights, nothing is stored).  Returns whether the tile is real.
    auto decode = [&](long long u, int& nt, int& tx, int& ty, int& n) -> bool {
        nt = (int)(u % p.tilesN);
        long long r = u / p.tilesN;
        if (PAIR) {
            const int tps = p.tilesX * p.tilesY, pps = (tps + 1) >> 1;
            n = (int)(r / pps);
            const int s = 2 * (int)(r - (long long)n * pps) + (int)cta;
            if (s >= tps) { tx = 0; ty = p.tilesY; return false; }
            ty = s / p.tilesX; tx = s - ty * p.tilesX;
            return true;
        }
        tx = (int)(r % p.tilesX); r /= p.tilesX;
        ty = (int)(r % p.tilesY);
        n = (int)(r / p.tilesY);
        return true;
    };

    if (warp < 4) {
    if (EW == 16) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");     // (warpgroup-uniform) registers for the epilogue
    if (warp == 0 && lane == 0) {
        // ===================== TMA producer =====================
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        for (long long u = u_begin; u < u_end; u += u_stride) {
            int nt, tx, ty, n;
            decode(u, nt, tx, ty, n);
            const int y0 = ty * TH - p.pad, x0 = tx * TW - p.pad_x;
            const int n0 = nt * BN + (int)cta * K::B_ROWS;
            for (int cb = 0; cb < CB; ++cb) {
                for (int kw = 0; kw < p.taps_w; ++kw) {
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    uint8_t* dst = smemA + sa * K::A_STAGE_BYTES;
                    if (!PAIR || cta == 0) mbar_expect_tx(&a_full[sa], NCTA * K::A_STAGE_BYTES);
                    if (PAIR) {
                        const uint32_t fb = mapa_u32(&a_full[sa], 0);        // the leader's barrier counts both CTAs' bytes
                        if (p.in_s2d) {
                            tma2_load_5d(dst, &tmA_hi, fb, 0, x0 + kw, cb, y0, n);
                            tma2_load_5d(dst + K::SLAB_BYTES, &tmA_lo, fb, 0, x0 + kw, cb, y0, n);
                        } else {
                            tma2_load_4d(dst, &tmA_hi, fb, cb * KB, x0 + kw, y0, n);
                            tma2_load_4d(dst + K::SLAB_BYTES, &tmA_lo, fb, cb * KB, x0 + kw, y0, n);
                        }
                    } else if (p.in_s2d) {       // channel block cb = row parity p of the [2H,2W,32] source (see make_act_map_s2d)
                        tma_load_5d(dst, &tmA_hi, &a_full[sa], 0, x0 + kw, cb, y0, n);
                        tma_load_5d(dst + K::SLAB_BYTES, &tmA_lo, &a_full[sa], 0, x0 + kw, cb, y0, n);
                    } else {
                        tma_load_4d(dst, &tmA_hi, &a_full[sa], cb * KB, x0 + kw, y0, n);
                        tma_load_4d(dst + K::SLAB_BYTES, &tmA_lo, &a_full[sa], cb * KB, x0 + kw, y0, n);
                    }
                    if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
                    for (int kh = 0; kh < p.taps; ++kh) {
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        uint8_t* bd = smemB + sb * K::B_STAGE_BYTES;
                        if (!PAIR || cta == 0) mbar_expect_tx(&b_full[sb], NCTA * K::B_STAGE_BYTES);
                        const int row = ((kh * p.taps_w + kw) * CB + cb) * p.OC + n0 + n * p.w_sample_rows;
                        if (PAIR) {
                            const uint32_t fb = mapa_u32(&b_full[sb], 0);
                            tma2_load_2d(bd, &tmB_hi, fb, 0, row);
                            tma2_load_2d(bd + K::BTILE_BYTES, &tmB_lo, fb, 0, row);
                        } else {
                            tma_load_2d(bd, &tmB_hi, &b_full[sb], 0, row);
                            tma_load_2d(bd + K::BTILE_BYTES, &tmB_lo, &b_full[sb], 0, row);
                        }
                        if (++sb == K::B_STAGES) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 && (!PAIR || cta == 0)) {
        // ===================== MMA issuer (the leader CTA of a pair) =====================
        // The whole warp runs the (warp-uniform) control flow and barrier waits; one elected lane issues.
        // Shared-memory descriptors are built once per stage; each tcgen05.mma then costs one 64-bit add per
        // operand, so the issue rate stays above the 32-64 tensor cycles an MMA takes.
        const uint32_t idesc = make_idesc_p<BN, PAIR>();
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        int as = 0; uint32_t pt = 0;
        for (long long u = u_begin; u < u_end; u += u_stride) {
            mbar_wait(&t_empty[as], pt ^ 1);
            tc_fence_after();
            const uint32_t acc_base = tb + (uint32_t)(as * K::NACC * BN);
            bool first = true;
            for (int cb = 0; cb < CB; ++cb) {
                for (int kw = 0; kw < p.taps_w; ++kw) {
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    const uint64_t ad_hi = make_sdesc(smem_u32(smemA + sa * K::A_STAGE_BYTES));
                    const uint64_t ad_lo = make_sdesc(smem_u32(smemA + sa * K::A_STAGE_BYTES + K::SLAB_BYTES));
                    for (int kh = 0; kh < p.taps; ++kh) {
                        mbar_wait(&b_full[sb], pb);
                        tc_fence_after();
                        const uint64_t bd_hi = make_sdesc(smem_u32(smemB + sb * K::B_STAGE_BYTES));
                        const uint64_t bd_lo = make_sdesc(smem_u32(smemB + sb * K::B_STAGE_BYTES + K::BTILE_BYTES));
                        if (elect_one()) {
#pragma unroll
                            for (int acc = 0; acc < K::NACC; ++acc) {
                                const uint64_t aoff = (uint64_t)((acc * 8 + kh) * (TW * 128 / 16));   // 16-byte units
                                const uint32_t d = acc_base + (uint32_t)(acc * BN);
#pragma unroll
                                for (int prod = 0; prod < 3; ++prod) {
                                    const uint64_t ad = (prod == 2 ? ad_lo : ad_hi) + aoff;
                                    const uint64_t bd = (prod == 1 ? bd_lo : bd_hi);
#pragma unroll
                                    for (int k4 = 0; k4 < 4; ++k4) {
                                        const uint32_t accum = (first && prod == 0 && k4 == 0) ? 0u : 1u;
                                        if (PAIR) tc_mma2_bf16(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, accum);
                                        else tc_mma_bf16(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, accum);
                                    }
                                }
                            }
                            if (PAIR) tc_commit2(&b_empty[sb]); else tc_commit(&b_empty[sb]);
                        }
                        __syncwarp();
                        first = false;
                        if (++sb == K::B_STAGES) { sb = 0; pb ^= 1; }
                    }
                    if (elect_one()) { if (PAIR) tc_commit2(&a_empty[sa]); else tc_commit(&a_empty[sa]); }
                    __syncwarp();
                    if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
                }
            }
            if (elect_one()) { if (PAIR) tc_commit2(&t_full[as]); else tc_commit(&t_full[as]); }
            __syncwarp();
            if (++as == 2) { as = 0; pt ^= 1; }
        }
    }
    } else {
        if (EW == 16) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================== epilogue (EW warps: EW / 4 per 32-lane TMEM quadrant) =====================
        // The two warps of a quadrant take alternate (accumulator, 32-channel chunk) items: the epilogue is a
        // chain of dependent global loads (addend / ReLU-mask reference) and stores per item, so a second warp
        // per quadrant doubles the memory-level parallelism of the memory-heavy launches (1x1 Gram backward,
        // 64-channel layers at 256^2).
        const int ew = (warp - 4) & 3;                 // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
        const int eg = (warp - 4) >> 2;                // 0 .. EW/4-1: which share of the items
        const int row = ew * 32 + lane;                // MMA row = pixel within the 8x16 sub-tile
        const int prow = row >> 4, pcol = row & 15;
        // fused InstanceNorm statistics (STATS): per-thread running (sum, sum of squares) over the pixels this thread
        // has seen of the current sample; a butterfly transposition leaves lane L with the totals of channel L of
        // the chunk, which go to the per-(sample, real channel) fp64 accumulators with one atomic each.  A CTA's
        // tiles are sample-major, so this happens about once per sample and warp.
        float st1[32], st2[32];
        int stat_n = -1;
        if (STATS) {
#pragma unroll
            for (int i = 0; i < 32; ++i) { st1[i] = 0.f; st2[i] = 0.f; }
        }
        auto stats_flush = [&]() {
            if (stat_n < 0) return;
            const float s1 = warp_transpose_sum(st1, lane);
            const float s2 = warp_transpose_sum(st2, lane);
            // STATS_REPLICAS copies of the accumulators (by CTA): same-address fp64 atomics serialise in L2 at ~50 ns
            // each, and with every CTA adding to the same N x C x 2 addresses they cost more than the pass they save
            const int cch = ((BN == 64 ? eg * 32 : 0) + lane) % p.stats_c;
            double* sp = p.stats + (((long long)(worker % STATS_REPLICAS) * p.N + stat_n) * p.stats_c + cch) * 2;
            atomicAdd(sp, (double)s1);
            atomicAdd(sp + 1, (double)s2);
#pragma unroll
            for (int i = 0; i < 32; ++i) { st1[i] = 0.f; st2[i] = 0.f; }
        };
        int as = 0; uint32_t pt = 0;
        for (long long u = u_begin; u < u_end; u += u_stride) {
            int nt, tx, ty, n;
            const bool tile_ok = decode(u, nt, tx, ty, n);
            if (STATS && tile_ok && n != stat_n) { stats_flush(); stat_n = n; }
            mbar_wait(&t_full[as], pt);
            tc_fence_after();
#pragma unroll 1
            for (int item = eg; item < K::NACC * (BN / 32); item += EW / 4) {
                const int acc = item / (BN / 32), ch = item - acc * (BN / 32);
                const int oy = ty * TH + acc * 8 + prow, ox = tx * TW + pcol;
                const bool ok = tile_ok && oy < p.OH && ox < p.OW;
                const long long pix = ((long long)n * p.OH + oy) * p.OW + ox;
                const float* addp = nullptr;
                if (p.addend && ok) {
                    int ay = oy - p.add_crop, ax = ox - p.add_crop;
                    if (ay >= 0 && ay < p.addH && ax >= 0 && ax < p.addW)
                        addp = p.addend + (((long long)n * p.addH + ay) * p.addW + ax) * p.OC;
                }
                {
                    // the ReLU reference (and the first chunk of the pooled gradient) are requested BEFORE the
                    // accumulator is pulled out of TMEM: four independent 256-bit loads in flight per thread instead of
                    // one load latency per 8-channel chunk (the epilogue of the memory-heavy launches was a chain of
                    // serialised HBM round trips: Gram backward at conv1_2 ran at 2.7 TB/s)
                    float rb[32];
                    float gq[8];
                    uint32_t cw8[8];
                    const int rc0 = nt * BN + ch * 32;
                    const float* gp = nullptr;
                    if (!STATS && p.ref8) {                    // one code byte per element instead of the fp32 reference
                        if (ok) {
                            ldg256_b32(p.ref8 + pix * p.OC + rc0, cw8);
                            if (p.pool_grad) {
                                gp = p.pool_grad + (((long long)n * ((p.OH + 1) >> 1) + (oy >> 1)) * ((p.OW + 1) >> 1) + (ox >> 1)) * p.OC + rc0;
                                ldg256(gp, gq);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) cw8[i] = 0u;
                        }
                    } else
                    if (F32REF && p.ref && !p.out_d2s) {
                        if (ok) {
                            const float* rp = p.ref + pix * p.OC + rc0;
#pragma unroll
                            for (int i = 0; i < 32; i += 8) ldg256(rp + i, rb + i);
                            if (p.pool_grad) {
                                gp = p.pool_grad + (((long long)n * ((p.OH + 1) >> 1) + (oy >> 1)) * ((p.OW + 1) >> 1) + (ox >> 1)) * p.OC + rc0;
                                ldg256(gp, gq);
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i) rb[i] = -INFINITY;
                        }
                    }
                    float v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) +
                                           (uint32_t)(as * K::NACC * BN + acc * BN + ch * 32);
                    tmem_ld32(taddr, v);                // warp-collective: executed by all lanes
                    if (STATS && ok) {
                        // InstanceNorm statistics of the raw conv output, fused here so that the activation is not
                        // read again: every thread keeps running sums of ITS pixels for the 32 channels of its
                        // chunk (the chunk's real channels are the same for every item this warp handles, see
                        // launch_conv3x3_tc); lanes are combined once per sample (stats_flush below).
#pragma unroll
                        for (int i = 0; i < 32; ++i) { st1[i] += v[i]; st2[i] = fmaf(v[i], v[i], st2[i]); }
                    }
                    if (ok && p.out_d2s) {
                        const int cq = nt * (BN / 32) + ch;           // which 32-channel quarter of OC = 128
                        float* op;
                        if (p.out_d2s == 1) {
                            // 32 output channels = one sub-pixel phase pq of the 2x upsampled result (OC = 4 * 32)
                            const long long pix2 = ((long long)n * (2 * p.OH) + 2 * oy + (cq >> 1)) * (2 * p.OW) + 2 * ox + (cq & 1);
                            op = p.out_f32 + pix2 * 32;
                        } else {
                            // paired pixels, OC = (e, p, q, 16): quarter (e, p) = both column phases q of row phase p of
                            // source pixel 2*ox+e -> 2 adjacent pixels x 16 channels of the [2OH, 4OW, 16] result
                            const long long pix2 = ((long long)n * (2 * p.OH) + 2 * oy + (cq & 1)) * (4 * p.OW) + 4 * ox + 2 * (cq >> 1);
                            op = p.out_f32 + pix2 * 16;
                        }
#pragma unroll
                        for (int i = 0; i < 32; i += 8) stg256(op + i, v + i);
                    } else if (ok || p.pool_grad || p.pool_hi) {   // (the pooling shuffles need every lane of the warp)
                        const int c0 = nt * BN + ch * 32;
                        if (p.bias) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + i));
                                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                            }
                        }
                        if (addp) {
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                float b[8];
                                ldg256(addp + c0 + i, b);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[i + j] += b[j];
                            }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (!STATS && p.ref8) {
                            // v = bit0 * (v + bit1 * pool_grad): ReLU mask and max-pool routing from the code bytes
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                float g[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) g[j] = p.pool_grad ? gq[j] : 0.f;
                                if (p.pool_grad && ok && i + 8 < 32) ldg256(gp + i + 8, gq);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint32_t code = cw8[(i + j) >> 2] >> (8 * ((i + j) & 3));
                                    const float a = v[i + j] + ((code & 2u) ? g[j] : 0.f);
                                    v[i + j] = (code & 1u) ? a : 0.f;
                                }
                            }
                        } else
                        if (F32REF && p.ref) {
                            // v = mask(ref > 0) * (v + route(pool_grad) + cw2 * (ref - ctarget)): the backward of
                            // ReLU, of a following 2x2 max-pool (gradient to the FIRST maximum of the window in scan
                            // order, as pool_bwd_combine_kernel) and of the content loss, fused.  The four pixels of a
                            // pooling window are lanes L, L^1, L^16, L^17 of this warp (a warp covers 2 rows x 16
                            // columns, both even-aligned); pixels outside the image count as -inf.
                            const float* tp = p.ctarget ? p.ctarget + pix * p.OC + c0 : nullptr;
                            const int kme = ((lane >> 4) & 1) * 2 + (lane & 1);
#pragma unroll
                            for (int i = 0; i < 32; i += 8) {
                                const float* b = rb + i;
                                if (p.pool_grad) {
                                    float g[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) g[j] = gq[j];
                                    if (ok && i + 8 < 32) ldg256(gp + i + 8, gq);     // next chunk's gradient while this one is routed
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float o1 = __shfl_xor_sync(0xffffffffu, b[j], 1);
                                        const float o2 = __shfl_xor_sync(0xffffffffu, b[j], 16);
                                        const float o3 = __shfl_xor_sync(0xffffffffu, b[j], 17);
                                        // position q = kme ^ m: earlier positions must be strictly smaller, later ones not larger
                                        const bool w1 = (kme ^ 1) < kme ? o1 < b[j] : o1 <= b[j];
                                        const bool w2 = (kme ^ 2) < kme ? o2 < b[j] : o2 <= b[j];
                                        const bool w3 = (kme ^ 3) < kme ? o3 < b[j] : o3 <= b[j];
                                        if (ok && w1 && w2 && w3) v[i + j] += g[j];
                                    }
                                }
                                if (tp && ok) {
                                    float t[8];
                                    ldg256(tp + i, t);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) v[i + j] = fmaf(p.cw2, b[j] - t[j], v[i + j]);
                                }
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[i + j] = b[j] > 0.f ? v[i + j] : 0.f;
                            }
                        }
                        uint32_t ismax = 0u;          // bit c: channel c of this pixel equals its window's maximum
                        if (p.pool_hi) {
                            // fused 2x2 stride-2 SAME max-pool of the result (libs/vgg16.py:67-71): the window's four
                            // pixels are lanes L, L^1, L^16, L^17; the top-left lane stores the pooled split planes
                            const bool writer = ok && (lane & 17) == 0;
                            const long long ppix = ((long long)n * ((p.OH + 1) >> 1) + (oy >> 1)) * ((p.OW + 1) >> 1) + (ox >> 1);
#pragma unroll
                            for (int i = 0; i < 32; i += 16) {
                                float m[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j) {
                                    float x = ok ? v[i + j] : -INFINITY;
                                    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
                                    x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 16));
                                    m[j] = x;
                                    if (ok && v[i + j] == x) ismax |= 1u << (i + j);
                                }
                                if (writer) {
                                    uint32_t hw[8], lw[8];
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        split_pair(m[2 * j], m[2 * j + 1], hw[j], lw[j]);
                                    }
                                    stg256_b32(p.pool_hi + ppix * p.OC + c0 + i, hw);
                                    stg256_b32(p.pool_lo + ppix * p.OC + c0 + i, lw);
                                }
                            }
                        }
                        if (p.out_code) {
                            // ReLU / arg-max codes for the backward pass (Conv3x3TcArgs::ref_code).  The first maximum in
                            // scan order: the earlier positions of the window (lanes L^1, L^16, L^17 by position) do not
                            // hold the maximum.  (Warp-collective: shuffles before the `ok` test.)
                            uint32_t first = 0u;
                            if (p.pool_hi) {
                                const int kme = ((lane >> 4) & 1) * 2 + (lane & 1);
                                const uint32_t e1 = __shfl_xor_sync(0xffffffffu, ismax, 1);
                                const uint32_t e2 = __shfl_xor_sync(0xffffffffu, ismax, 16);
                                const uint32_t e3 = __shfl_xor_sync(0xffffffffu, ismax, 17);
                                uint32_t earlier = 0u;
                                if ((kme ^ 1) < kme) earlier |= e1;
                                if ((kme ^ 2) < kme) earlier |= e2;
                                if ((kme ^ 3) < kme) earlier |= e3;
                                first = ismax & ~earlier;
                            }
                            if (ok) {
                                uint32_t pos = 0u;
#pragma unroll
                                for (int i = 0; i < 32; ++i) pos |= (v[i] > 0.f ? 1u : 0u) << i;
                                uint32_t words[8];
#pragma unroll
                                for (int w = 0; w < 8; ++w) {
                                    const uint32_t pb = (pos >> (4 * w)) & 0xFu, fb = (first >> (4 * w)) & 0xFu;
                                    const uint32_t ps = (pb & 1u) | ((pb & 2u) << 7) | ((pb & 4u) << 14) | ((pb & 8u) << 21);
                                    const uint32_t fs4 = (fb & 1u) | ((fb & 2u) << 7) | ((fb & 4u) << 14) | ((fb & 8u) << 21);
                                    words[w] = ps | (fs4 << 1);
                                }
                                stg256_b32(p.out_code + pix * p.OC + c0, words);
                            }
                        }
                        if (ok) {
                        if (p.out_f32) {
                            float* op = p.out_f32 + pix * p.OC + c0;
#pragma unroll
                            for (int i = 0; i < 32; i += 8) stg256(op + i, v + i);
                        }
                        if (p.out_hi) {
                            __nv_bfloat16* hp = p.out_hi + pix * p.OC + c0;
                            __nv_bfloat16* lp = p.out_lo + pix * p.OC + c0;
#pragma unroll
                            for (int i = 0; i < 32; i += 16) {
                                uint32_t hw[8], lw[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    split_pair(v[i + 2 * j], v[i + 2 * j + 1], hw[j], lw[j]);
                                }
                                stg256_b32(hp + i, hw);
                                stg256_b32(lp + i, lw);
                            }
                        }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(mapa_u32(&t_empty[as], 0));    // the leader's barrier (its own warps too)
                else mbar_arrive(&t_empty[as]);
            }
            if (++as == 2) { as = 0; pt ^= 1; }
        }
        if (STATS) stats_flush();
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all();         // the peer's shared memory and barriers stay alive until both CTAs are done
    else __syncthreads();
    if (warp == 2) {
        if (PAIR) tmem_dealloc2(tmem_base, K::TMEM_COLS);
        else tmem_dealloc(tmem_base, K::TMEM_COLS);
    }
}

// ------------------------------------------------------------------ helpers kernels
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n4) {
    FS_PDL_ENTER();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 v = *reinterpret_cast<const float4*>(x + i * 4);
    float f[4] = {v.x, v.y, v.z, v.w};
    uint32_t h[2], l[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * j]), h1 = __float2bfloat16_rn(f[2 * j + 1]);
        __nv_bfloat16 l0 = __float2bfloat16_rn(f[2 * j] - __bfloat162float(h0));
        __nv_bfloat16 l1 = __float2bfloat16_rn(f[2 * j + 1] - __bfloat162float(h1));
        h[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        l[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    *reinterpret_cast<uint2*>(hi + i * 4) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2*>(lo + i * 4) = make_uint2(l[0], l[1]);
}

__global__ void pack_w3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, int Ci, int Co, int mode) {
    FS_PDL_ENTER();
    // output index: (((tap*CB + cb)*Nn + n)*64 + k)
    const int Kc = mode == 0 ? Ci : Co;        // reduction channels
    const int Nn = mode == 0 ? Co : Ci;        // output channels
    const int CB = Kc / 64;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = 9LL * Kc * Nn;
    if (i >= total) return;
    int k = (int)(i % 64);
    long long r = i / 64;
    int n = (int)(r % Nn); r /= Nn;
    int cb = (int)(r % CB);
    int tap = (int)(r / CB);
    int kh = tap / 3, kw = tap % 3;
    float v;
    if (mode == 0) v = w[(((long long)kh * 3 + kw) * Ci + cb * 64 + k) * Co + n];
    else v = w[(((long long)(2 - kh) * 3 + (2 - kw)) * Ci + n) * Co + cb * 64 + k];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void pack_taps_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, int T, int Kc, int Nn, int flip) {
    FS_PDL_ENTER();
    // output index: (((tap'*CB + cb)*Nn + n)*64 + k)  <-  w[tap][cb*64+k][n]
    const int CB = Kc / 64;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)T * Kc * Nn) return;
    int k = (int)(i % 64);
    long long r = i / 64;
    int n = (int)(r % Nn); r /= Nn;
    int cb = (int)(r % CB);
    int tp = (int)(r / CB);
    int tap = flip ? T - 1 - tp : tp;
    float v = w[((long long)tap * Kc + cb * 64 + k) * Nn + n];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

struct PackBatch { const float* w[16]; __nv_bfloat16* hi[16]; __nv_bfloat16* lo[16]; };

__global__ void pack_w3x3_batch_kernel(const PackBatch pb, int Ci, int Co, int mode) {
    FS_PDL_ENTER();
    const float* __restrict__ w = pb.w[blockIdx.y];
    __nv_bfloat16* __restrict__ hi = pb.hi[blockIdx.y];
    __nv_bfloat16* __restrict__ lo = pb.lo[blockIdx.y];
    const int Kc = mode == 0 ? Ci : Co, Nn = mode == 0 ? Co : Ci, CB = Kc / 64;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 9LL * Kc * Nn) return;
    int k = (int)(i % 64);
    long long r = i / 64;
    int n = (int)(r % Nn); r /= Nn;
    int cb = (int)(r % CB);
    int tap = (int)(r / CB);
    int kh = tap / 3, kw = tap % 3;
    float v = mode == 0 ? w[(((long long)kh * 3 + kw) * Ci + cb * 64 + k) * Co + n]
                        : w[(((long long)(2 - kh) * 3 + (2 - kw)) * Ci + n) * Co + cb * 64 + k];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

int make_act_map(CUtensorMap* tm, const __nv_bfloat16* base, int N, int H, int W, int C, int box_h) {
    EncodeTiledFn enc = get_encode_fn();
    FS_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)KB, (cuuint32_t)TW, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", N, H, W, C, (int)r);
    return 0;
}

// Space-to-depth view S[n, y, x, (p, q, c)] = T[n, 2y+p, 2x+q, c] of a plain [N, 2H, 2W, 32] tensor as a 5-D map
// {64 = (q,c) contiguous, W (x), 2 (p), H (y), N}: the 64-channel block cb of S is the row parity p.
int make_act_map_s2d(CUtensorMap* tm, const __nv_bfloat16* base, int N, int H, int W, int box_h) {
    EncodeTiledFn enc = get_encode_fn();
    FS_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    const cuuint64_t C0 = 32, FW = 2 * (cuuint64_t)W, FH = 2 * (cuuint64_t)H;
    cuuint64_t dims[5] = {64, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[4] = {2 * C0 * 2, FW * C0 * 2, 2 * FW * C0 * 2, FH * FW * C0 * 2};
    cuuint32_t box[5] = {(cuuint32_t)KB, (cuuint32_t)TW, 1, (cuuint32_t)box_h, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(s2d activation %dx%dx%d) failed: %d", N, H, W, (int)r);
    return 0;
}

int make_w_map(CUtensorMap* tm, const __nv_bfloat16* base, long long rows, int box_rows) {
    EncodeTiledFn enc = get_encode_fn();
    FS_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t dims[2] = {(cuuint64_t)KB, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)KB * 2};
    cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights rows=%lld) failed: %d", rows, (int)r);
    return 0;
}

int g_tc_pair = -1;     // CTA-pair (cta_group::2) kernel: -1 = read FS_TC_PAIR (default on), 0 off, 1 on
int g_tc_epi_warps = -1;   // lean launches: 16 epilogue warps (default) or 8 (FS_TC_EPI_WARPS=8 / set_tc_epi_warps: A/B switch)

template <int TH, int BN, int EPI, bool PAIR, int HALO>
int launch_cfg_s(const Conv3x3TcArgs& a, cudaStream_t st) {
    using K = Cfg<TH, BN, PAIR, HALO>;
    CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
    if (a.in_s2d) {
        FS_TRY(make_act_map_s2d(&tmA_hi, a.x.hi, a.N, a.H, a.W, TH + HALO));
        FS_TRY(make_act_map_s2d(&tmA_lo, a.x.lo, a.N, a.H, a.W, TH + HALO));
    } else {
        FS_TRY(make_act_map(&tmA_hi, a.x.hi, a.N, a.H, a.W, a.C, TH + HALO));
        FS_TRY(make_act_map(&tmA_lo, a.x.lo, a.N, a.H, a.W, a.C, TH + HALO));
    }
    const int taps = a.one_by_one ? 1 : (a.taps ? a.taps : 3);
    const int taps_w = a.taps_w ? a.taps_w : taps;
    const long long wrows = a.one_by_one ? (long long)(a.per_sample_w ? a.N : 1) * (a.C / KB) * a.OC
                                         : (long long)taps * taps_w * (a.C / KB) * a.OC;
    FS_TRY(make_w_map(&tmB_hi, a.w.hi, wrows, K::B_ROWS));
    FS_TRY(make_w_map(&tmB_lo, a.w.lo, wrows, K::B_ROWS));
    TcParams p;
    p.N = a.N; p.H = a.H; p.W = a.W; p.C = a.C; p.OH = a.OH; p.OW = a.OW; p.OC = a.OC; p.pad = a.pad;
    p.taps = taps; p.taps_w = taps_w; p.pad_x = a.taps_w ? a.pad_x : a.pad; p.in_s2d = a.in_s2d; p.out_d2s = a.out_d2s;
    p.w_sample_rows = a.one_by_one && a.per_sample_w ? (a.C / KB) * a.OC : 0;
    p.tilesX = cdiv(a.OW, TW); p.tilesY = cdiv(a.OH, TH); p.tilesN = a.OC / BN;
    const long long tps = (long long)p.tilesX * p.tilesY;
    p.total_tiles = (long long)a.N * (PAIR ? (tps + 1) / 2 : tps) * p.tilesN;
    p.bias = a.bias; p.addend = a.addend; p.ref = a.ref; p.relu = a.relu;
    p.pool_grad = a.pool_grad; p.ctarget = a.ctarget; p.cw2 = a.cw2;
    p.pool_hi = a.pool_split.hi; p.pool_lo = a.pool_split.lo;
    p.ref8 = a.ref_code; p.out_code = a.out_code;
    p.add_crop = a.add_crop; p.addH = a.addH; p.addW = a.addW;
    p.out_f32 = a.out_f32; p.out_hi = a.out_split.hi; p.out_lo = a.out_split.lo;
    p.stats = a.stats; p.stats_c = a.stats_c;
    FS_DYN_SMEM((conv3x3_tc_kernel<TH, BN, EPI, PAIR, HALO>), K::SMEM_BYTES);
    const int sms = num_sms();
    if (PAIR) {
        const int clusters = (int)(p.total_tiles < sms / 2 ? p.total_tiles : sms / 2);
        launch_k_cluster2((conv3x3_tc_kernel<TH, BN, EPI, PAIR, HALO>), dim3(2 * clusters), dim3(tc_threads(EPI)), K::SMEM_BYTES, st,
                          tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);
    } else {
        const int grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
        launch_k((conv3x3_tc_kernel<TH, BN, EPI, PAIR, HALO>), dim3(grid), dim3(tc_threads(EPI)), K::SMEM_BYTES, st, tmA_hi, tmA_lo,
                 tmB_hi, tmB_lo, p);
    }
    FS_LAUNCH_CHECK();
    return 0;
}

template <int TH, int BN, int HALO = 2>
int launch_cfg(const Conv3x3TcArgs& a, bool pair, cudaStream_t st) {
    if (g_tc_epi_warps < 0) {
        const char* e = getenv("FS_TC_EPI_WARPS");
        g_tc_epi_warps = (e && atoi(e) == 8) ? 8 : 16;
    }
    // (the fp32-reference variant with no reference tensor is exactly the lean epilogue on 8 warps.)  Measured: the 16-warp
    // epilogue pays on the 1x1 per-sample GEMMs of the Gram backward (all epilogue, 0.256 -> 0.240 ms per step) and costs
    // ~1 % on the 3x3 layers (fewer registers for the same work), so only the 1x1 launches take it.
    const int epi = a.stats ? 1 : ((a.ref || g_tc_epi_warps == 8 || !a.one_by_one) ? 2 : 0);
    if (pair) return epi == 1 ? launch_cfg_s<TH, BN, 1, true, HALO>(a, st)
                   : epi == 2 ? launch_cfg_s<TH, BN, 2, true, HALO>(a, st) : launch_cfg_s<TH, BN, 0, true, HALO>(a, st);
    return epi == 1 ? launch_cfg_s<TH, BN, 1, false, HALO>(a, st)
         : epi == 2 ? launch_cfg_s<TH, BN, 2, false, HALO>(a, st) : launch_cfg_s<TH, BN, 0, false, HALO>(a, st);
}

}  // namespace

int tc_make_map_nhwc(CUtensorMap* tm, const void* base, int elem_bytes, int N, int H, int W, int C, int boxC, int boxW,
                     int boxH, int swizzle_bytes) {
    EncodeTiledFn enc = get_encode_fn();
    FS_CHECK(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    FS_CHECK(boxC * elem_bytes == swizzle_bytes && (swizzle_bytes == 64 || swizzle_bytes == 128), "tc_make_map_nhwc: bad box");
    const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const cuuint64_t e = (cuuint64_t)elem_bytes;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * e, (cuuint64_t)W * C * e, (cuuint64_t)H * W * C * e};
    cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, dt, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FS_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(store map %dx%dx%dx%d, %d B) failed: %d", N, H, W, C, elem_bytes, (int)r);
    return 0;
}

bool conv3x3_tc_supported(int C, int OC, int W, int OW) {
    return C % 64 == 0 && OC % 64 == 0 && C >= 64 && OC >= 64 && W >= 1 && OW >= 1;
}

int launch_conv3x3_tc(const Conv3x3TcArgs& a, cudaStream_t st) {
    FS_CHECK(conv3x3_tc_supported(a.C, a.OC, a.W, a.OW), "conv3x3_tc: needs C%%64==0 and OC%%64==0 (C=%d OC=%d)", a.C, a.OC);
    FS_CHECK(a.x.hi && a.x.lo && a.w.hi && a.w.lo, "conv3x3_tc: NULL operand planes");
    FS_CHECK(a.out_f32 || a.out_split.hi || a.pool_split.hi, "conv3x3_tc: no output requested");
    FS_CHECK(!a.ref_code || (!a.ref && !a.ctarget && !a.out_d2s && !a.stats), "conv3x3_tc: ref_code excludes ref / ctarget / d2s / stats");
    FS_CHECK(!a.out_code || (!a.out_d2s && !a.stats && a.relu), "conv3x3_tc: out_code is for plain ReLU outputs");
    FS_CHECK((a.pool_split.hi == nullptr) == (a.pool_split.lo == nullptr) && !(a.pool_split.hi && a.out_d2s),
             "conv3x3_tc: pooled split output needs both planes and a plain layout");
    FS_CHECK((a.out_split.hi == nullptr) == (a.out_split.lo == nullptr), "conv3x3_tc: split output needs both planes");
    FS_CHECK(a.OH > 0 && a.OW > 0 && a.N > 0, "conv3x3_tc: empty output");
    FS_CHECK(a.taps == 0 || a.taps == 2 || a.taps == 3 || a.taps == 9, "conv3x3_tc: taps must be 2, 3 or 9");
    FS_CHECK(a.taps != 9 || (a.taps_w >= 1 && a.taps_w <= 3 && !a.in_s2d && !a.out_d2s && !a.one_by_one),
             "conv3x3_tc: the 9-row form takes 1..3 horizontal taps and plain layouts");
    FS_CHECK((!a.pool_grad && !a.ctarget) || ((a.ref || (a.ref_code && !a.ctarget)) && !a.out_d2s),
             "conv3x3_tc: pool_grad / ctarget need the ReLU reference tensor");
    // fused statistics: the real channel (output channel % stats_c) of element i of a warp's 32-channel chunk must
    // not depend on the chunk: stats_c divides 32, or one 64-channel tile whose two chunks go to different warps
    FS_CHECK(!a.stats || ((a.stats_c == 4 || a.stats_c == 8 || a.stats_c == 16 || a.stats_c == 32 ||
                           (a.stats_c == 64 && a.OC == 64)) && !a.bias && !a.relu && !a.addend && !a.ref),
             "conv3x3_tc: fused statistics are for raw transform-net conv outputs (16/32/64 real channels)");
    FS_CHECK(!a.in_s2d || (a.C == 128 && !a.one_by_one), "conv3x3_tc: the space-to-depth input view needs C == 128");
    FS_CHECK(!a.out_d2s || (a.OC == 128 && a.out_f32 && !a.out_split.hi && !a.bias && !a.addend && !a.ref && !a.relu),
             "conv3x3_tc: the depth-to-space store needs OC == 128 and a plain fp32 output");
    if (g_tc_pair < 0) {
        const char* e = getenv("FS_TC_PAIR");
        g_tc_pair = (e && e[0] == '0') ? 0 : 1;
    }
    const bool pair = g_tc_pair != 0;
    if (a.taps == 9)         // 9-row forms (x16 space-to-depth 9x9 layers): 8-row tiles under a 16-row slab
        return a.OC % 128 == 0 ? launch_cfg<8, 128, 8>(a, pair, st) : launch_cfg<8, 64, 8>(a, pair, st);
    bool tall = a.OH > 8;
    // 16-row tiles amortise the slab halo (18 rows loaded per 16) but small problems leave SMs idle or end on a
    // ragged last wave: estimate both tilings in units of one 8-row tile (x1.1 for the 10-rows-per-8 halo) and
    // take the cheaper.  E.g. the residual convs (200 vs 400 tiles on 148 SMs) and conv4_1's data gradient
    // (64 vs 128 tiles) run faster on 8-row tiles.  The CTA-pair kernel schedules PAIRS of tiles on 74 clusters.
    if (tall) {
        const long long bn = a.OC % 128 == 0 ? 128 : 64;
        const long long nn = a.OC / bn;
        long long t16 = (long long)cdiv(a.OW, TW) * cdiv(a.OH, 16), t8 = (long long)cdiv(a.OW, TW) * cdiv(a.OH, 8);   // per sample
        long long workers = num_sms();
        if (pair) { t16 = (t16 + 1) / 2; t8 = (t8 + 1) / 2; workers /= 2; }
        t16 *= nn * a.N; t8 *= nn * a.N;
        const double cost16 = 2.0 * (double)((t16 + workers - 1) / workers), cost8 = 1.1 * (double)((t8 + workers - 1) / workers);
        if (cost8 < 0.85 * cost16) tall = false;
        // the small 64->64 residual convs: 8-row tiles + their 3-deep slab ring also hide the L2 latency (measured)
        if (a.C == 64 && a.OC == 64 && !a.one_by_one && t16 * 2 < 3 * workers) tall = false;
        // MMA-bound layers (>= 256 channels) on small planes: 8-row tiles.  With 16-row tiles a CTA pair gets one or
        // two work units (conv4_x: 64 units on 74 pairs, conv3_x: 128), so the first load latency and the whole last
        // epilogue stay un-overlapped; twice as many half-size units hide them (measured: 4.390 -> 4.339 ms per step).
        // FS_TC_TH8 (bit mask, default 3): 1 = planes of <= 32 rows, 2 = planes of 33..64 rows, 4 = 65..128 rows.
        static int th8 = -1;
        if (th8 < 0) { const char* e = getenv("FS_TC_TH8"); th8 = e ? atoi(e) : 3; }
        if (!a.one_by_one && a.C >= 256 && (((th8 & 1) && a.OH <= 32) || ((th8 & 2) && a.OH > 32 && a.OH <= 64))) tall = false;
        if (!a.one_by_one && a.C >= 128 && (th8 & 4) && a.OH > 64 && a.OH <= 128) tall = false;
    }
    if (a.OC % 128 == 0) return tall ? launch_cfg<16, 128>(a, pair, st) : launch_cfg<8, 128>(a, pair, st);
    return tall ? launch_cfg<16, 64>(a, pair, st) : launch_cfg<8, 64>(a, pair, st);
}

void set_tc_pair(int on) { g_tc_pair = on ? 1 : 0; }
void set_tc_epi_warps(int warps) { g_tc_epi_warps = warps == 8 ? 8 : 16; }

int split_bf16(const float* x, SplitPtr out, long long n, cudaStream_t st) {
    FS_CHECK(n % 4 == 0, "split_bf16: n %% 4 != 0");
    long long n4 = n / 4;
    launch_k(split_bf16_kernel, dim3(cdiv(n4, 256)), dim3(256), 0, st, x, out.hi, out.lo, n4);
    FS_LAUNCH_CHECK();
    return 0;
}

int pack_w3x3_tc_batch(const float* const* w, const SplitPtr* out, int count, int Ci, int Co, int mode, cudaStream_t st) {
    FS_CHECK(count >= 1 && count <= 16, "pack_w3x3_tc_batch: 1..16 matrices per launch");
    FS_CHECK(Ci % 64 == 0 && Co % 64 == 0, "pack_w3x3_tc_batch: channels must be multiples of 64");
    PackBatch pb;
    for (int i = 0; i < count; ++i) { pb.w[i] = w[i]; pb.hi[i] = out[i].hi; pb.lo[i] = out[i].lo; }
    dim3 grid(cdiv(9LL * Ci * Co, 256), count);
    launch_k(pack_w3x3_batch_kernel, dim3(grid), dim3(256), 0, st, pb, Ci, Co, mode);
    FS_LAUNCH_CHECK();
    return 0;
}

int pack_taps_tc(const float* w, SplitPtr out, int T, int K, int Nn, int flip, cudaStream_t st) {
    FS_CHECK(K % 64 == 0 && Nn % 64 == 0 && T >= 1, "pack_taps_tc: K and N must be multiples of 64");
    long long total = (long long)T * K * Nn;
    launch_k(pack_taps_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, w, out.hi, out.lo, T, K, Nn, flip);
    FS_LAUNCH_CHECK();
    return 0;
}

int pack_w3x3_tc(const float* w, SplitPtr out, int Ci, int Co, int mode, cudaStream_t st) {
    FS_CHECK(Ci % 64 == 0 && Co % 64 == 0, "pack_w3x3_tc: channels must be multiples of 64");
    long long total = 9LL * Ci * Co;
    launch_k(pack_w3x3_kernel, dim3(cdiv(total, 256)), dim3(256), 0, st, w, out.hi, out.lo, Ci, Co, mode);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
