// FFMA (CUDA-core, exact fp32) implicit-GEMM kernels.
//
// These are the generic, any-shape convolution engines of the library: forward
// conv, data gradient (incl. stride-2 "fractionally strided"), the collapsed
// resize-conv (depth-to-space store / space-to-depth gather) and the per-sample
// Gram backward GEMM all run through igemm_kernel; weight gradients and the Gram
// forward (reductions over pixels) run through wgrad_kernel.  The tcgen05 path
// (conv3x3_tc.cu) takes over the 64-channel-multiple 3x3 layers; the kernels here
// remain the path for tiny-N / tiny-K layers (Cin=3, Cout=3..32) that cannot feed the
// tensor pipe.  Tile shapes are picked per output-channel count so
// that every thread keeps a >= 8x4 (or 8x8) register tile: 32 FMAs per 3 LDS.128.
//
// Reference semantics reproduced: tf.nn.conv2d NHWC/HWIO with TF SAME/VALID
// padding (reference im_transf_net.py:115-118, libs/vgg16.py:48-52).
#include "common.cuh"

namespace fs {

namespace {

struct RowInfo { int oy, ox; bool ok; int off; unsigned mh, mw; };

__device__ __forceinline__ float4 ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

// Gather one float4 (4 consecutive channels of one tap) of the implicit A matrix.
__device__ __forceinline__ float4 gather_a(const IGemmArgs& a, const float* in_n, int oy, int ox,
                                           bool row_ok, int kh, int kw, int c, bool k_ok) {
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!row_ok || !k_ok) return z;
    int iy, ix;
    if (a.gather == 0) {
        iy = oy * a.stride - a.pad_t + kh;
        ix = ox * a.stride - a.pad_l + kw;
    } else {
        int ny = oy + a.pad_t - kh, nx = ox + a.pad_l - kw;
        if (ny < 0 || nx < 0) return z;
        iy = ny / a.stride; ix = nx / a.stride;
        if (iy * a.stride != ny || ix * a.stride != nx) return z;
    }
    if (iy < 0 || iy >= a.H || ix < 0 || ix >= a.W) return z;
    long long off;
    if (a.in_mode == 0) {
        off = ((long long)iy * a.W + ix) * a.C + c;
    } else {
        int C0 = a.C >> 2;
        int pq = c / C0, co = c - pq * C0;
        off = ((long long)(2 * iy + (pq >> 1)) * (2 * a.W) + 2 * ix + (pq & 1)) * C0 + co;
    }
    return ldg4(in_n + off);
}

// BM x BN output tile, BK reduction slice, TM x TN register tile, NT threads.
template <int BM, int BN, int TM, int TN, int NT, int BK>
__global__ void __launch_bounds__(NT) igemm_kernel(const IGemmArgs a) {
    FS_PDL_ENTER();
    constexpr int TXN = BN / TN;
    constexpr int TYN = NT / TXN;
    static_assert(TYN * TM == BM && TXN * TN == BN, "tile/thread mismatch");
    constexpr int K4 = BK / 4;                        // float4 per A row slice
    static_assert(NT % K4 == 0, "thread count must be a multiple of BK/4");
    constexpr int RSTEP = NT / K4;                    // rows covered per load round
    static_assert(BM % RSTEP == 0, "BM must be a multiple of NT/(BK/4)");
    constexpr int A_F4 = BM / RSTEP;
    constexpr int B_TOT = BK * BN / 4;
    constexpr int B_F4 = (B_TOT + NT - 1) / NT;
    constexpr int NG = TN / 4;                        // float4 column groups per thread
    constexpr int GSTRIDE = BN / NG;

    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int t = threadIdx.x;
    const int n = blockIdx.z;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int M = a.OH * a.OW;
    const int Ktot = a.KH * a.KW * a.C;
    const int C4 = a.C >> 2;
    const int nk = (Ktot + BK - 1) / BK;
    const float* in_n = a.in + (long long)n * a.in_bs;
    const float* w_n = a.w + (long long)n * a.w_bs;

    // rows this thread gathers (its float4 column k4 is the same for all of them)
    const int k4 = t % K4;
    const int r0 = t / K4;
    RowInfo rows[A_F4];
#pragma unroll
    for (int i = 0; i < A_F4; ++i) {
        int m = m0 + r0 + i * RSTEP;
        rows[i].ok = m < M;
        int mm = rows[i].ok ? m : 0;
        rows[i].oy = mm / a.OW;
        rows[i].ox = mm - rows[i].oy * a.OW;
    }
    // Fast gather (plain NHWC input, forward conv or stride-1 data gradient): the element offset is
    // row_offset + tap_offset and the bounds test is two precomputed per-row bit masks.
    const bool fast = a.in_mode == 0 && (a.gather == 0 || a.stride == 1);
    const int sgn = a.gather == 0 ? 1 : -1;
    if (fast) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            int by = a.gather == 0 ? rows[i].oy * a.stride - a.pad_t : rows[i].oy + a.pad_t;
            int bx = a.gather == 0 ? rows[i].ox * a.stride - a.pad_l : rows[i].ox + a.pad_l;
            unsigned mh = 0, mw = 0;
            for (int k = 0; k < a.KH; ++k) { int y = by + sgn * k; if (y >= 0 && y < a.H) mh |= 1u << k; }
            for (int k = 0; k < a.KW; ++k) { int x = bx + sgn * k; if (x >= 0 && x < a.W) mw |= 1u << k; }
            rows[i].off = (by * a.W + bx) * a.C;
            rows[i].mh = rows[i].ok ? mh : 0u;
            rows[i].mw = mw;
        }
    }

    float4 ra[A_F4];
    float4 rb[B_F4];

    auto load_tile = [&](int kt) {
        int kg4 = kt * K4 + k4;
        bool k_ok = kg4 * 4 < Ktot;
        int tap = kg4 / C4;
        int c = (kg4 - tap * C4) * 4;
        int kh = tap / a.KW;
        int kw = tap - kh * a.KW;
        if (fast) {
            const int tapoff = sgn * (kh * a.W + kw) * a.C + c;
#pragma unroll
            for (int i = 0; i < A_F4; ++i) {
                bool v = k_ok && (((rows[i].mh >> kh) & (rows[i].mw >> kw)) & 1u);
                ra[i] = v ? ldg4(in_n + rows[i].off + tapoff) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
#pragma unroll
            for (int i = 0; i < A_F4; ++i)
                ra[i] = gather_a(a, in_n, rows[i].oy, rows[i].ox, rows[i].ok, kh, kw, c, k_ok);
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            int f = t + i * NT;
            rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < B_TOT) {
                int bk = f / (BN / 4), bn4 = f - bk * (BN / 4);
                int kg = kt * BK + bk, col = n0 + bn4 * 4;
                if (kg < Ktot && col < a.OC) rb[i] = ldg4(w_n + (long long)kg * a.OC + col);
            }
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_F4; ++i) {
            int r = r0 + i * RSTEP;
            As[buf][k4 * 4 + 0][r] = ra[i].x;
            As[buf][k4 * 4 + 1][r] = ra[i].y;
            As[buf][k4 * 4 + 2][r] = ra[i].z;
            As[buf][k4 * 4 + 3][r] = ra[i].w;
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            int f = t + i * NT;
            if (f < B_TOT) {
                int bk = f / (BN / 4), bn4 = f - bk * (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][bk][bn4 * 4]) = rb[i];
            }
        }
    };

    const int tx = t % TXN, ty = t / TXN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float ar[TM], br[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) ar[i] = As[buf][kk][ty * TM + i];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * GSTRIDE + tx * 4]);
                br[g * 4 + 0] = v.x; br[g * 4 + 1] = v.y; br[g * 4 + 2] = v.z; br[g * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        if (kt + 1 < nk) store_tile(buf ^ 1);
        __syncthreads();
    }

    // ---------------------------------------------------------------- epilogue
    float* out_n = a.out + (long long)n * a.out_bs;
    const float* ref_n = a.ref ? a.ref + (long long)n * a.out_bs : nullptr;
    const float* add_n = a.addend ? a.addend + (long long)n * a.add_bs : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
        int oy = m / a.OW, ox = m - oy * a.OW;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            int col = n0 + g * GSTRIDE + tx * 4;
            if (col >= a.OC) continue;
            float4 v = make_float4(acc[i][g * 4 + 0], acc[i][g * 4 + 1], acc[i][g * 4 + 2],
                                   acc[i][g * 4 + 3]);
            if (a.bias) {
                float4 b = ldg4(a.bias + col);
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (add_n) {
                int ay = oy - a.add_crop, ax = ox - a.add_crop;
                if (ay >= 0 && ay < a.addH && ax >= 0 && ax < a.addW) {
                    float4 d = *reinterpret_cast<const float4*>(
                        add_n + ((long long)ay * a.addW + ax) * a.OC + col);
                    v.x += d.x; v.y += d.y; v.z += d.z; v.w += d.w;
                }
            }
            if (a.relu) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f);
                v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            }
            long long off;
            if (a.out_mode == 0) {
                off = (long long)m * a.OC + col;
            } else {
                int C0 = a.OC >> 2;
                int pq = col / C0, co = col - pq * C0;
                off = ((long long)(2 * oy + (pq >> 1)) * (2 * a.OW) + 2 * ox + (pq & 1)) * C0 + co;
            }
            if (ref_n) {
                float4 r = *reinterpret_cast<const float4*>(ref_n + off);
                v.x = r.x > 0.f ? v.x : 0.f; v.y = r.y > 0.f ? v.y : 0.f;
                v.z = r.z > 0.f ? v.z : 0.f; v.w = r.w > 0.f ? v.w : 0.f;
            }
            *reinterpret_cast<float4*>(out_n + off) = v;
        }
    }
}

template <int BM, int BN, int TM, int TN, int NT, int BK>
int launch_cfg(const IGemmArgs& a, cudaStream_t st) {
    int M = a.OH * a.OW;
    dim3 grid(cdiv(M, BM), cdiv(a.OC, BN), a.N);
    launch_k((igemm_kernel<BM, BN, TM, TN, NT, BK>), dim3(grid), dim3(NT), 0, st, a);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int launch_igemm(const IGemmArgs& a, cudaStream_t st) {
    FS_CHECK(a.C % 4 == 0 && a.OC % 4 == 0, "igemm: channel counts must be multiples of 4 (C=%d OC=%d)", a.C, a.OC);
    FS_CHECK(a.in_mode == 0 || (a.C % 16 == 0), "igemm: s2d gather needs C%%16==0");
    FS_CHECK(a.out_mode == 0 || (a.OC % 16 == 0), "igemm: d2s store needs OC%%16==0");
    FS_CHECK(a.N > 0 && a.N <= 65535, "igemm: batch %d out of range", a.N);
    FS_CHECK(a.stride >= 1, "igemm: bad stride");
    FS_CHECK(a.KH <= 16 && a.KW <= 16, "igemm: kernel larger than 16x16");
    FS_CHECK((long long)a.H * a.W * a.C < (1LL << 30), "igemm: per-sample input too large for 32-bit offsets");
    if (a.OH <= 0 || a.OW <= 0) return 0;
    if (a.OC >= 128) return launch_cfg<128, 128, 8, 8, 256, 16>(a, st);
    if (a.OC > 32) return launch_cfg<128, 64, 8, 4, 256, 16>(a, st);
    // small-N layers are bound by the im2col gather, not by FMA issue: many light threads
    // (high occupancy to hide the gather latency) beat big register tiles here (measured).
    if (a.OC > 16) return launch_cfg<128, 32, 4, 4, 256, 16>(a, st);
    if (a.OC > 4) return launch_cfg<128, 16, 2, 4, 256, 16>(a, st);
    return launch_cfg<256, 4, 1, 4, 256, 16>(a, st);
}

// =====================================================================
// Reduction-over-pixels GEMM: weight gradient / Gram forward
//   out[g][k, j] = scale * sum_{pix} A[pix, k] * B[pix, j]
// Deterministic: each split writes a partial tile, wgrad_reduce_kernel sums the
// splits in a fixed order.
// =====================================================================
namespace {

struct WGeom {
    int M, Ktot, C4, groups, splits, pix_per_group, chunk;
};

// WK x WN output tile (k rows x out channels), WP pixels per smem step, TK x 4 register tile.
template <int WK, int WN, int TK, int WP, int NT>
__global__ void __launch_bounds__(NT) wgrad_kernel(const WGradArgs a, const WGeom g) {
    FS_PDL_ENTER();
    constexpr int TXN = WN / 4;
    constexpr int TYN = NT / TXN;
    static_assert(TYN * TK == WK, "tile/thread mismatch");
    constexpr int K4 = WK / 4;                               // float4 per A row
    constexpr int KPT = (K4 + NT - 1) / NT;                  // distinct k4 columns per thread
    constexpr int PSTEP = NT >= K4 ? NT / K4 : 1;            // pixel rows covered per load round
    static_assert(WP % PSTEP == 0, "WP must be a multiple of the pixel step");
    constexpr int PPT = WP / PSTEP;                          // pixel rows per thread per k4 column
    constexpr int B_TOT = WP * WN / 4;
    constexpr int B_F4 = (B_TOT + NT - 1) / NT;

    __shared__ __align__(16) float As[2][WP][WK + 4];
    __shared__ __align__(16) float Bs[2][WP][WN + 4];
    const int t = threadIdx.x;
    const int k0 = blockIdx.x * WK, j0 = blockIdx.y * WN;
    const int grp = blockIdx.z / g.splits, sp = blockIdx.z - grp * g.splits;
    const int pbeg = sp * g.chunk;
    const int pend = min(pbeg + g.chunk, g.pix_per_group);
    const int nsteps = pend > pbeg ? (pend - pbeg + WP - 1) / WP : 0;
    const int samples_per_group = g.pix_per_group / g.M;

    // this thread's k columns (fixed for the whole loop): decode (tap, channel) once
    int kc[KPT], kkh[KPT], kkw[KPT]; bool kok[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        int kg4 = (k0 >> 2) + (t % K4) + j * NT;
        bool in_tile = ((t % K4) + j * NT) < K4;
        kok[j] = in_tile && kg4 * 4 < g.Ktot;
        int tap = kg4 / g.C4;
        kc[j] = (kg4 - tap * g.C4) * 4;
        kkh[j] = tap / a.KW;
        kkw[j] = tap - kkh[j] * a.KW;
    }
    const int p0 = NT >= K4 ? t / K4 : 0;

    // Pixel trackers: every thread follows PPT (A side) + B_F4 (dy side) pixels that advance by WP per step;
    // (sample, oy, ox) are updated incrementally instead of two integer divisions per pixel per step.
    struct Trk { int pix, ns, oy, ox; };
    auto trk_init = [&](Trk& k, int pix) {
        k.pix = pix;
        int pp = pix < pend ? pix : pbeg;
        k.ns = pp / g.M;
        int m = pp - k.ns * g.M;
        k.oy = m / a.OW;
        k.ox = m - k.oy * a.OW;
    };
    auto trk_adv = [&](Trk& k) {
        k.pix += WP;
        k.ox += WP;
        while (k.ox >= a.OW) {
            k.ox -= a.OW;
            if (++k.oy >= a.OH) { k.oy = 0; ++k.ns; }
        }
    };
    Trk ta[PPT], tb[B_F4];
#pragma unroll
    for (int i = 0; i < PPT; ++i) trk_init(ta[i], pbeg + p0 + i * PSTEP);
    // dy-side constants of this thread (its column group never changes)
    int bn4[B_F4]; bool bok[B_F4]; int bpq[B_F4], bco[B_F4];
#pragma unroll
    for (int i = 0; i < B_F4; ++i) {
        int f = t + i * NT;
        int pb = f / (WN / 4);
        bn4[i] = f - pb * (WN / 4);
        int jcol = j0 + bn4[i] * 4;
        bok[i] = f < B_TOT && jcol < a.OC;
        int C0 = a.OC >> 2;
        bpq[i] = a.dy_mode ? jcol / C0 : 0;
        bco[i] = a.dy_mode ? jcol - bpq[i] * C0 : jcol;
        trk_init(tb[i], pbeg + pb);
    }
    int apq[KPT], aco[KPT];
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        int C0 = a.C >> 2;
        apq[j] = a.in_mode ? kc[j] / C0 : 0;
        aco[j] = a.in_mode ? kc[j] - apq[j] * C0 : kc[j];
    }

    float4 ra[KPT][PPT], rb[B_F4];
    auto load_step = [&](int) {
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const bool pok = ta[i].pix < pend;
            const float* in_n = a.in + (long long)(grp * samples_per_group + ta[i].ns) * a.in_bs;
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (pok && kok[j]) {
                    int iy = ta[i].oy * a.stride - a.pad_t + kkh[j], ix = ta[i].ox * a.stride - a.pad_l + kkw[j];
                    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
                        long long off;
                        if (a.in_mode == 0) off = ((long long)iy * a.W + ix) * a.C + kc[j];
                        else off = ((long long)(2 * iy + (apq[j] >> 1)) * (2 * a.W) + 2 * ix + (apq[j] & 1)) * (a.C >> 2) + aco[j];
                        v = ldg4(in_n + off);
                    }
                }
                ra[j][i] = v;
            }
            trk_adv(ta[i]);
        }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            rb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bok[i] && tb[i].pix < pend) {
                long long off;
                if (a.dy_mode == 0) off = ((long long)tb[i].oy * a.OW + tb[i].ox) * a.OC + bco[i];
                else off = ((long long)(2 * tb[i].oy + (bpq[i] >> 1)) * (2 * a.OW) + 2 * tb[i].ox + (bpq[i] & 1)) * (a.OC >> 2) + bco[i];
                rb[i] = ldg4(a.dy + (long long)(grp * samples_per_group + tb[i].ns) * a.dy_bs + off);
            }
            trk_adv(tb[i]);
        }
    };
    auto store_step = [&](int buf) {
#pragma unroll
        for (int i = 0; i < PPT; ++i)
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                int kq = (t % K4) + j * NT;
                if (kq < K4) *reinterpret_cast<float4*>(&As[buf][p0 + i * PSTEP][kq * 4]) = ra[j][i];
            }
#pragma unroll
        for (int i = 0; i < B_F4; ++i) {
            int f = t + i * NT;
            if (f < B_TOT) {
                int p = f / (WN / 4), n4 = f - p * (WN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][p][n4 * 4]) = rb[i];
            }
        }
    };

    const int tx = t % TXN, ty = t / TXN;
    float acc[TK][4];
#pragma unroll
    for (int i = 0; i < TK; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (nsteps > 0) {
        load_step(0);
        store_step(0);
    }
    __syncthreads();
    for (int s = 0; s < nsteps; ++s) {
        const int buf = s & 1;
        if (s + 1 < nsteps) load_step(s + 1);
#pragma unroll
        for (int p = 0; p < WP; ++p) {
            float ar[TK];
#pragma unroll
            for (int i = 0; i < TK; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&As[buf][p][ty * TK + i]);
                ar[i] = v.x; ar[i + 1] = v.y; ar[i + 2] = v.z; ar[i + 3] = v.w;
            }
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][p][tx * 4]);
            float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TK; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
        if (s + 1 < nsteps) store_step(buf ^ 1);
        __syncthreads();
    }
    // partial[z][k][j]
    float* part = a.partial + (long long)blockIdx.z * g.Ktot * a.OC;
#pragma unroll
    for (int i = 0; i < TK; ++i) {
        int k = k0 + ty * TK + i;
        int col = j0 + tx * 4;
        if (k < g.Ktot && col < a.OC)
            *reinterpret_cast<float4*>(part + (long long)k * a.OC + col) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
}

// Sums the split partials of one float4 column in a fixed order: 32 split-lanes each take splits
// y, y+32, ... (independent loads in flight, 8 columns = one 128-byte line per split row), then lane 0 of the
// column folds the 32 partial sums in order 0..31.  8 columns per CTA: a [K, OC] tile of a few thousand floats
// still spreads over >100 CTAs (the first version used 32 columns x 8 lanes and ran 58 us on 36 CTAs).
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, long long tile_elems,
                    int splits, float scale) {
    FS_PDL_ENTER();
    __shared__ float4 sm[32][8];
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const long long n4 = tile_elems / 4;
    const long long e = (long long)blockIdx.x * 8 + tx;
    const long long grp = blockIdx.y;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < n4) {
        const float4* p = reinterpret_cast<const float4*>(partial) + grp * splits * n4 + e;
#pragma unroll 4
        for (int k = ty; k < splits; k += 32) {
            float4 v = p[(long long)k * n4];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
    }
    sm[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && e < n4) {
#pragma unroll
        for (int k = 1; k < 32; ++k) {
            float4 v = sm[k][tx];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        s.x *= scale; s.y *= scale; s.z *= scale; s.w *= scale;
        reinterpret_cast<float4*>(out)[grp * n4 + e] = s;
    }
}

struct WCfg { int WK, WN, WP; };
WCfg pick_wcfg(int OC) {
    if (OC > 32) return {64, 64, 16};
    if (OC > 16) return {64, 32, 16};          // 17..32 output channels: a 64-wide tile would be half padding
    if (OC > 4) return {256, 16, 16};
    return {512, 4, 8};
}

int choose_splits(int tiles, int groups, int pix_per_group, int WP) {
    const int target = 148 * 4;
    int s = (target + tiles * groups - 1) / (tiles * groups);
    int max_s = pix_per_group / (WP * 8);
    if (max_s < 1) max_s = 1;
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    if (s > 256) s = 256;
    return s;
}

}  // namespace

long long wgrad_partial_floats(int K, int OC, int groups) {
    // upper bound used by workspace planning (choose_splits only ever lowers the split count)
    WCfg c = pick_wcfg(OC);
    long long tiles = (long long)cdiv(K, c.WK) * cdiv(OC, c.WN);
    long long s = (148 * 4 + tiles * groups - 1) / (tiles * groups);
    if (s < 1) s = 1;
    if (s > 256) s = 256;
    return s * groups * (long long)K * OC;
}

int launch_wgrad(const WGradArgs& a, cudaStream_t st) {
    FS_CHECK(a.C % 4 == 0 && a.OC % 4 == 0, "wgrad: channel counts must be multiples of 4");
    WCfg c = pick_wcfg(a.OC);
    WGeom g;
    g.M = a.OH * a.OW;
    g.Ktot = a.KH * a.KW * a.C;
    g.C4 = a.C >> 2;
    g.groups = a.per_sample ? a.N : 1;
    g.pix_per_group = a.per_sample ? g.M : a.N * g.M;
    int tiles = cdiv(g.Ktot, c.WK) * cdiv(a.OC, c.WN);
    g.splits = choose_splits(tiles, g.groups, g.pix_per_group, c.WP);
    g.chunk = cdiv(cdiv(g.pix_per_group, g.splits), c.WP) * c.WP;
    long long need = (long long)g.groups * g.splits * g.Ktot * a.OC;
    FS_CHECK(need <= a.partial_cap, "wgrad: partial workspace too small (%lld > %lld floats)", need, a.partial_cap);
    FS_CHECK((long long)g.groups * g.splits <= 65535, "wgrad: too many z blocks");
    dim3 grid(cdiv(g.Ktot, c.WK), cdiv(a.OC, c.WN), g.groups * g.splits);
    if (c.WN == 64) launch_k((wgrad_kernel<64, 64, 8, 16, 128>), dim3(grid), dim3(128), 0, st, a, g);
    else if (c.WN == 32) launch_k((wgrad_kernel<64, 32, 4, 16, 128>), dim3(grid), dim3(128), 0, st, a, g);   // 4 warps: the stride-2 gather is latency-bound
    else if (c.WN == 16) launch_k((wgrad_kernel<256, 16, 8, 16, 128>), dim3(grid), dim3(128), 0, st, a, g);
    else launch_k((wgrad_kernel<512, 4, 8, 8, 64>), dim3(grid), dim3(64), 0, st, a, g);
    FS_LAUNCH_CHECK();
    long long tile_elems = (long long)g.Ktot * a.OC;
    launch_k(wgrad_reduce_kernel, dim3(dim3((unsigned)cdiv(tile_elems / 4, 8), g.groups)), dim3(256), 0, st, 
        a.partial, a.out, tile_elems, g.splits, a.scale);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
