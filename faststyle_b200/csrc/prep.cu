// Table-driven weight preparation: every per-step layout transform of the transform-net weights (channel padding,
// resize-conv / stride-2 collapses, pixel pairing, tap transposition, split-bf16 packing for the tensor path) and of
// the weight gradients on the way back (un-pairing, collapse adjoints, un-padding) is a JOB; one launch runs all
// jobs of a dependency phase (blockIdx.y = job).  The 37 short launches per train step of the one-kernel-per-
// transform version (0.27 ms of a 5.4 ms step, launch-bound) become 4 + 2.
// The index arithmetic of each job kind is the same as the stand-alone kernels in ops.cu / conv3x3_tc.cu /
// direct9x9.cu (which remain for the single-op C-ABI entry points).
#include <cuda_bf16.h>
#include "prep.cuh"

namespace fs {

namespace {

// R_0 = [[1,1,1],[0,0,0]],  R_1 = [[1,1,0],[0,0,1]]   (rows a, cols kh) - resize-conv collapse
__device__ __forceinline__ bool rsel(int p, int a, int k) {
    return p == 0 ? (a == 0) : (a == 0 ? k < 2 : k == 2);
}

__device__ __forceinline__ void split_store(__nv_bfloat16* hi, __nv_bfloat16* lo, long long i, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__device__ __forceinline__ void run_job(const PrepJob& j, long long i) {
    const float* __restrict__ src = j.src;
    float* __restrict__ dst = j.dst;
    switch (j.kind) {
    case PJ_PAD: {            // a=T b=Ci c=Co d=Cip e=Cop: dst[T][Cip][Cop] <- src[T][Ci][Co], zero fill
        const int Ci = j.b, Co = j.c, Cip = j.d, Cop = j.e;
        int o = (int)(i % Cop);
        long long r = i / Cop;
        int ci = (int)(r % Cip);
        int t = (int)(r / Cip);
        dst[i] = (ci < Ci && o < Co) ? src[((long long)t * Ci + ci) * Co + o] : 0.f;
        break;
    }
    case PJ_UNPAD: {          // dst[T][Ci][Co] <- src[T][Cip][Cop]
        const int Ci = j.b, Co = j.c, Cip = j.d, Cop = j.e;
        int o = (int)(i % Co);
        long long r = i / Co;
        int ci = (int)(r % Ci);
        int t = (int)(r / Ci);
        dst[i] = src[((long long)t * Cip + ci) * Cop + o];
        break;
    }
    case PJ_TRANSPOSE: {      // a=T b=Ci c=Co: dst[t][o][ci] = src[t][ci][o]
        const int Ci = j.b, Co = j.c;
        int ci = (int)(i % Ci);
        long long r = i / Ci;
        int o = (int)(r % Co);
        int t = (int)(r / Co);
        dst[i] = src[((long long)t * Ci + ci) * Co + o];
        break;
    }
    case PJ_FLIP_TRANSPOSE: { // a=T b=Ci c=Co: dst[T-1-t][o][ci] = src[t][ci][o]
        const int T = j.a, Ci = j.b, Co = j.c;
        int ci = (int)(i % Ci);
        long long r = i / Ci;
        int co = (int)(r % Co);
        int tf = (int)(r / Co);
        dst[i] = src[((long long)(T - 1 - tf) * Ci + ci) * Co + co];
        break;
    }
    case PJ_UPCONV_COLLAPSE: {        // a=Ci b=Co: Wc[a][b][ci][(p*2+q)*Co + co] = (R_p W R_q^T)[a][b]
        const int Ci = j.a, Co = j.b;
        int co = (int)(i % Co);
        long long r = i / Co;
        int pq = (int)(r % 4); r /= 4;
        int ci = (int)(r % Ci); r /= Ci;
        int b = (int)(r % 2), a = (int)(r / 2);
        int p = pq >> 1, q = pq & 1;
        float s = 0.f;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw)
                if (rsel(p, a, kh) && rsel(q, b, kw)) s += src[(((long long)kh * 3 + kw) * Ci + ci) * Co + co];
        dst[i] = s;
        break;
    }
    case PJ_UPCONV_COLLAPSE_GRAD: {   // a=Ci b=Co: adjoint of the above, dst[3][3][Ci][Co]
        const int Ci = j.a, Co = j.b;
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
                            s += src[((((long long)a * 2 + b) * Ci + ci) * 4 + (p * 2 + q)) * Co + co];
        dst[i] = s;
        break;
    }
    case PJ_S2_DGRAD_COLLAPSE: {      // a=Ci b=Co: Wd[a][b][co][(p*2+q)*Ci + ci]
        const int Ci = j.a, Co = j.b;
        int ci = (int)(i % Ci);
        long long r = i / Ci;
        int pq = (int)(r % 4); r /= 4;
        int co = (int)(r % Co); r /= Co;
        int b = (int)(r % 2), a = (int)(r / 2);
        int p = pq >> 1, q = pq & 1;
        int kh = p == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : -1);
        int kw = q == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : -1);
        dst[i] = (kh >= 0 && kw >= 0) ? src[(((long long)kh * 3 + kw) * Ci + ci) * Co + co] : 0.f;
        break;
    }
    case PJ_S2_FWD_COLLAPSE: {        // a=Ci b=Co: Wf[a][b][(p*2+q)*Ci + ci][co] = W[2a+p][2b+q]
        const int Ci = j.a, Co = j.b;
        int co = (int)(i % Co);
        long long r = i / Co;
        int ci = (int)(r % Ci); r /= Ci;
        int pq = (int)(r % 4); r /= 4;
        int b = (int)(r % 2), a = (int)(r / 2);
        int kh = 2 * a + (pq >> 1), kw = 2 * b + (pq & 1);
        dst[i] = (kh < 3 && kw < 3) ? src[(((long long)kh * 3 + kw) * Ci + ci) * Co + co] : 0.f;
        break;
    }
    case PJ_S2_FWD_COLLAPSE_GRAD: {   // a=Ci b=Co: dW[kh][kw][ci][co] = dWf[kh>>1][kw>>1][((kh&1)*2+(kw&1))*Ci + ci][co]
        const int Ci = j.a, Co = j.b;
        int co = (int)(i % Co);
        long long r = i / Co;
        int ci = (int)(r % Ci); r /= Ci;
        int kw = (int)(r % 3), kh = (int)(r / 3);
        int a = kh >> 1, b = kw >> 1, pq = (kh & 1) * 2 + (kw & 1);
        dst[i] = src[((((long long)a * 2 + b) * 4 + pq) * Ci + ci) * Co + co];
        break;
    }
    case PJ_PAIR: {           // a=K0 b=N0 c=kmode d=gather: W [2][2][K0][N0] -> Wp [2][2][2K0][2N0] (see ops.cu pair_taps_kernel)
        const int K0 = j.a, N0 = j.b, kmode = j.c, gather = j.d;
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
        dst[i] = (b == 0 || b == 1) ? src[(((long long)a * 2 + b) * K0 + k) * N0 + n] : 0.f;
        break;
    }
    case PJ_UNPAIR: {         // a=K0 b=N0 c=kmode d=n_s2d: adjoint of PJ_PAIR (correlation form), dst [2][2][K0][N0]
        const int K0 = j.a, N0 = j.b, kmode = j.c, n_s2d = j.d;
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
                s += src[(((long long)a * 2 + kwp) * (2 * K0) + kp) * (2 * N0) + col];
            }
        dst[i] = s;
        break;
    }
    case PJ_PACK_TAPS: {      // a=T b=Kc c=Nn d=flip: B[tap'][cb][n][k] = w[tap][cb*64+k][n] as split planes
        const int T = j.a, Kc = j.b, Nn = j.c, flip = j.d;
        const int CB = Kc / 64;
        int k = (int)(i % 64);
        long long r = i / 64;
        int n = (int)(r % Nn); r /= Nn;
        int cb = (int)(r % CB);
        int tp = (int)(r / CB);
        int tap = flip ? T - 1 - tp : tp;
        split_store(j.hi, j.lo, i, src[((long long)tap * Kc + cb * 64 + k) * Nn + n]);
        break;
    }
    case PJ_PACK_W3X3: {      // a=Ci b=Co c=mode (0 forward, 1 data gradient): see tc.cuh pack_w3x3_tc
        const int Ci = j.a, Co = j.b, mode = j.c;
        const int Kc = mode == 0 ? Ci : Co, Nn = mode == 0 ? Co : Ci, CB = Kc / 64;
        int k = (int)(i % 64);
        long long r = i / 64;
        int n = (int)(r % Nn); r /= Nn;
        int cb = (int)(r % CB);
        int tap = (int)(r / CB);
        int kh = tap / 3, kw = tap % 3;
        float v = mode == 0 ? src[(((long long)kh * 3 + kw) * Ci + cb * 64 + k) * Co + n]
                            : src[(((long long)(2 - kh) * 3 + (2 - kw)) * Ci + n) * Co + cb * 64 + k];
        split_store(j.hi, j.lo, i, v);
        break;
    }
    case PJ_IN15: {           // dst[0..7] = {g0,g1,g2,0, b0,b1,b2,0}: 4-channel staging of the last layer's IN vectors
        const int c = (int)(i & 3);
        dst[i] = c < 3 ? (i < 4 ? src[c] : j.src2[c]) : 0.f;
        break;
    }
    case PJ_X16: {            // a=A b=B (src W[9][9][A][B]) c=KP d=NP e=mode: the 9x9 stride-1 SAME conv as a 9x2-tap conv over
                              // 16-pixel groups: dst[kh][t][dxi*KP + k][dxo*NP + n], kw = 16 t + dxi - dxo (see Engine::tc9)
        // e bit 1: the x8 form - 8-pixel output groups over 16-pixel input windows (one horizontal tap, N = 8 * NP):
        // half the MMA work of the x16 form when the input side has 4 channels (16 px x 4 ch = one 64-channel K block)
        const int A = j.a, B = j.b, KP = j.c, NP = j.d, mode = j.e & 1, x8 = (j.e >> 1) & 1;
        const int PXO = x8 ? 8 : 16, TWn = x8 ? 1 : 2;
        int np, kp, tap;
        if (j.hi) {           // straight into the packed tensor-path layout B[tap][cb][n][k] (split planes)
            const int Nn = PXO * NP, CB = 16 * KP / 64;
            int k = (int)(i % 64);
            long long r = i / 64;
            np = (int)(r % Nn); r /= Nn;
            int cb = (int)(r % CB);
            tap = (int)(r / CB);
            kp = cb * 64 + k;
        } else {
            np = (int)(i % (PXO * NP));
            long long r = i / (PXO * NP);
            kp = (int)(r % (16 * KP));
            tap = (int)(r / (16 * KP));
        }
        int kh = tap / TWn, t = tap - kh * TWn;
        int dxo = np / NP, n = np - dxo * NP;
        int dxi = kp / KP, k = kp - dxi * KP;
        int kw = 16 * t + dxi - dxo;
        float v = 0.f;
        if (kw >= 0 && kw <= 8) {
            if (mode == 0) {          // forward: K = input channel, N = output channel
                if (k < A && n < B) v = src[(((long long)kh * 9 + kw) * A + k) * B + n];
            } else {                  // data gradient: K = output channel, N = input channel, flipped taps
                if (k < B && n < A) v = src[(((long long)(8 - kh) * 9 + (8 - kw)) * A + n) * B + k];
            }
        }
        if (j.hi) split_store(j.hi, j.lo, i, v);
        else dst[i] = v;
        break;
    }
    case PJ_COPY:             // plain copy (a 3-float gradient slot out of its 4-channel staging)
        dst[i] = src[i];
        break;
    default:
        break;
    }
}

__global__ void __launch_bounds__(256) prep_table_kernel(const __grid_constant__ PrepTable t, int first) {
    FS_PDL_ENTER();
    const PrepJob& j = t.j[first + blockIdx.y];
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= j.total) return;
    run_job(j, i);
}

}  // namespace

long long prep_job_total(const PrepJob& j) {
    switch (j.kind) {
    case PJ_PAD: return (long long)j.a * j.d * j.e;
    case PJ_UNPAD: return (long long)j.a * j.b * j.c;
    case PJ_TRANSPOSE: case PJ_FLIP_TRANSPOSE: return (long long)j.a * j.b * j.c;
    case PJ_UPCONV_COLLAPSE: case PJ_S2_DGRAD_COLLAPSE: case PJ_S2_FWD_COLLAPSE: return 16LL * j.a * j.b;
    case PJ_UPCONV_COLLAPSE_GRAD: case PJ_S2_FWD_COLLAPSE_GRAD: return 9LL * j.a * j.b;
    case PJ_PAIR: return 16LL * j.a * j.b;
    case PJ_UNPAIR: return 4LL * j.a * j.b;
    case PJ_PACK_TAPS: return (long long)j.a * j.b * j.c;
    case PJ_PACK_W3X3: return 9LL * j.a * j.b;
    case PJ_IN15: return 8;
    case PJ_COPY: return j.a;
    case PJ_X16: return ((j.e >> 1) & 1) ? 9LL * 16 * j.c * 8 * j.d : 18LL * 16 * j.c * 16 * j.d;
    default: return 0;
    }
}

int PrepPlan::add(int phase, const PrepJob& job) {
    FS_CHECK(n < PREP_MAX_JOBS, "prep plan: more than %d jobs", PREP_MAX_JOBS);
    FS_CHECK(phase >= 0 && phase < PREP_MAX_PHASES, "prep plan: bad phase %d", phase);
    FS_CHECK(n == 0 || phase >= phase_of[n - 1], "prep plan: jobs must be added in phase order");
    t.j[n] = job;
    t.j[n].total = prep_job_total(job);
    FS_CHECK(t.j[n].total > 0, "prep plan: empty job of kind %d", job.kind);
    phase_of[n] = phase;
    ++n;
    return 0;
}

// one launch per phase; launches are stream-ordered, so a phase sees everything the previous ones wrote
int PrepPlan::run(cudaStream_t st) const {
    int i = 0;
    while (i < n) {
        int k = i;
        long long mx = 0;
        while (k < n && phase_of[k] == phase_of[i]) { mx = t.j[k].total > mx ? t.j[k].total : mx; ++k; }
        dim3 grid((unsigned)cdiv(mx, 256), (unsigned)(k - i));
        launch_k(prep_table_kernel, grid, dim3(256), 0, st, t, i);
        FS_LAUNCH_CHECK();
        i = k;
    }
    return 0;
}

}  // namespace fs
