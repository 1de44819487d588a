// Data-parallel optimiser step as ONE kernel over NVLink peer memory: gradient all-reduce(SUM) + TF-Adam.
//
// Every rank owns a flat exchange buffer [gradients | loss scalars] that all ranks of the node can address (CUDA IPC
// mappings of the peers' buffers; NVLink / NVSwitch peer loads).  After the backward pass each rank
//   1. signals "my buffer of step t is complete" by storing t into slot [rank] of every peer's flag array
//      (st.release.sys after a system fence),
//   2. waits until its own flag array shows t for every rank (ld.acquire.sys on local memory),
//   3. reads element i of ALL world buffers (peer loads), adds them in rank order 0..world-1 - the same order on every
//      rank, so every replica computes bit-identical sums and the parameters stay identical without a broadcast - and
//      applies the TF-Adam update to its own replica in the same pass.
// The 1.7 MB all-gather-style read replaces NCCL's latency-bound all-reduce launch plus the separate Adam launch (the
// gradient never goes back to memory in reduced form).  The exchange buffers are double-buffered by step parity: a
// rank can run at most one step ahead of the slowest peer (step t+1's flags need every rank's step-t kernel to have
// finished), so buffer t&1 is never overwritten while a peer still reads it.
// Reference semantics: train.py:198-204 (minimize(loss) on the summed losses) on a batch sharded over ranks.
#include "ops.cuh"

namespace fs {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {          // peer data: bypass L1 (written by another GPU)
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

constexpr int DP_MAX_WORLD = 16;
struct DpPeers { const float* grad[DP_MAX_WORLD]; unsigned* flags[DP_MAX_WORLD]; };

__global__ void __launch_bounds__(256)
dp_allreduce_adam_kernel(const DpPeers peers, int rank, int world, unsigned tag, long long n, long long extra_off, int n_extra,
                         float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, float lr, float b1, float b2,
                         float eps, const int* __restrict__ step, float* __restrict__ extra_out, int* __restrict__ err) {
    FS_PDL_ENTER();                      // this rank's backward pass (previous kernels of the stream) is complete
    __shared__ float lr_t_s;
    const int t = threadIdx.x;
    if (tag == 0u) tag = (unsigned)(*step + 1);      // graph replay: the step number lives on the device
    if (blockIdx.x == 0 && t < world) {
        __threadfence_system();
        st_release_sys(peers.flags[t] + rank, tag);
    }
    if (t == 0) {
        const int ts = *step + 1;
        lr_t_s = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)ts)) / (1.0 - pow((double)b1, (double)ts)));
    }
    if (t < world) {                     // every CTA waits on the LOCAL flag array (cheap polling, no peer traffic)
        const unsigned* f = peers.flags[rank] + t;
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - tag) < 0) {
            if (clock64() - t0 > 20000000000LL) {          // ~10 s: a peer died; report instead of hanging the GPU
                if (err) atomicExch(err, 1);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    const float lr_t = lr_t_s;
    const long long n4 = n >> 2;
    const long long i = (long long)blockIdx.x * blockDim.x + t;
    if (i < n4) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 q = ld_peer4(peers.grad[r] + 4 * i);
            g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
        }
        float4 mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i], pi = reinterpret_cast<float4*>(p)[i];
        const float gg[4] = {g.x, g.y, g.z, g.w};
        float mm[4] = {mi.x, mi.y, mi.z, mi.w}, vv[4] = {vi.x, vi.y, vi.z, vi.w}, pp[4] = {pi.x, pi.y, pi.z, pi.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            mm[j] = b1 * mm[j] + (1.f - b1) * gg[j];
            vv[j] = b2 * vv[j] + (1.f - b2) * gg[j] * gg[j];
            pp[j] = pp[j] - lr_t * mm[j] / (sqrtf(vv[j]) + eps);
        }
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
        reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
    }
    // tail elements (n % 4) and the extra scalars (loss terms, summed for logging) by the last CTA
    if (blockIdx.x == gridDim.x - 1) {
        for (long long k = 4 * n4 + t; k < n; k += 256) {
            float g = 0.f;
            for (int r = 0; r < world; ++r) g += ld_peer1(peers.grad[r] + k);
            const float mk = b1 * m[k] + (1.f - b1) * g, vk = b2 * v[k] + (1.f - b2) * g * g;
            m[k] = mk; v[k] = vk;
            p[k] = p[k] - lr_t * mk / (sqrtf(vk) + eps);
        }
        if (extra_out && t < n_extra) {
            float s = 0.f;
            for (int r = 0; r < world; ++r) s += ld_peer1(peers.grad[r] + extra_off + t);
            extra_out[t] = s;
        }
    }
}

__global__ void dp_incr_kernel(int* c) { FS_PDL_ENTER(); *c += 1; }

}  // namespace

int dp_allreduce_adam(const void* const* peer_bufs, int rank, int world, long long grad_off, long long n,
                      long long extra_off, int n_extra, long long flag_off_bytes, unsigned tag, float* params, float* m,
                      float* v, float lr, float b1, float b2, float eps, int* step_counter, float* extra_out, int* err,
                      cudaStream_t st) {
    FS_CHECK(world >= 1 && world <= DP_MAX_WORLD && rank >= 0 && rank < world, "dp_allreduce_adam: bad rank / world (%d / %d)", rank, world);
    FS_CHECK(peer_bufs && params && m && v && step_counter && n > 0, "dp_allreduce_adam: NULL argument");
    FS_CHECK(grad_off % 4 == 0 && flag_off_bytes % 4 == 0 && n_extra >= 0 && n_extra <= 256, "dp_allreduce_adam: bad layout");
    DpPeers P;
    memset(&P, 0, sizeof(P));
    for (int r = 0; r < world; ++r) {
        FS_CHECK(peer_bufs[r] != nullptr, "dp_allreduce_adam: NULL buffer of rank %d", r);
        P.grad[r] = reinterpret_cast<const float*>(peer_bufs[r]) + grad_off;
        P.flags[r] = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(const_cast<void*>(peer_bufs[r])) + flag_off_bytes);
    }
    const long long n4 = n >> 2;
    const int grid = n4 > 0 ? (int)((n4 + 255) / 256) : 1;      // the last CTA also takes the n % 4 tail and the extras
    launch_k(dp_allreduce_adam_kernel, dim3(grid), dim3(256), 0, st, P, rank, world, tag, n, extra_off, n_extra, params, m, v,
             lr, b1, b2, eps, (const int*)step_counter, extra_out, err);
    FS_LAUNCH_CHECK();
    launch_k(dp_incr_kernel, dim3(1), dim3(1), 0, st, step_counter);
    FS_LAUNCH_CHECK();
    return 0;
}

}  // namespace fs
