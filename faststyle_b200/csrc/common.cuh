// Shared declarations for libfaststyle_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace fs {

// ---------------------------------------------------------------- errors
// Every C-ABI entry point returns 0 on success or a negative code and records
// a message retrievable through fs_last_error() (thread-local).
void set_error(const char* fmt, ...);
const char* get_error();

#define FS_CHECK(cond, ...)                                   \
    do {                                                      \
        if (!(cond)) { fs::set_error(__VA_ARGS__); return -1; } \
    } while (0)

#define FS_CUDA(expr)                                                          \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            fs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                 \
            return -2;                                                         \
        }                                                                      \
    } while (0)

// every kernel launch of this library goes through FS_LAUNCH_CHECK: it also counts launches
// (fs_launch_count) so that callers can report how many of OUR kernels ran in a timed region.
extern long long g_launches;
#define FS_LAUNCH_CHECK()                \
    do {                                 \
        ++fs::g_launches;                \
        FS_CUDA(cudaGetLastError());     \
    } while (0)

#define FS_TRY(expr)                   \
    do {                               \
        int _r = (expr);               \
        if (_r != 0) return _r;        \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Raises a kernel's dynamic shared-memory limit once per (device, kernel); thread-safe, so engines on several
// devices / host threads of one process can launch concurrently (the attribute is per device).
int ensure_dyn_smem(const void* fn, int bytes);
#define FS_DYN_SMEM(kern, bytes) FS_TRY(fs::ensure_dyn_smem(reinterpret_cast<const void*>(kern), (int)(bytes)))
// SM count of the current device (cached per device)
int num_sms();

// ---------------------------------------------------------------- programmatic dependent launch
// The step is a chain of ~240 short kernels (mean 30 us).  Every kernel is launched with the
// programmatic-stream-serialization attribute and starts with FS_PDL_ENTER(): it releases its own
// dependents at once and then waits for the WHOLE preceding grid to complete and flush
// (griddepcontrol.wait), so memory ordering is exactly that of plain stream order while the launch
// latency, CTA scheduling and the predecessor's tail overlap.  Nothing may touch global memory before
// FS_PDL_ENTER() (the tensor-core kernels split it: FS_PDL_TRIGGER() first, then barrier init / TMEM
// allocation / descriptor prefetch -- none of which reads global memory -- then FS_PDL_WAIT() by every
// thread).  FS_PDL=0 in the environment turns the attribute off (the device side is then a no-op).
extern int g_pdl;   // -1 unknown, 0 off, 1 on
int pdl_enabled();

#ifdef __CUDACC__
#define FS_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#define FS_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define FS_PDL_ENTER()    \
    do {                  \
        FS_PDL_TRIGGER(); \
        FS_PDL_WAIT();    \
    } while (0)

template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);   // the error is picked up by FS_LAUNCH_CHECK
}
// same, as clusters of two CTAs (CTA pairs for tcgen05 cta_group::2): grid.x must be even
template <typename... KArgs, typename... Args>
static inline void launch_k_cluster2(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    (void)cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);   // the error is picked up by FS_LAUNCH_CHECK
}
#endif

// TF 'SAME' padding rule: out = ceil(n/s),
// total = max((out-1)*s + k - n, 0), before = total/2.
static inline void tf_same(int n, int k, int s, int* out, int* before) {
    int o = (n + s - 1) / s;
    int total = (o - 1) * s + k - n;
    if (total < 0) total = 0;
    *out = o;
    *before = total / 2;
}

// ---------------------------------------------------------------- implicit GEMM
// One descriptor drives the FFMA implicit-GEMM kernels (forward conv, data
// gradient, fused resize-conv, Gram backward).  GEMM view per sample n:
//   out[m, j] = sum_k A[m, k] * Wm[k, j],   m = oy*OW+ox,  k = (kh*KW+kw)*C + c
// A is gathered on the fly from the NHWC input.
struct IGemmArgs {
    const float* in;  const float* w;  float* out;
    int H, W, C;                 // gathered tensor: logical spatial dims, channels per tap (C%4==0)
    int in_mode;                 // 0: plain NHWC [H,W,C]; 1: space-to-depth view of [2H,2W,C/4]
    long long in_bs;             // input batch stride (elements)
    int KH, KW, stride, pad_t, pad_l;
    int gather;                  // 0: iy = oy*stride - pad_t + kh      (forward)
                                 // 1: iy = (oy + pad_t - kh)/stride if divisible (data gradient)
    int OH, OW, OC;              // GEMM M = OH*OW, N = OC (OC%4==0)
    int out_mode;                // 0: plain [OH,OW,OC]; 1: depth-to-space into [2OH,2OW,OC/4]
    long long out_bs;            // output batch stride (elements)
    long long w_bs;              // weight batch stride (0 = shared across the batch)
    // epilogue: v += bias[j]; v += addend[...]; if relu v=max(v,0); if ref && ref<=0 v=0
    const float* bias; const float* addend; const float* ref;
    int relu;
    int add_crop, addH, addW;    // addend is [N,addH,addW,OC]; pixel (y,x) reads (y-crop, x-crop)
    long long add_bs;
    int N;
};
int launch_igemm(const IGemmArgs& a, cudaStream_t st);

// Reduction-over-pixels GEMM (weight gradient, Gram forward):
//   out[g][k, j] = scale * sum_pix A[pix, k] * B[pix, j]
struct WGradArgs {
    const float* in;  const float* dy;  float* out;  float* partial;  // partial: workspace
    long long partial_cap;       // capacity of partial in floats
    int H, W, C; int in_mode; long long in_bs;      // gathered A side (as IGemmArgs)
    int KH, KW, stride, pad_t, pad_l;
    int OH, OW, OC;              // dy logical dims [OH,OW,OC]
    int dy_mode;                 // 0 plain; 1 space-to-depth view of [2OH,2OW,OC/4]
    long long dy_bs;
    int N;
    int per_sample;              // 1: separate output per sample (Gram), 0: reduce over batch
    float scale;
};
int launch_wgrad(const WGradArgs& a, cudaStream_t st);
long long wgrad_partial_floats(int K, int OC, int groups);

// direct 9x9 stride-1 SAME weight gradient for (CI,CO) in {(16,4),(4,16)}  (direct9x9.cu)
long long wgrad9x9_partial_floats(int N, int H, int W);
int launch_conv9x9(const float* in, const float* w, float* out, int N, int H, int W, int CI, int CO, cudaStream_t st);
int flip_transpose_taps(const float* W, float* Wf, int T, int Ci, int Co, cudaStream_t st);
int launch_conv3x3_c4_fwd(const float* in, const float* w, const float* bias, float* out, void* split_hi, void* split_lo,
                          int N, int H, int W, cudaStream_t st, unsigned char* code = nullptr);
int launch_dgrad3x3_c4(const float* P, const float* wf, float* dx, int N, int H, int W, cudaStream_t st);
int launch_wgrad9x9(const float* in, const float* dy, float* out, float* partial, long long partial_cap, int N,
                    int H, int W, int CI, int CO, cudaStream_t st);

}  // namespace fs
