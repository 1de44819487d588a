// Tensor-core (tcgen05 / TMEM / TMA) path: 3x3 stride-1 implicit-GEMM convolution with
// split-bf16 operands.  sm_100a only.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace fs {

// A "split" fp32 tensor: x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi)  (16 mantissa bits).
struct SplitPtr { __nv_bfloat16* hi; __nv_bfloat16* lo; };

// fp32 [n] -> hi/lo bf16 planes (n % 4 == 0)
int split_bf16(const float* x, SplitPtr out, long long n, cudaStream_t st);

// Pack HWIO fp32 weights [3,3,Ci,Co] for the tensor path.
//   mode 0 (forward):        B[tap=(kh,kw)][cb][n=co][k=ci%64]          = W[kh,kw,cb*64+k,n]
//   mode 1 (data gradient):  B[tap=(2-kh,2-kw)][cb][n=ci][k=co%64]      = W[kh,kw,n,cb*64+k]
// Output: hi/lo planes of 9*(K/64)*N*64 bf16 each (K = reduction channels, N = output channels).
int pack_w3x3_tc(const float* w, SplitPtr out, int Ci, int Co, int mode, cudaStream_t st);
// Generic tap-major GEMM weights w [T][K][Nn] fp32 -> B[tap'][cb][n][k] = w[tap][cb*64+k][n] split planes
// (tap = T-1-tap' when flip: the data-gradient forms correlate with the flipped kernel)
int pack_taps_tc(const float* w, SplitPtr out, int T, int K, int Nn, int flip, cudaStream_t st);
// same for up to 16 equally shaped weight tensors in ONE launch
int pack_w3x3_tc_batch(const float* const* w, const SplitPtr* out, int count, int Ci, int Co, int mode, cudaStream_t st);

constexpr int STATS_REPLICAS = 16;

struct Conv3x3TcArgs {
    SplitPtr x;                // input  [N,H,W,C]  split bf16 planes (C % 64 == 0)
    SplitPtr w;                // packed weights (pack_w3x3_tc), reduction channels = C, outputs = OC
    int N, H, W, C;
    int OH, OW, OC;            // output dims (OC % 64 == 0)
    int pad;                   // zero padding on top/left (SAME: 1, VALID: 0, VALID data gradient: 2)
    int taps;                  // kernel extent KH = KW: 0/3 = 3x3 (default), 2 = 2x2 (the collapsed stride-2 / resize
                               // convolutions and their data gradients; weights packed by pack_taps_tc)
    int taps_w, pad_x;         // non-square forms: horizontal taps / left padding (0 = same as taps / pad).  taps = 9 with
                               // taps_w = 2 is the 9x9 stride-1 SAME conv of the transform net over the x16
                               // space-to-depth view of a zero-margined input (see Engine::tc9)
    int in_s2d;                // 1: x is the space-to-depth view [N,H,W,C] of a plain [N,2H,2W,C/4] tensor (C == 128)
    int out_d2s;               // 1: the [OH,OW,OC] result is stored depth-to-space into out_f32 [N,2OH,2OW,OC/4]
                               // 2: same for pixel-paired operands, OC = (e,p,q,16) -> out_f32 [N,2OH,4OW,16]
                               //    (OC == 128; fp32 output only, no bias/addend/ref/relu/split)
    int one_by_one;            // 1: 1x1 'convolution' (pure GEMM over channels); pad must be 0
    int per_sample_w;          // with one_by_one: weights are [N][C/64][OC][64] (one matrix per sample)
    // epilogue: v = acc + bias[c] + addend[pix,c]; relu; mask by ref[pix,c] > 0
    const float* bias; const float* addend; const float* ref;
    // with ref: v += route(pool_grad) + cw2 * (ref - ctarget) before the mask.  pool_grad [N,ceil(OH/2),ceil(OW/2),OC]
    // is the gradient w.r.t. the 2x2 max-pool of ref (routed to the first maximum of each window); ctarget
    // [N,OH,OW,OC] the content target of this layer (content-loss gradient 2w/(hwc) (f - t), cw2 = 2w/(hwc))
    const float* pool_grad; const float* ctarget; float cw2;
    // ref_code: the same reference as one byte per element instead of fp32 (written by a forward launch through
    // out_code): bit 0 = value > 0 (the ReLU mask), bit 1 = this element is the FIRST maximum of its 2x2 pooling window
    // (scan order) - the max-pool gradient goes to it.  Excludes ref / ctarget.  A quarter of the reference traffic and
    // no window shuffles in the backward epilogue; the forward launch stores 1 byte instead of 4 per element.
    const unsigned char* ref_code;
    int relu;
    int add_crop, addH, addW;  // addend is [N,addH,addW,OC]; output pixel (y,x) reads (y-crop, x-crop)
    double* stats; int stats_c;    // optional: accumulate per-(sample, real channel) sum / sum of squares of the raw
                               // output into stats[STATS_REPLICAS][N][stats_c][2] (real channel = output channel %
                               // stats_c: the depth-to-space / paired forms carry several pixels' channels side by
                               // side; replica = CTA index % STATS_REPLICAS, summed by instnorm_stats_from_sums)
    float* out_f32;            // [N,OH,OW,OC] fp32 (may be null)
    SplitPtr out_split;        // split planes of the same tensor (may be null)
    SplitPtr pool_split;       // split planes of the 2x2 stride-2 SAME max-pool of the result [N,ceil(OH/2),ceil(OW/2),OC] (may be null)
    unsigned char* out_code;   // [N,OH,OW,OC] ReLU / arg-max codes of the result (see ref_code; bit 1 only with pool_split)
};
int launch_conv3x3_tc(const Conv3x3TcArgs& a, cudaStream_t st);
// CTA-pair (tcgen05 cta_group::2) variant of the kernel on / off (default: on; FS_TC_PAIR=0 in the environment)
void set_tc_pair(int on);
// epilogue warps of the launches without fused statistics / fp32 reference: 16 (default) or 8 (FS_TC_EPI_WARPS=8)
void set_tc_epi_warps(int warps);
bool conv3x3_tc_supported(int C, int OC, int W, int OW);

// Tiled tensor map over an NHWC tensor [N,H,W,C] of 1- / 2- / 4-byte elements with box {boxC, boxW, boxH, 1}; the box's
// inner extent (boxC * elem bytes) must equal swizzle_bytes (64 or 128).  Used for TMA stores of epilogue tiles.
int tc_make_map_nhwc(CUtensorMap* tm, const void* base, int elem_bytes, int N, int H, int W, int C, int boxC, int boxW,
                     int boxH, int swizzle_bytes);

// VGG conv1_1 (3(4) -> 64, 3x3 SAME, bias + ReLU) with the im2col tile built in shared memory (conv1_1_tc.cu):
// in [N,H,W,4] fp32, w [9,4,64] fp32 -> any of fp32 out / split planes / ReLU code bytes
int launch_conv1_1_tc(const float* in, const float* w, const float* bias, float* out, void* split_hi, void* split_lo,
                      unsigned char* code, int N, int H, int W, cudaStream_t st);

// tcgen05 weight gradient of a 3x3 stride-1 64->64 convolution (wgrad_tc.cu):
// x [N,H,W,64], dy [N,OH,OW,64] split planes -> out [3,3,64,64] fp32 (HWIO)
long long wgrad3x3_tc_partial_floats();
int launch_wgrad3x3_tc(SplitPtr x, SplitPtr dy, float* out, float* partial, long long partial_cap, int N, int H,
                       int W, int OH, int OW, int pad, cudaStream_t st);


// the same for up to WGRAD_MULTI_MAX convolutions (equal batch N) in one launch + one reduction
constexpr int WGRAD_MULTI_MAX = 10;
struct WgradMultiItem { SplitPtr x, dy; float* out; int H, W, OH, OW, pad; };
int launch_wgrad3x3_tc_multi(const WgradMultiItem* items, int count, float* partial, long long partial_cap, int N,
                             cudaStream_t st);

// tcgen05 weight gradient of the 9x9 layers in their x8 forms (wgrad9_tc.cu): xw = windowed planes [N,H,G,64] of the
// 4-channel side, dg = plain planes [N,H,8G,16] of the 16-channel side; out [9,9,A,16] (mode 0: the 4-channel side is the
// conv input) or [9,9,16,A] (mode 1: it is the output gradient), A = real channels of that side (3)
long long wgrad9_x8_partial_floats();
int launch_wgrad9_x8_tc(SplitPtr xw, SplitPtr dg, float* out, float* partial, long long partial_cap, int N, int H, int G,
                        int A, int mode, cudaStream_t st);

// tcgen05 weight gradient of the 2x2-tap forms (wgrad_tc.cu): out [2,2,Cx,Cy], (Cx,Cy) = (64,128) | (128,64)
long long wgrad2x2_tc_partial_floats();
int launch_wgrad2x2_tc(SplitPtr x, int x_s2d, SplitPtr dy, int dy_s2d, float* out, float* partial, long long partial_cap,
                       int N, int GH, int GW, cudaStream_t st);

// B[n][cb][j][k] = S[n][cb*64+k][j] as split planes: packs per-sample [C,C] matrices for a 1x1 tensor-path GEMM
int pack_gemm_b_tc(const float* S, SplitPtr out, int N, int C, cudaStream_t st);

// tcgen05 Gram matrix G[n] = scale * F[n]^T F[n]; F [N,HW,C] split planes, G [N,C,C] fp32 (C % 64 == 0)
long long gram_tc_partial_floats(int N, int HW, int C);
int launch_gram_tc(SplitPtr f, float* G, float* partial, long long partial_cap, int N, int HW, int C, float scale,
                   cudaStream_t st);

}  // namespace fs
