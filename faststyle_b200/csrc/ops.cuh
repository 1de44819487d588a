// Launchers for the HBM-bound (non-GEMM) kernels of the hot path.
#pragma once
#include "common.cuh"

namespace fs {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH255 = 2 };

// K1: tf.pad REFLECT + channel pad 3->4           (reference im_transf_net.py:78-88)
int reflect_pad_c4(const float* x, float* out, int N, int H, int W, int pad, cudaStream_t st, void* x16hi = nullptr,
                   void* x16lo = nullptr, int x8 = 0);
// K9: x - VGG mean, channel pad 3->4             (reference libs/vgg16.py:41-42)
int vgg_preprocess_c4(const float* x, float* out, long long npix, cudaStream_t st);

// uint8 RGB/BGR frames <-> fp32 (streaming path, reference stylize_webcam.py:82-90)
int frame_u8_to_f32(const unsigned char* in, float* out, long long n, cudaStream_t st);
int frame_f32_to_u8(const float* in, unsigned char* out, long long npix, int swap_rb, cudaStream_t st);

// input pipeline: tf.image.resize_images(method=2) of TF 1.0 on a uint8 HWC image (reference datapipe.py:25)
int resize_bicubic_tf1_u8(const unsigned char* in, float* out, int H, int W, int OH, int OW, cudaStream_t st);

// K3: InstanceNorm                               (reference im_transf_net.py:218-247)
int in_chunks(int N, int HW);
int instnorm_stats(const float* x, float* mean, float* rstd, int N, int HW, int C, float eps,
                   double* partial, cudaStream_t st);
// statistics from per-(n,c) (sum, sum of squares) accumulated by a conv epilogue; the sums are zeroed again
int instnorm_stats_from_sums(double* sums, float* mean, float* rstd, int N, int HW, int C, float eps, int reps, cudaStream_t st);
int instnorm_apply(const float* x, const float* mean, const float* rstd, const float* scale,
                   const float* shift, const float* skip, float* out, int N, int H, int W, int C,
                   int act, int out3, cudaStream_t st, void* split_hi = nullptr, void* split_lo = nullptr);
// backward: dz = dY*act'(z); dgamma[c] += sum dz*xhat; dbeta[c] += sum dz;
// dx = scale*rstd*(dz - mean(dz) - xhat*mean(dz*xhat))
int instnorm_bwd(const float* dY, const float* x, const float* mean, const float* rstd,
                 const float* scale, const float* shift, float* dx, float* dgamma, float* dbeta,
                 int N, int HW, int C, int act, double* partial, float* m12 /*[N][C][2]*/,
                 cudaStream_t st, void* split_hi = nullptr, void* split_lo = nullptr);

// fused forms (one / two launches per site instead of two / three).  `sums` / `bs`: [reps][N*C][2] fp64 accumulators
// (filled by a conv epilogue / by the reduction launch); `zero_buf` (zero_n doubles): the other ping-pong set, zeroed
// for the next site.  instnorm_apply_from_sums also publishes mean / rstd for the backward pass.
int instnorm_apply_from_sums(const float* x, const double* sums, int reps, float eps, float* mean, float* rstd,
                             double* zero_buf, long long zero_n, const float* scale, const float* shift,
                             const float* skip, float* out, int N, int H, int W, int C, int act, int out3,
                             cudaStream_t st, void* split_hi = nullptr, void* split_lo = nullptr);
int instnorm_bwd_sums(const float* dY, const float* x, const float* mean, const float* rstd, const float* scale,
                      const float* shift, float* dx, float* dgamma, float* dbeta, int N, int HW, int C, int act,
                      double* bs, int reps, double* zero_buf, long long zero_n, cudaStream_t st,
                      void* split_hi = nullptr, void* split_lo = nullptr);

// data-parallel step: all-reduce(SUM) of the ranks' flat gradient buffers over peer memory fused with TF-Adam
// (dp_adam.cu).  peer_bufs[r]: rank r's exchange buffer as addressable from THIS device; gradients of this step at
// float offset grad_off, n_extra scalars (summed into extra_out) at grad_off + extra_off, the flag array
// (unsigned[world]) at byte offset flag_off_bytes; tag: step number, strictly increasing.
int dp_allreduce_adam(const void* const* peer_bufs, int rank, int world, long long grad_off, long long n,
                      long long extra_off, int n_extra, long long flag_off_bytes, unsigned tag, float* params, float* m,
                      float* v, float lr, float b1, float b2, float eps, int* step_counter, float* extra_out, int* err,
                      cudaStream_t st);

// K8: 2x2 s2 SAME max-pool                        (reference libs/vgg16.py:67-71)
int maxpool2x2_fwd(const float* x, float* out, int N, int H, int W, int C, cudaStream_t st,
                   void* split_hi = nullptr, void* split_lo = nullptr);
// out = [mask(act>0)] * ( route(gpool) + cw2*(act - ctarget) )
int pool_bwd_combine(const float* act, const float* gpool, const float* ctarget, float cw2,
                     int apply_mask, float* out, int N, int H, int W, int C, cudaStream_t st);

// K11/K12 losses. loss_acc: device double[4] = {content, style, tv, unused}
int sqdiff_sum(const float* a, const float* b, long long n, double scale, double* acc, cudaStream_t st);
// S[n] = coef*(G[n]-T); acc += lscale*sum (G-T)^2
int style_loss_grad(const float* G, const float* T, float* S, int N, int CC, float coef,
                    double lscale, double* acc, cudaStream_t st);
// TV on Y [N,H,W,3]; acc += beta*tv; dY4 [N,H,W,4] += beta*dTV/dY  (dY4 may be null)
int tv_loss_grad(const float* Y, float* dY4, int N, int H, int W, float beta, double* acc, cudaStream_t st);
int finalize_losses(const double* acc, float* out4, cudaStream_t st);

// K13: TF-Adam on a flat buffer                   (reference train.py:203)
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1,
              float b2, float eps, int* step_counter, cudaStream_t st);

// weight layout helpers
int pad_taps(const float* src, float* dst, int T, int Ci, int Co, int Cip, int Cop, cudaStream_t st);
int unpad_taps(const float* src, float* dst, int T, int Ci, int Co, int Cip, int Cop, cudaStream_t st);
int transpose_taps(const float* src, float* dst, int T, int Ci, int Co, cudaStream_t st);
// resize-conv collapse W[3,3,Ci,Co] -> W'[2,2,Ci,4Co] and its adjoint for gradients
int upconv_collapse(const float* W, float* Wc, int Ci, int Co, cudaStream_t st);
int upconv_collapse_grad(const float* dWc, float* dW, int Ci, int Co, cudaStream_t st);
// 3x3 stride-2 (pad 0/1) conv: weights of the 4-phase 2x2 data-gradient form, [2][2][Co][4*Ci]
int s2_dgrad_collapse(const float* W, float* Wd, int Ci, int Co, cudaStream_t st);

// 3x3 stride-2 (pad 0/1) conv: weights of the 2x2 stride-1 form over the space-to-depth input, [2][2][4*Ci][Co]
int s2_fwd_collapse(const float* W, float* Wf, int Ci, int Co, cudaStream_t st);

// pixel pairing of a 2x2-tap GEMM, W [2][2][K0][N0] -> Wp [2][2][2*K0][2*N0] (see pair_taps_kernel)
int pair_taps(const float* W, float* Wp, int K0, int N0, int kmode, int gather, cudaStream_t st);

// adjoints (weight gradients back through pair_taps / s2_fwd_collapse)
int unpair_taps(const float* dWp, float* dW, int K0, int N0, int kmode, int n_s2d, cudaStream_t st);
int s2_fwd_collapse_grad(const float* dWf, float* dW, int Ci, int Co, cudaStream_t st);

// fp32 [rows, W, C] -> split-bf16 planes [rows, W + 16, C], row content at column offset 4 (margins untouched)
int split_pad_x16(const float* x, void* hi, void* lo, long long rows, int W, int C, cudaStream_t st, int x8 = 0);
int fill_zero(void* p, size_t bytes, cudaStream_t st);

}  // namespace fs
