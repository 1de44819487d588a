// Engine: plans (shapes + workspace layout) and whole-path composites.
#pragma once
#include "common.cuh"
#include "ops.cuh"
#include "tc.cuh"
#include "prep.cuh"
#include <vector>

namespace fs {

// ---------------------------------------------------------------- transform net
constexpr int T_NCONV = 16;         // 3 init + 10 residual + 3 "upsample" convs
constexpr long long T_NPARAMS = 424102;

struct TConv {
    int k, stride, same, cin, cout;     // reference geometry (im_transf_net.py:37-70)
    int upconv;                         // 1: resize-conv (NN x4 + 3x3 s2 SAME), run collapsed
    int act;                            // activation after InstanceNorm
    long long offW, offG, offB;         // offsets into the flat parameter buffer (floats)
    // runtime geometry for one plan
    int inH, inW, outH, outW;           // conv input / output spatial dims
    int cin_s, cout_s;                  // stored channel counts (3 -> 4)
    int pad_t, pad_l;
};

struct TBuf { float *raw, *act, *mean, *rstd; };

// ---------------------------------------------------------------- VGG16 (conv1_1..conv4_3)
constexpr int V_NCONV = 10;
struct VConv {
    int cin, cout, cin_s, pool_after;
    long long offW, offB, offWT;        // fp32: padded HWIO weights, bias, transposed weights
    long long offTcF, offTcD;           // tensor path (layers >= 1): packed split-bf16 weights for the
                                        // forward / data-gradient conv; hi plane, lo plane follows
    int H, W;
};
long long vgg_flat_floats();        // unpacked: W,b per layer in order (HWIO)
long long vgg_packed_floats();      // packed: padded W, b, transposed W per layer

struct LossConfig {
    int n_content; int content_layer[V_NCONV]; float content_w[V_NCONV];
    int n_style;   int style_layer[V_NCONV];   float style_w[V_NCONV];
    float beta;
};

enum ProfCat { PC_TC_VGG_FWD = 0, PC_TC_VGG_DGRAD, PC_TC_RES_FWD, PC_TC_RES_DGRAD, PC_FFMA_CONV, PC_WGRAD,
               PC_GRAM_FWD, PC_GRAM_BWD, PC_IN_STATS, PC_IN_APPLY, PC_IN_BWD, PC_POINTWISE, PC_LOSS, PC_PREP,
               PC_TC_S2_FWD, PC_TC_S2_DGRAD, PC_TC9_FWD, PC_TC9_DGRAD, PC_TC_C11_FWD, PC_COUNT };
enum EngineFlags { ENG_TRANSFORM = 1, ENG_TRANSFORM_BWD = 2, ENG_VGG = 4, ENG_VGG_BWD = 8, ENG_DECONV = 16 };

struct Arena {
    char* base = nullptr; size_t off = 0; size_t cap = 0;
    template <typename T> T* take(long long n) {
        size_t bytes = ((size_t)n * sizeof(T) + 255) & ~(size_t)255;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

struct Engine {
    int N, H, W, flags;
    unsigned content_mask = 0, style_mask = 0;   // VGG conv indices with loss taps
    // transform plan
    TConv tc[T_NCONV];
    int Hp, Wp, OH, OW;                  // padded-input dims, output dims
    // vgg plan
    VConv vc[V_NCONV];
    int VH, VW;
    // workspace
    size_t ws_bytes = 0;
    bool bound = false;
    // --- device buffers (assigned by layout())
    float* xpad4 = nullptr;              // [N,Hp,Wp,4]
    TBuf tb[T_NCONV];
    float* weff[T_NCONV];                // effective forward weights (null = use flat params)
    float* wefft[T_NCONV];               // transposed weights for the data gradient
    float* y3 = nullptr;                 // [N,OH,OW,3] when the caller does not supply an output
    double* in_partial = nullptr;
    // [STATS_REPLICAS][N][64][2] fp64 accumulators in two ping-pong sets.  Forward: layer l's conv epilogue adds
    // (sum, sum sq) of its raw output to in_sums2[l & 1]; the fused apply of layer l reads them and zeroes set (l+1)&1
    // (last read one layer earlier).  Backward: the same scheme on bw_sums2 with (sum dz, sum dz*xhat).
    double* in_sums2[2] = {nullptr, nullptr};
    double* bw_sums2[2] = {nullptr, nullptr};
    long long sums_n = 0;                // doubles per set
    int fuse_in = 1;                     // FS_IN_FUSE=0: separate finalize launches (stats -> finalize -> apply chains)
    int in_epi = 1;                      // FS_IN_EPILOGUE=0: separate statistics pass for every layer
    float* in15 = nullptr;               // 4-channel staging of the last layer's IN scale/shift
    float* gb_tmp = nullptr;
    float* wtmp15 = nullptr;             // staging for the deconv variant of the last layer's weights
    float* m12 = nullptr;
    float* tgrad[3];                     // rotating gradient buffers (transform bwd)
    float* wg_partial = nullptr; long long wg_partial_cap = 0;
    float* wg_tmp = nullptr;             // padded / collapsed weight-gradient staging
    float* wg_tmp2 = nullptr;            // un-paired weight-gradient staging (tensor-path 2x2 forms)
    // vgg
    float* v_in4 = nullptr;              // [N,VH,VW,4]
    float* vact[V_NCONV]; float* vpool[V_NCONV];
    float* vgrad[4]; long long vgrad_floats = 0;
    float* gram[V_NCONV]; float* gramS[V_NCONV];
    float* ctarget[V_NCONV];
    float* dY4 = nullptr;                // [N,VH,VW,4] gradient w.r.t. the VGG input image
    double* loss_acc = nullptr;          // double[4]
    // tensor-core path (tcgen05): split-bf16 companions of the tensors feeding 3x3 convs
    int use_tc = 1;
    // inference with fixed weights: prepare (pad / collapse / pair / pack) once, skip the ~25 preparation launches
    // of every later forward.  Cleared by any call that may have changed the parameters or the path.
    int frozen_weights = 0; bool weights_prepared = false;
    SplitPtr vsplit[V_NCONV];            // input planes of VGG conv l (l >= 1)
    SplitPtr vgsplit[4];                 // planes of vgrad[i]
    SplitPtr vtsplit[V_NCONV];           // planes of a style-tapped activation when no later conv holds them
    SplitPtr gsS[V_NCONV];               // packed per-sample Gram-space gradients S (tensor-path Gram backward)
    SplitPtr tsplit[T_NCONV];            // input planes of the tensor-path transform convs (2, 3..12, 13)
    SplitPtr tgsplit[3];                 // planes of tgrad[i] (dRaw of the tensor-path transform convs)
    SplitPtr rg[T_NCONV];                // residual convs (3..12): planes of dRaw kept per layer, so that their ten
                                         // weight gradients run as ONE batched launch after the backward sweep
    // 9x9 layers on the tensor path (initconv_0 forward, upsample_2 forward + data gradient): the stride-1 SAME 9x9
    // convolution over [H, W, c] is run as a 9x2-tap convolution over [H, W/16 (+1), 16c] - sixteen horizontally
    // adjacent pixels form one GEMM pixel - with Toeplitz-expanded weights (PJ_X16).  The input planes carry a zero
    // margin of 4 pixels on the left and 12 on the right (p9[.], written by split_pad_x16).
    SplitPtr p9[3];                      // inputs: initconv_0 forward, upsample_2 forward, upsample_2 data gradient
    SplitPtr tw9[3];                     // Toeplitz-expanded weights, packed split-bf16 [18][16 KP / 64][16 NP][64]
    int tc9_on = 1;                      // FS_TC9=0: 9x9 layers on the direct FFMA kernels
    // The two forms with a 4-channel input side (initconv_0 forward, upsample_2 data gradient) use 8-pixel output groups
    // over 16-pixel input windows instead (planes [H, W/8, 16 px x 4]: every pixel stored into the two windows that
    // overlap it): ONE horizontal tap, K = 64, N = 8 x 16 = 128 - half the MMA work of the x16 form.  FS_TC9_X8=0: x16.
    int tc9_x8 = 1;
    bool tc9() const;
    int tc9_conv(int which, const float* src_f32, float* out, bool stats, cudaStream_t st, bool planes_ready = false);
    // weight gradients of the two 9x9 layers on tcgen05 in their x8 forms (wgrad9_tc.cu): the windowed planes p9[0] /
    // p9[2] against the plain planes of layer 0's raw-output gradient (g0split) / of layer 15's input (a14split).
    // FS_WGRAD9_TC=0: the direct CUDA-core kernels.
    int wgrad9_tc = 1;
    SplitPtr g0split = {nullptr, nullptr}, a14split = {nullptr, nullptr};
    bool wg9() const { return wgrad9_tc && tc9_x8 && tc9(); }
    int keep_acts = 0;                   // 1: write every fp32 activation / gradient even where only split planes are read (debug taps)
    int fuse_pool = 1;                   // FS_FUSE_POOL=0: separate max-pool kernel after the VGG conv
    int fold_pool = 1;                   // FS_FOLD_POOL=0: separate pool_bwd_combine pass before the Gram backward
    // ReLU / arg-max code bytes (Conv3x3TcArgs::ref_code) instead of fp32 copies of the VGG activations in the
    // training pass: the backward epilogues need only "value > 0" and "first maximum of its pooling window" of a
    // layer that is no content target.  FS_RELU_CODES=0: fp32 references as before.
    int relu_codes = 1;
    int conv11_tc = 1;                   // FS_CONV11_TC=0: VGG conv1_1 on the exact-fp32 direct kernel
    unsigned char* vcode[V_NCONV];
    int batch_wgrad = 1;                 // FS_BATCH_WGRAD=0: one weight-gradient launch per residual conv
    SplitPtr tw_f[T_NCONV], tw_d[T_NCONV];   // packed weights (forward / data gradient)
    float* w2f = nullptr;                // initconv_1/2 weights in the 2x2 space-to-depth form (fp32 staging)
    float* wpair = nullptr;              // pixel-paired 2x2 weights (fp32 staging, one layer at a time)
    // table-driven preparation (prep.cu): every transform has its own staging buffer so that all transforms of a
    // dependency phase run in ONE launch
    float* w2f_b = nullptr;              // initconv_2's space-to-depth form (w2f holds initconv_1's)
    float* wpair_x[4] = {nullptr, nullptr, nullptr, nullptr};   // paired forms: l=1 fwd, l=14 fwd, l=1 bwd, l=14 bwd
    float* wgs[T_NCONV];                 // weight-gradient staging of the layers whose gradient needs a layout adjoint
    float* wgu[2] = {nullptr, nullptr};  // un-paired weight gradients of l=14 / l=1
    int fast_prep = 1;                   // FS_FAST_PREP=0: one launch per transform (the round-1 path)
    bool use_fast_prep() const;
    int prep_transform_weights_table(const float* params, bool need_bwd, cudaStream_t st);
    int finish_weight_grads_table(float* grads, cudaStream_t st);
    // stride-2 layers on the tensor path in their collapsed 2x2 stride-1 forms: initconv_2 (3x3 s2 32->64) and
    // upsample_0 (resize-conv 64->32) have 64/128 channels on both sides as they are; initconv_1 (16->32) and
    // upsample_1 (32->16) get there by pairing horizontally adjacent pixels (pair_taps).  Forward + data gradient.
    bool tc2(int l) const;
    bool tc2_layout(int l) const;        // geometry part of tc2 (independent of use_tc)
    int tc2_args(int l, bool bwd, int ri, float* out, Conv3x3TcArgs& ta) const;

    // live per-kernel timing (bench.py roofline)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev; std::vector<int> prof_cat; std::vector<double> prof_flops, prof_bytes;
    double next_bytes = 0.0;             // algorithmic bytes of the next bracketed launch (PROFB)
    int pbegin(int cat, double flops, cudaStream_t st);
    int pend(cudaStream_t st);
    int prof_read(int ncat, float* ms, double* flops, int* launches);
    int prof_read_bytes(int ncat, double* bytes);
    int prof_records(int max_rec, int* cat, float* ms, double* flops, int* count);

    int plan();                          // fill geometry; returns 0 / error
    void layout(Arena& a);               // assign (or just size) workspace
    int bind(void* ws, size_t bytes);

    // composites (all asynchronous on st)
    int prep_transform_weights(const float* params, bool need_bwd, cudaStream_t st);
    int transform_forward(const float* params, const float* x3, float* y3_out, cudaStream_t st);
    int transform_backward(const float* params, const float* dY4_in, float* grads, cudaStream_t st);
    // code_mask: bit l set = layer l stores code bytes (vcode[l]) instead of its fp32 activation
    int vgg_forward(const float* packed, const float* img3, int upto, float* const* act_override,
                    cudaStream_t st, unsigned code_mask = 0);
    int vgg_content_targets(const float* packed, const float* img3, const LossConfig& lc, cudaStream_t st);
    int vgg_loss_backward(const float* packed, const float* img3, const LossConfig& lc,
                          const float* const* target_grams, bool need_grad, cudaStream_t st);
    int train_fwd_bwd(const float* params, const float* packed, const float* x3, const LossConfig& lc,
                      const float* const* target_grams, float* grads, float* losses4, float* y3_out,
                      cudaStream_t st);
};

int vgg_pack(const float* flat, float* packed, cudaStream_t st);
void transform_param_table(TConv* tc);        // fills reference geometry + offsets

}  // namespace fs
