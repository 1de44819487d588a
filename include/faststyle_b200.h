/* faststyle_b200 — C-ABI of the B200-native fast-style-transfer hot path.
 *
 * The reference (ghwatson/faststyle) has no FFI: its hot path is reached through
 * Python functions that build TensorFlow-1 graphs and `tf.Session.run` calls.
 * Each entry point below names the reference call site(s) whose device work it
 * replaces.  Conventions:
 *   - every function returns 0 on success, non-zero on failure;
 *     fs_last_error() returns a thread-local message for the last failure;
 *   - all tensor arguments are DEVICE pointers to fp32, NHWC activations and
 *     HWIO weights, exactly the reference's layouts (im_transf_net.py:113,
 *     libs/vgg16.py:46); the caller owns every buffer;
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it;
 *   - no hidden allocation: composites run out of a caller-provided workspace
 *     whose size is queried from the engine plan.
 * The library has no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef FASTSTYLE_B200_H
#define FASTSTYLE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FS_VERSION 100

const char* fs_last_error(void);
int fs_version(void);
/* number of kernels this library has launched so far in this process */
long long fs_launch_count(void);

/* ------------------------------------------------------------------ layouts
 * Transform-net parameters live in ONE flat fp32 buffer whose order is the
 * byte-sorted checkpoint key order of the reference's models/\*.ckpt
 * (48 tensors, 424102 floats; reference train.py:198-199 `img_t_net` scope).
 * conv index: 0-2 initconv_k, 3+2r / 4+2r resblock_r W1 / W2, 13-15 upsample_k. */
#define FS_T_NCONV 16
long long fs_transform_param_count(void);
/* which: 0 = W, 1 = INscale, 2 = INshift */
int fs_transform_param_slot(int conv_index, int which, long long* offset, long long* count);

/* VGG16 conv1_1..conv4_3 (reference libs/vgg16.py:45-173, load_weights :257-266).
 * flat   = conv1_1_W, conv1_1_b, conv1_2_W, ... conv4_3_b  (HWIO, RGB first layer)
 * packed = engine-internal layout produced once by fs_vgg_pack. */
#define FS_V_NCONV 10
long long fs_vgg_flat_floats(void);
long long fs_vgg_packed_floats(void);
int fs_vgg_pack(const float* flat, float* packed, void* stream);

/* ------------------------------------------------------------------ engine plan */
typedef struct fs_engine fs_engine;
/* FS_ENG_DECONV: the transform net's 'deconv' upsampling variant (conv2d_transpose layers, reference
 * im_transf_net.py:57-63,158-190; forward and backward).  Without it the 'resize' variant (resize-convolution)
 * is planned. */
enum { FS_ENG_TRANSFORM = 1, FS_ENG_TRANSFORM_BWD = 2, FS_ENG_VGG = 4, FS_ENG_VGG_BWD = 8, FS_ENG_DECONV = 16 };

/* Plan for batch N of HxW RGB images.  content_mask / style_mask: bit l set =
 * VGG conv index l (0 = conv1_1 ... 9 = conv4_3) carries a content / style tap. */
int fs_engine_create(int N, int H, int W, int flags, unsigned content_mask, unsigned style_mask,
                     fs_engine** out);
int fs_engine_destroy(fs_engine* e);
size_t fs_engine_workspace_bytes(const fs_engine* e);
int fs_engine_bind(fs_engine* e, void* workspace, size_t bytes);
/* 1 (default): 3x3 convolutions with 64-multiple channels run on the tcgen05 tensor path
 * (split-bf16 x3); 0: every convolution on the exact-fp32 FFMA path.  Also settable
 * through the environment variable FS_TENSOR_PATH=0 read at fs_engine_create. */
int fs_engine_set_tensor_path(fs_engine* e, int enabled);
/* Process-wide: 1 (default) runs the tcgen05 convolution as CTA pairs (clusters of two CTAs, tcgen05.mma
 * cta_group::2 with M = 256, the weight tile split between the two CTAs' shared memories); 0: one CTA per tile
 * (cta_group::1).  Same results up to fp32 summation order.  Also FS_TC_PAIR=0 in the environment. */
int fs_set_tc_pair(int enabled);
/* Process-wide: epilogue warps of the tcgen05 conv launches without fused statistics / fp32 reference tensor - 16
 * (default; registers re-divided with setmaxnreg) or 8 (the previous shape; an A/B switch).  Also FS_TC_EPI_WARPS=8. */
int fs_set_tc_epilogue_warps(int warps);
/* keep = 1: the composites also write the fp32 copies of activations / gradients whose only consumers are tensor-path
 * kernels reading split-bf16 planes (they are skipped by default); needed before fs_engine_transform_activation on
 * such layers.  Also FS_KEEP_ACTS=1 in the environment. */
int fs_engine_keep_activations(fs_engine* e, int keep);
/* frozen = 1: the transform parameters passed to fs_transform_forward do not change between calls (inference):
 * their per-call preparation (padding, collapsing, pairing, bf16 packing: ~25 short launches) runs once and is
 * skipped afterwards.  Call again (any value) after changing the parameter buffer. */
int fs_engine_set_frozen_weights(fs_engine* e, int frozen);
/* Live per-kernel timing: while enabled, the GEMM-class launches of every composite are
 * bracketed by CUDA events on the launching stream.  fs_engine_profile_read synchronises, sums
 * elapsed ms / algorithmic FLOPs / launch counts per category and resets.  Categories:
 * 0 tc VGG conv fwd, 1 tc VGG dgrad, 2 tc residual conv fwd, 3 tc residual dgrad,
 * 4 FFMA implicit-GEMM / direct conv + dgrad, 5 weight gradients, 6 Gram forward, 7 Gram backward,
 * 8 InstanceNorm statistics, 9 InstanceNorm apply, 10 InstanceNorm backward, 11 pooling / padding /
 * split / other element-wise, 12 losses, 13 per-step weight preparation, 14 tc stride-2 / resize conv fwd
 * (collapsed 2x2 forms), 15 their data gradients, 16 tc 9x9 conv fwd (x16 space-to-depth forms), 17 its data
 * gradient, 18 tc VGG conv1_1 fwd (im2col tile built in shared memory). */
#define FS_PROF_NCAT 19
int fs_engine_profile(fs_engine* e, int enabled);
int fs_engine_profile_read(fs_engine* e, int ncat, float* ms, double* flops, int* launches);
/* algorithmic bytes per category of the launches recorded so far (tensor-path convolutions: every operand plane
 * read once + every output plane written once; 0 for categories without a byte model); call before _read */
int fs_engine_profile_bytes(fs_engine* e, int ncat, double* bytes);
/* per-launch records (category, ms, algorithmic FLOPs) in issue order; call before _read */
int fs_engine_profile_records(fs_engine* e, int max_rec, int* cat, float* ms, double* flops, int* count);
/* output dims of the transform net (== VGG input dims when both are planned) */
int fs_engine_output_dims(const fs_engine* e, int* OH, int* OW);
/* device pointer + dims of a saved VGG activation (post-ReLU conv output) */
int fs_engine_vgg_activation(const fs_engine* e, int layer, const float** ptr, int* H, int* W, int* C);
/* saved transform-net activation: stage 0 = raw conv output, 1 = post IN+activation */
int fs_engine_transform_activation(const fs_engine* e, int conv_index, int stage, const float** ptr,
                                   int* H, int* W, int* C);

typedef struct {
    int n_content; int content_layer[FS_V_NCONV]; float content_w[FS_V_NCONV];
    int n_style;   int style_layer[FS_V_NCONV];   float style_w[FS_V_NCONV];
    float beta;
} fs_loss_config;

/* ------------------------------------------------------------------ composites */
/* create_net forward, upsample_method='resize' (reference im_transf_net.py:14-75;
 * replaces sess.run(Y) at stylize_image.py:75).  x3 [N,H,W,3] 0..255 -> y3 [N,OH,OW,3]. */
int fs_transform_forward(fs_engine* e, const float* params, const float* x3, float* y3, void* stream);

/* VGG16 forward to conv index `upto`; activations stay in the workspace
 * (reference libs/vgg16.py:36-173; replaces sess.run(content_layers) train.py:250). */
int fs_vgg_forward(fs_engine* e, const float* packed, const float* img3, int upto, void* stream);

/* Gram matrices G = F^T F / (h*w*c) of the listed layers of img3
 * (reference utils.py:66-83; replaces sess.run(get_grams) train.py:150).
 * out_grams[i] -> device [N, C_i, C_i]. */
int fs_vgg_grams(fs_engine* e, const float* packed, const float* img3, int n_layers, const int* layers,
                 float* const* out_grams, void* stream);

/* Store VGG(content_img3) activations of cfg's content layers as the engine's
 * content targets (reference slow_style.py:133-137). */
int fs_vgg_set_content_targets(fs_engine* e, const float* packed, const float* img3,
                               const fs_loss_config* cfg, void* stream);

/* Perceptual loss of an image (+ gradient w.r.t. its pixels when grad3 != NULL):
 * content + style + beta*TV (reference losses.py:12-97, slow_style.py:140-154).
 * losses4 (device float[4]) = {content, style, beta*tv, total}. */
int fs_perceptual_loss(fs_engine* e, const float* packed, const float* img3, const fs_loss_config* cfg,
                       const float* const* target_grams, float* losses4, float* grad3, void* stream);

/* One training step's forward + backward (reference train.py:250-255,274-275 up to
 * the gradient): content targets from the raw batch, transform fwd, VGG fwd, losses,
 * VGG data-gradient, transform backward.  grads: flat [424102] (sum over the local
 * batch - the reference's losses sum over the batch, losses.py:32-37,63-64).
 * y3 may be NULL. */
int fs_train_fwd_bwd(fs_engine* e, const float* params, const float* packed, const float* x3,
                     const fs_loss_config* cfg, const float* const* target_grams, float* grads,
                     float* losses4, float* y3, void* stream);

/* tf.train.AdamOptimizer update on a flat buffer (reference train.py:203,
 * slow_style.py:153): epsilon outside the bias correction.  step_counter: device
 * int holding the number of updates applied so far (incremented by the call). */
int fs_adam_step(float* params, const float* grads, float* m, float* v, long long n, float lr,
                 float beta1, float beta2, float eps, int* step_counter, void* stream);

/* Data-parallel optimiser step in ONE kernel over NVLink peer memory (reference: train.py:198-204 on a batch sharded
 * over the GPUs of a node): all-reduce(SUM) of every rank's flat gradient buffer + the TF-Adam update of this rank's
 * replica.  peer_bufs: HOST array of `world` device pointers - rank r's exchange buffer as addressable from this
 * device (CUDA IPC mapping / peer access; entry [rank] is the local buffer).  Layout of every exchange buffer:
 * this step's n gradients at float offset grad_off_floats (double-buffer by step parity), n_extra scalars to be
 * summed into extra_out (loss terms; may be 0 / NULL) at grad_off_floats + extra_off_floats, an unsigned[world] flag
 * array (zero-initialised, then owned by the kernel) at byte offset flag_off_bytes.  tag: the step number
 * (1, 2, 3, ... strictly increasing; the same on all ranks); 0 = take *step_counter + 1 on the device (for CUDA-graph
 * replay, where host arguments are frozen).  Sums are formed in rank order on every rank, so the
 * replicas stay bit-identical.  err_flag (device int, may be NULL) is set to 1 when a peer's flag did not arrive
 * within ~10 s.  Every rank must make the call once per step. */
/* Let kernels of the CURRENT device dereference memory of `peer_device` (cudaDeviceEnablePeerAccess; a no-op when it
 * is already enabled or peer_device is the current device).  Needed once per peer before fs_dp_allreduce_adam. */
int fs_enable_peer_access(int peer_device);
/* Exchange buffers for fs_dp_allreduce_adam.  create: a zero-filled device allocation on the current device + its
 * 64-byte CUDA IPC handle (send it to the peer processes of the node); open: map a peer's buffer into the CURRENT
 * device's address space (peer access over NVLink is enabled by the open); close / free release them. */
int fs_peer_buffer_create(size_t bytes, void** dev_ptr, unsigned char* handle64);
int fs_peer_buffer_open(const unsigned char* handle64, void** dev_ptr);
int fs_peer_buffer_close(void* dev_ptr);
int fs_peer_buffer_free(void* dev_ptr);
int fs_dp_allreduce_adam(const void* const* peer_bufs, int rank, int world, long long grad_off_floats, long long n,
                         long long extra_off_floats, int n_extra, long long flag_off_bytes, unsigned tag, float* params,
                         float* m, float* v, float lr, float beta1, float beta2, float eps, int* step_counter,
                         float* extra_out, int* err_flag, void* stream);

/* ------------------------------------------------------------------ single ops
 * (exported for op-level parity tests and the Python op mirror) */
/* tf.nn.conv2d NHWC/HWIO; padding_same: 1 = TF 'SAME', 0 = 'VALID'; C and OC must be
 * multiples of 4 (reference im_transf_net.py:115, libs/vgg16.py:48-52). */
int fs_conv2d_forward(const float* x, const float* w, const float* bias, float* y, int N, int H, int W,
                      int C, int KH, int KW, int OC, int stride, int padding_same, int relu, void* stream);
/* data gradient of the above; wt_scratch: KH*KW*C*OC floats */
int fs_conv2d_dgrad(const float* dy, const float* w, float* dx, float* wt_scratch, int N, int H, int W,
                    int C, int KH, int KW, int OC, int stride, int padding_same, void* stream);
/* weight gradient; scratch: fs_conv2d_wgrad_scratch_floats() floats */
long long fs_conv2d_wgrad_scratch_floats(int C, int KH, int KW, int OC);
int fs_conv2d_wgrad(const float* x, const float* dy, float* dw, float* scratch, long long scratch_floats,
                    int N, int H, int W, int C, int KH, int KW, int OC, int stride, int padding_same,
                    void* stream);
/* upconv2d = NN resize x4 + 3x3 stride-2 SAME conv, fused (reference im_transf_net.py:122-155);
 * wc_scratch: 16*C*OC floats.  x [N,H,W,C] -> y [N,2H,2W,OC]. */
int fs_upconv2d_forward(const float* x, const float* w, float* y, float* wc_scratch, int N, int H, int W,
                        int C, int OC, void* stream);
/* inst_norm + activation (0 none, 1 relu, 2 scaled tanh) (reference im_transf_net.py:193-247);
 * scratch_doubles >= N*64*2*C doubles, stats: 2*N*C floats (mean, rstd) */
int fs_instnorm_forward(const float* x, const float* scale, const float* shift, float* y, float* stats,
                        double* scratch, int N, int H, int W, int C, float eps, int act, void* stream);
int fs_maxpool2x2(const float* x, float* y, int N, int H, int W, int C, void* stream);
/* G = F^T F/(h*w*c), F [N,H,W,C] -> G [N,C,C] (reference utils.py:76-81) */
long long fs_gram_scratch_floats(int N, int C);
int fs_gram_forward(const float* f, float* g, float* scratch, long long scratch_floats, int N, int H, int W,
                    int C, void* stream);

/* Loss single ops (reference losses.py): acc = device double[4] scratch, out = device float[4];
 * the result is out[0] for sqdiff (content), out[1] for style, out[2] for tv.
 *   fs_loss_sqdiff: scale * sum (a-b)^2             (content_loss term, losses.py:32-37)
 *   fs_loss_style : scale * sum (G[n]-T)^2, T [CC] broadcast over N (style_loss term, :61-64)
 *   fs_loss_tv    : sum of squared forward differences of Y3 [N,H,W,3]  (tv_loss, :86-95) */
int fs_loss_sqdiff(const float* a, const float* b, long long n, double scale, double* acc, float* out, void* stream);
int fs_loss_style(const float* G, const float* T, int N, int CC, double scale, double* acc, float* out, void* stream);
int fs_loss_tv(const float* Y3, int N, int H, int W, double* acc, float* out, void* stream);

/* ------------------------------------------------------------------ streaming frames
 * Device-side uint8 <-> fp32 conversion around fs_transform_forward for the frame-by-frame path
 * (reference stylize_webcam.py:82-90: the uint8 frame is fed as-is - the float cast happens inside
 * feed_dict -, the result is cast with numpy astype(uint8), i.e. truncated, and channels 0/2 are
 * swapped by cv2.cvtColor(COLOR_BGR2RGB)).  in/out are DEVICE pointers; n = number of bytes (4-byte
 * aligned input, 16-byte aligned output); npix = number of 3-channel pixels.  swap_rb bit 0: swap channels 0/2;
 * bit 1: round half to even before the saturating cast (what cv2.imwrite does to a float image, reference
 * utils.py:51-52 / stylize_image.py:79-80) instead of truncating. */
int fs_frame_u8_to_f32(const unsigned char* in, float* out, long long n, void* stream);
int fs_frame_f32_to_u8(const float* in, unsigned char* out, long long npix, int swap_rb, void* stream);

/* ------------------------------------------------------------------ input pipeline
 * tf.image.resize_images(img, resize_shape, method=2) of TensorFlow 1.0 (reference datapipe.py:14-26,
 * `preprocessing`): legacy bicubic (A = -0.75, 1024-entry coefficient table, no half-pixel centres, clamped
 * borders) of one decoded uint8 HWC RGB image into a float32 [OH, OW, 3] slot of the batch.  Device pointers. */
int fs_resize_bicubic_tf1_u8(const unsigned char* in, float* out, int H, int W, int OH, int OW, void* stream);

/* ------------------------------------------------------------------ tensor-core path
 * 3x3 stride-1 convolution on the tcgen05 tensor pipe with split-bf16 operands
 * (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM); C and OC multiples of 64.
 * Same semantics as fs_conv2d_forward / fs_conv2d_dgrad with KH=KW=3, stride 1.
 * scratch: fs_conv3x3_tc_scratch_bytes() bytes, 1024-byte aligned. */
size_t fs_conv3x3_tc_scratch_bytes(int N, int H, int W, int C, int OC);
int fs_conv3x3_tc_forward(const float* x, const float* w, const float* bias, float* y, void* scratch,
                          size_t scratch_bytes, int N, int H, int W, int C, int OC, int padding_same,
                          int relu, void* stream);
int fs_conv3x3_tc_dgrad(const float* dy, const float* w, float* dx, void* scratch, size_t scratch_bytes,
                        int N, int H, int W, int C, int OC, int padding_same, void* stream);

/* weight gradient of a 3x3 stride-1 64->64 convolution on the tensor path: x [N,H,W,64],
 * dy [N,OH,OW,64] -> dw [3,3,64,64]; same semantics as fs_conv2d_wgrad */
size_t fs_wgrad3x3_tc_scratch_bytes(int N, int H, int W);
int fs_wgrad3x3_tc(const float* x, const float* dy, float* dw, void* scratch, size_t scratch_bytes, int N, int H,
                   int W, int padding_same, void* stream);

/* Gram matrix on the tensor path (the kernel the engine uses for the style taps): same semantics as
 * fs_gram_forward, C a multiple of 64.  FS_GRAM_V2=0 in the environment selects the first-generation
 * 128x64-tile kernel instead of the full-width one. */
size_t fs_gram_tc_scratch_bytes(int N, int H, int W, int C);
int fs_gram_tc_forward(const float* f, float* g, void* scratch, size_t scratch_bytes, int N, int H, int W, int C,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif
