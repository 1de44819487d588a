// extern "C" surface of libfaststyle_b200 (see include/faststyle_b200.h).
#include "../../include/faststyle_b200.h"
#include "engine.cuh"
#include "tc.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
#include <mutex>
#include <new>
#include <set>
#include <utility>

namespace fs {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }
long long g_launches = 0;

int ensure_dyn_smem(const void* fn, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<int, const void*>> done;
    int dev = 0;
    FS_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    if (done.count(std::make_pair(dev, fn))) return 0;
    FS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done.insert(std::make_pair(dev, fn));
    return 0;
}

int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
int g_pdl = -1;
int pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("FS_PDL");
        g_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_pdl;
}
}  // namespace fs

using namespace fs;

struct fs_engine { Engine e; };

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

static int to_lc(const fs_loss_config* c, LossConfig& lc) {
    FS_CHECK(c != nullptr, "loss config is NULL");
    FS_CHECK(c->n_content >= 0 && c->n_content <= V_NCONV && c->n_style >= 0 && c->n_style <= V_NCONV,
             "loss config: bad layer counts");
    lc.n_content = c->n_content; lc.n_style = c->n_style; lc.beta = c->beta;
    for (int i = 0; i < V_NCONV; ++i) {
        lc.content_layer[i] = c->content_layer[i]; lc.content_w[i] = c->content_w[i];
        lc.style_layer[i] = c->style_layer[i]; lc.style_w[i] = c->style_w[i];
    }
    return 0;
}

extern "C" {

const char* fs_last_error(void) { return get_error(); }
int fs_version(void) { return FS_VERSION; }
long long fs_launch_count(void) { return g_launches; }

long long fs_transform_param_count(void) { return T_NPARAMS; }

int fs_transform_param_slot(int conv_index, int which, long long* offset, long long* count) {
    FS_CHECK(conv_index >= 0 && conv_index < T_NCONV && which >= 0 && which <= 2, "bad slot query");
    TConv tc[T_NCONV];
    transform_param_table(tc);
    const TConv& c = tc[conv_index];
    if (which == 0) { *offset = c.offW; *count = (long long)c.k * c.k * c.cin * c.cout; }
    else { *offset = which == 1 ? c.offG : c.offB; *count = c.cout; }
    return 0;
}

long long fs_vgg_flat_floats(void) { return vgg_flat_floats(); }
long long fs_vgg_packed_floats(void) { return vgg_packed_floats(); }
int fs_vgg_pack(const float* flat, float* packed, void* stream) { return vgg_pack(flat, packed, S(stream)); }

int fs_engine_create(int N, int H, int W, int flags, unsigned content_mask, unsigned style_mask,
                     fs_engine** out) {
    FS_CHECK(out != nullptr, "fs_engine_create: out is NULL");
    FS_CHECK(flags != 0 && (flags & ~31) == 0, "fs_engine_create: bad flags 0x%x", flags);
    FS_CHECK(!(flags & ENG_DECONV) || (flags & ENG_TRANSFORM), "DECONV selects a transform-net variant: needs TRANSFORM");
    FS_CHECK(!(flags & ENG_TRANSFORM_BWD) || (flags & ENG_TRANSFORM), "TRANSFORM_BWD needs TRANSFORM");
    FS_CHECK(!(flags & ENG_VGG_BWD) || (flags & ENG_VGG), "VGG_BWD needs VGG");
    FS_CHECK((content_mask >> V_NCONV) == 0 && (style_mask >> V_NCONV) == 0, "loss masks address layers > conv4_3");
    fs_engine* h = new (std::nothrow) fs_engine();
    FS_CHECK(h != nullptr, "out of host memory");
    h->e.N = N; h->e.H = H; h->e.W = W; h->e.flags = flags;
    h->e.content_mask = content_mask; h->e.style_mask = style_mask;
    const char* env = getenv("FS_TENSOR_PATH");
    h->e.use_tc = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_IN_EPILOGUE");
    h->e.in_epi = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_TC9");
    h->e.tc9_on = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_KEEP_ACTS");
    h->e.keep_acts = (env && env[0] == '1') ? 1 : 0;
    env = getenv("FS_WGRAD9_TC");
    h->e.wgrad9_tc = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_TC9_X8");
    h->e.tc9_x8 = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_CONV11_TC");
    h->e.conv11_tc = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_RELU_CODES");
    h->e.relu_codes = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_IN_FUSE");
    h->e.fuse_in = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_FUSE_POOL");
    h->e.fuse_pool = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_FOLD_POOL");
    h->e.fold_pool = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_BATCH_WGRAD");
    h->e.batch_wgrad = (env && env[0] == '0') ? 0 : 1;
    env = getenv("FS_FAST_PREP");
    h->e.fast_prep = (env && env[0] == '0') ? 0 : 1;
    int r = h->e.plan();
    if (r != 0) { delete h; return r; }
    Arena a;
    h->e.layout(a);
    h->e.ws_bytes = a.off;
    *out = h;
    return 0;
}

int fs_engine_set_tensor_path(fs_engine* e, int enabled) {
    FS_CHECK(e, "NULL engine");
    e->e.use_tc = enabled ? 1 : 0;
    e->e.weights_prepared = false;
    return 0;
}
int fs_engine_keep_activations(fs_engine* e, int keep) {
    FS_CHECK(e, "NULL engine");
    e->e.keep_acts = keep ? 1 : 0;
    return 0;
}
int fs_set_tc_pair(int enabled) { fs::set_tc_pair(enabled); return 0; }
int fs_set_tc_epilogue_warps(int warps) {
    FS_CHECK(warps == 8 || warps == 16, "fs_set_tc_epilogue_warps: 8 or 16");
    fs::set_tc_epi_warps(warps);
    return 0;
}
int fs_engine_profile(fs_engine* e, int enabled) {
    FS_CHECK(e, "NULL engine");
    e->e.prof_on = enabled != 0;
    return 0;
}
int fs_engine_profile_read(fs_engine* e, int ncat, float* ms, double* flops, int* launches) {
    FS_CHECK(e && ms && flops && launches && ncat >= 1 && ncat <= 64, "fs_engine_profile_read: bad argument");
    return e->e.prof_read(ncat, ms, flops, launches);
}
int fs_engine_profile_bytes(fs_engine* e, int ncat, double* bytes) {
    FS_CHECK(e && bytes && ncat >= 1 && ncat <= 64, "fs_engine_profile_bytes: bad argument");
    return e->e.prof_read_bytes(ncat, bytes);
}
int fs_engine_profile_records(fs_engine* e, int max_rec, int* cat, float* ms, double* flops, int* count) {
    FS_CHECK(e && cat && ms && flops && count && max_rec >= 0, "fs_engine_profile_records: bad argument");
    return e->e.prof_records(max_rec, cat, ms, flops, count);
}
int fs_engine_destroy(fs_engine* e) { delete e; return 0; }
size_t fs_engine_workspace_bytes(const fs_engine* e) { return e ? e->e.ws_bytes : 0; }
int fs_engine_bind(fs_engine* e, void* workspace, size_t bytes) {
    FS_CHECK(e && workspace, "fs_engine_bind: NULL argument");
    return e->e.bind(workspace, bytes);
}
int fs_engine_output_dims(const fs_engine* e, int* OH, int* OW) {
    FS_CHECK(e, "NULL engine");
    *OH = e->e.VH; *OW = e->e.VW;
    return 0;
}
int fs_engine_vgg_activation(const fs_engine* e, int layer, const float** ptr, int* H, int* W, int* C) {
    FS_CHECK(e && e->e.bound && (e->e.flags & ENG_VGG), "engine has no bound VGG plan");
    FS_CHECK(layer >= 0 && layer < V_NCONV, "bad VGG layer %d", layer);
    *ptr = e->e.vact[layer]; *H = e->e.vc[layer].H; *W = e->e.vc[layer].W; *C = e->e.vc[layer].cout;
    return 0;
}
int fs_engine_transform_activation(const fs_engine* e, int conv_index, int stage, const float** ptr,
                                   int* H, int* W, int* C) {
    FS_CHECK(e && e->e.bound && (e->e.flags & ENG_TRANSFORM), "engine has no bound transform plan");
    FS_CHECK(conv_index >= 0 && conv_index < T_NCONV && (stage == 0 || stage == 1), "bad activation query");
    const TConv& c = e->e.tc[conv_index];
    *ptr = stage == 0 ? e->e.tb[conv_index].raw : e->e.tb[conv_index].act;
    FS_CHECK(*ptr != nullptr, "activation not stored (the last layer's output is y3)");
    *H = c.outH; *W = c.outW; *C = c.cout_s;
    return 0;
}

int fs_engine_set_frozen_weights(fs_engine* e, int frozen) {
    FS_CHECK(e, "NULL engine");
    e->e.frozen_weights = frozen ? 1 : 0;
    e->e.weights_prepared = false;
    return 0;
}

int fs_transform_forward(fs_engine* e, const float* params, const float* x3, float* y3, void* stream) {
    FS_CHECK(e && params && x3 && y3, "fs_transform_forward: NULL argument");
    if (!(e->e.frozen_weights && e->e.weights_prepared)) {
        FS_TRY(e->e.prep_transform_weights(params, false, S(stream)));
        e->e.weights_prepared = true;
    }
    return e->e.transform_forward(params, x3, y3, S(stream));
}

int fs_vgg_forward(fs_engine* e, const float* packed, const float* img3, int upto, void* stream) {
    FS_CHECK(e && packed && img3, "fs_vgg_forward: NULL argument");
    return e->e.vgg_forward(packed, img3, upto, nullptr, S(stream));
}

int fs_vgg_grams(fs_engine* e, const float* packed, const float* img3, int n_layers, const int* layers,
                 float* const* out_grams, void* stream) {
    FS_CHECK(e && packed && img3 && layers && out_grams, "fs_vgg_grams: NULL argument");
    FS_CHECK(n_layers >= 1 && n_layers <= V_NCONV, "fs_vgg_grams: bad layer count");
    Engine& E = e->e;
    int top = 0;
    for (int i = 0; i < n_layers; ++i) {
        FS_CHECK(layers[i] >= 0 && layers[i] < V_NCONV, "fs_vgg_grams: bad layer %d", layers[i]);
        top = layers[i] > top ? layers[i] : top;
    }
    FS_TRY(E.vgg_forward(packed, img3, top, nullptr, S(stream)));
    for (int i = 0; i < n_layers; ++i) {
        const VConv& v = E.vc[layers[i]];
        WGradArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.in = E.vact[layers[i]]; wa.dy = wa.in; wa.out = out_grams[i];
        wa.partial = E.wg_partial; wa.partial_cap = E.wg_partial_cap;
        wa.H = v.H; wa.W = v.W; wa.C = v.cout; wa.in_bs = (long long)v.H * v.W * v.cout;
        wa.KH = wa.KW = 1; wa.stride = 1; wa.OH = v.H; wa.OW = v.W; wa.OC = v.cout; wa.dy_bs = wa.in_bs;
        wa.N = E.N; wa.per_sample = 1; wa.scale = (float)(1.0 / ((double)v.H * v.W * v.cout));
        FS_TRY(launch_wgrad(wa, S(stream)));
    }
    return 0;
}

int fs_vgg_set_content_targets(fs_engine* e, const float* packed, const float* img3,
                               const fs_loss_config* cfg, void* stream) {
    FS_CHECK(e && packed && img3, "fs_vgg_set_content_targets: NULL argument");
    LossConfig lc;
    FS_TRY(to_lc(cfg, lc));
    return e->e.vgg_content_targets(packed, img3, lc, S(stream));
}

int fs_perceptual_loss(fs_engine* e, const float* packed, const float* img3, const fs_loss_config* cfg,
                       const float* const* target_grams, float* losses4, float* grad3, void* stream) {
    FS_CHECK(e && packed && img3 && losses4, "fs_perceptual_loss: NULL argument");
    LossConfig lc;
    FS_TRY(to_lc(cfg, lc));
    Engine& E = e->e;
    FS_CHECK(E.bound && (E.flags & ENG_VGG), "engine has no bound VGG plan");
    cudaStream_t st = S(stream);
    FS_TRY(fill_zero(E.loss_acc, 4 * sizeof(double), st));
    FS_TRY(E.vgg_loss_backward(packed, img3, lc, target_grams, grad3 != nullptr, st));
    if (lc.beta != 0.f) FS_TRY(tv_loss_grad(img3, grad3 ? E.dY4 : nullptr, E.N, E.VH, E.VW, lc.beta, E.loss_acc + 2, st));
    if (grad3) FS_TRY(unpad_taps(E.dY4, grad3, E.N * E.VH * E.VW, 1, 3, 1, 4, st));
    return finalize_losses(E.loss_acc, losses4, st);
}

int fs_train_fwd_bwd(fs_engine* e, const float* params, const float* packed, const float* x3,
                     const fs_loss_config* cfg, const float* const* target_grams, float* grads,
                     float* losses4, float* y3, void* stream) {
    FS_CHECK(e && params && packed && x3 && grads && losses4, "fs_train_fwd_bwd: NULL argument");
    LossConfig lc;
    FS_TRY(to_lc(cfg, lc));
    return e->e.train_fwd_bwd(params, packed, x3, lc, target_grams, grads, losses4, y3, S(stream));
}

int fs_adam_step(float* params, const float* grads, float* m, float* v, long long n, float lr,
                 float beta1, float beta2, float eps, int* step_counter, void* stream) {
    FS_CHECK(params && grads && m && v && step_counter && n > 0, "fs_adam_step: bad argument");
    return adam_step(params, grads, m, v, n, lr, beta1, beta2, eps, step_counter, S(stream));
}

int fs_enable_peer_access(int peer_device) {
    int cur = 0;
    FS_CUDA(cudaGetDevice(&cur));
    if (cur == peer_device) return 0;
    int can = 0;
    FS_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
    FS_CHECK(can, "fs_enable_peer_access: device %d cannot access device %d (no NVLink / PCIe peer path)", cur, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return 0; }
    FS_CUDA(e);
    return 0;
}

// Exchange buffers of the data-parallel step: plain cudaMalloc allocations (exact base pointers, not sub-allocations of
// a caching allocator) exported / imported through CUDA IPC under the CONSUMER's device, so the driver maps the peer's
// memory into this device's address space and enables peer access (NVLink) as part of the open.
int fs_peer_buffer_create(size_t bytes, void** dev_ptr, unsigned char* handle64) {
    FS_CHECK(dev_ptr && handle64 && bytes > 0, "fs_peer_buffer_create: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    FS_CUDA(cudaMalloc(&p, bytes));
    FS_CUDA(cudaMemset(p, 0, bytes));
    FS_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); FS_CUDA(e); }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return 0;
}
int fs_peer_buffer_open(const unsigned char* handle64, void** dev_ptr) {
    FS_CHECK(dev_ptr && handle64, "fs_peer_buffer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    FS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return 0;
}
int fs_peer_buffer_close(void* dev_ptr) {
    if (dev_ptr) FS_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}
int fs_peer_buffer_free(void* dev_ptr) {
    if (dev_ptr) FS_CUDA(cudaFree(dev_ptr));
    return 0;
}

int fs_dp_allreduce_adam(const void* const* peer_bufs, int rank, int world, long long grad_off_floats, long long n,
                         long long extra_off_floats, int n_extra, long long flag_off_bytes, unsigned tag, float* params,
                         float* m, float* v, float lr, float beta1, float beta2, float eps, int* step_counter,
                         float* extra_out, int* err_flag, void* stream) {
    return dp_allreduce_adam(peer_bufs, rank, world, grad_off_floats, n, extra_off_floats, n_extra, flag_off_bytes, tag,
                             params, m, v, lr, beta1, beta2, eps, step_counter, extra_out, err_flag, S(stream));
}

// ------------------------------------------------------------------ single ops
static int conv_geom(int H, int W, int KH, int KW, int stride, int same, int* OH, int* OW, int* pt, int* pl) {
    FS_CHECK(stride >= 1 && KH >= 1 && KW >= 1, "conv: bad geometry");
    if (same) { tf_same(H, KH, stride, OH, pt); tf_same(W, KW, stride, OW, pl); }
    else { *OH = (H - KH) / stride + 1; *OW = (W - KW) / stride + 1; *pt = *pl = 0; }
    FS_CHECK(*OH > 0 && *OW > 0, "conv: empty output (%dx%d input, %dx%d kernel)", H, W, KH, KW);
    return 0;
}

int fs_conv2d_forward(const float* x, const float* w, const float* bias, float* y, int N, int H, int W,
                      int C, int KH, int KW, int OC, int stride, int padding_same, int relu, void* stream) {
    FS_CHECK(x && w && y, "fs_conv2d_forward: NULL argument");
    IGemmArgs a;
    memset(&a, 0, sizeof(a));
    FS_TRY(conv_geom(H, W, KH, KW, stride, padding_same, &a.OH, &a.OW, &a.pad_t, &a.pad_l));
    a.in = x; a.w = w; a.out = y; a.N = N; a.H = H; a.W = W; a.C = C; a.in_bs = (long long)H * W * C;
    a.KH = KH; a.KW = KW; a.stride = stride; a.OC = OC; a.out_bs = (long long)a.OH * a.OW * OC;
    a.bias = bias; a.relu = relu;
    return launch_igemm(a, S(stream));
}

int fs_conv2d_dgrad(const float* dy, const float* w, float* dx, float* wt_scratch, int N, int H, int W,
                    int C, int KH, int KW, int OC, int stride, int padding_same, void* stream) {
    FS_CHECK(dy && w && dx && wt_scratch, "fs_conv2d_dgrad: NULL argument");
    IGemmArgs a;
    memset(&a, 0, sizeof(a));
    int OH, OW;
    FS_TRY(conv_geom(H, W, KH, KW, stride, padding_same, &OH, &OW, &a.pad_t, &a.pad_l));
    FS_TRY(transpose_taps(w, wt_scratch, KH * KW, C, OC, S(stream)));
    a.in = dy; a.w = wt_scratch; a.out = dx; a.N = N; a.gather = 1;
    a.H = OH; a.W = OW; a.C = OC; a.in_bs = (long long)OH * OW * OC;
    a.KH = KH; a.KW = KW; a.stride = stride;
    a.OH = H; a.OW = W; a.OC = C; a.out_bs = (long long)H * W * C;
    return launch_igemm(a, S(stream));
}

long long fs_conv2d_wgrad_scratch_floats(int C, int KH, int KW, int OC) {
    return wgrad_partial_floats(KH * KW * C, OC, 1);
}

int fs_conv2d_wgrad(const float* x, const float* dy, float* dw, float* scratch, long long scratch_floats,
                    int N, int H, int W, int C, int KH, int KW, int OC, int stride, int padding_same,
                    void* stream) {
    FS_CHECK(x && dy && dw && scratch, "fs_conv2d_wgrad: NULL argument");
    WGradArgs wa;
    memset(&wa, 0, sizeof(wa));
    FS_TRY(conv_geom(H, W, KH, KW, stride, padding_same, &wa.OH, &wa.OW, &wa.pad_t, &wa.pad_l));
    wa.in = x; wa.dy = dy; wa.out = dw; wa.partial = scratch; wa.partial_cap = scratch_floats;
    wa.H = H; wa.W = W; wa.C = C; wa.in_bs = (long long)H * W * C;
    wa.KH = KH; wa.KW = KW; wa.stride = stride; wa.OC = OC; wa.dy_bs = (long long)wa.OH * wa.OW * OC;
    wa.N = N; wa.per_sample = 0; wa.scale = 1.f;
    return launch_wgrad(wa, S(stream));
}

int fs_upconv2d_forward(const float* x, const float* w, float* y, float* wc_scratch, int N, int H, int W,
                        int C, int OC, void* stream) {
    FS_CHECK(x && w && y && wc_scratch, "fs_upconv2d_forward: NULL argument");
    FS_TRY(upconv_collapse(w, wc_scratch, C, OC, S(stream)));
    IGemmArgs a;
    memset(&a, 0, sizeof(a));
    a.in = x; a.w = wc_scratch; a.out = y; a.N = N; a.H = H; a.W = W; a.C = C; a.in_bs = (long long)H * W * C;
    a.KH = a.KW = 2; a.stride = 1; a.OH = H; a.OW = W; a.OC = 4 * OC; a.out_mode = 1;
    a.out_bs = 4LL * H * W * OC;
    return launch_igemm(a, S(stream));
}

int fs_instnorm_forward(const float* x, const float* scale, const float* shift, float* y, float* stats,
                        double* scratch, int N, int H, int W, int C, float eps, int act, void* stream) {
    FS_CHECK(x && scale && shift && y && stats && scratch, "fs_instnorm_forward: NULL argument");
    FS_TRY(instnorm_stats(x, stats, stats + (long long)N * C, N, H * W, C, eps, scratch, S(stream)));
    return instnorm_apply(x, stats, stats + (long long)N * C, scale, shift, nullptr, y, N, H, W, C, act, 0, S(stream));
}

int fs_maxpool2x2(const float* x, float* y, int N, int H, int W, int C, void* stream) {
    FS_CHECK(x && y, "fs_maxpool2x2: NULL argument");
    return maxpool2x2_fwd(x, y, N, H, W, C, S(stream));
}

long long fs_gram_scratch_floats(int N, int C) { return wgrad_partial_floats(C, C, N); }

int fs_gram_forward(const float* f, float* g, float* scratch, long long scratch_floats, int N, int H, int W,
                    int C, void* stream) {
    FS_CHECK(f && g && scratch, "fs_gram_forward: NULL argument");
    WGradArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.in = f; wa.dy = f; wa.out = g; wa.partial = scratch; wa.partial_cap = scratch_floats;
    wa.H = H; wa.W = W; wa.C = C; wa.in_bs = (long long)H * W * C;
    wa.KH = wa.KW = 1; wa.stride = 1; wa.OH = H; wa.OW = W; wa.OC = C; wa.dy_bs = wa.in_bs;
    wa.N = N; wa.per_sample = 1; wa.scale = (float)(1.0 / ((double)H * W * C));
    return launch_wgrad(wa, S(stream));
}

// loss single ops: acc is a device double (zeroed by the caller), out a device float
int fs_loss_sqdiff(const float* a, const float* b, long long n, double scale, double* acc, float* out, void* stream) {
    FS_CHECK(a && b && acc && out && n % 4 == 0, "fs_loss_sqdiff: bad argument (n must be a multiple of 4)");
    FS_TRY(fill_zero(acc, 4 * sizeof(double), S(stream)));
    FS_TRY(sqdiff_sum(a, b, n, scale, acc, S(stream)));
    return finalize_losses(acc, out, S(stream));
}
int fs_loss_style(const float* G, const float* T, int N, int CC, double scale, double* acc, float* out, void* stream) {
    FS_CHECK(G && T && acc && out, "fs_loss_style: NULL argument");
    FS_TRY(fill_zero(acc, 4 * sizeof(double), S(stream)));
    FS_TRY(style_loss_grad(G, T, nullptr, N, CC, 0.f, scale, acc + 1, S(stream)));
    return finalize_losses(acc, out, S(stream));
}
int fs_frame_u8_to_f32(const unsigned char* in, float* out, long long n, void* stream) {
    FS_CHECK(in && out && n > 0, "fs_frame_u8_to_f32: bad argument");
    return frame_u8_to_f32(in, out, n, S(stream));
}
int fs_frame_f32_to_u8(const float* in, unsigned char* out, long long npix, int swap_rb, void* stream) {
    FS_CHECK(in && out && npix > 0, "fs_frame_f32_to_u8: bad argument");
    return frame_f32_to_u8(in, out, npix, swap_rb, S(stream));
}

int fs_resize_bicubic_tf1_u8(const unsigned char* in, float* out, int H, int W, int OH, int OW, void* stream) {
    FS_CHECK(in && out, "fs_resize_bicubic_tf1_u8: NULL argument");
    return resize_bicubic_tf1_u8(in, out, H, W, OH, OW, S(stream));
}

int fs_loss_tv(const float* Y3, int N, int H, int W, double* acc, float* out, void* stream) {
    FS_CHECK(Y3 && acc && out, "fs_loss_tv: NULL argument");
    FS_TRY(fill_zero(acc, 4 * sizeof(double), S(stream)));
    FS_TRY(tv_loss_grad(Y3, nullptr, N, H, W, 1.f, acc + 2, S(stream)));
    return finalize_losses(acc, out, S(stream));
}

// ------------------------------------------------------------------ tensor-core path
static size_t al1k(size_t b) { return (b + 1023) & ~(size_t)1023; }

size_t fs_conv3x3_tc_scratch_bytes(int N, int H, int W, int C, int OC) {
    size_t act = (size_t)N * H * W * (C > OC ? C : OC) * 2;      // one bf16 plane of the larger tensor
    size_t wts = (size_t)9 * C * OC * 2;
    return 2 * al1k(act) + 2 * al1k(wts) + 1024;
}

static int tc_scratch(void* scratch, size_t bytes, size_t act_elems, size_t w_elems, SplitPtr* xs, SplitPtr* ws) {
    FS_CHECK(scratch && ((uintptr_t)scratch & 1023) == 0, "tc scratch must be 1024-byte aligned");
    size_t a = al1k(act_elems * 2), w = al1k(w_elems * 2);
    FS_CHECK(bytes >= 2 * a + 2 * w, "tc scratch too small");
    char* p = (char*)scratch;
    xs->hi = (__nv_bfloat16*)p; xs->lo = (__nv_bfloat16*)(p + a);
    ws->hi = (__nv_bfloat16*)(p + 2 * a); ws->lo = (__nv_bfloat16*)(p + 2 * a + w);
    return 0;
}

int fs_conv3x3_tc_forward(const float* x, const float* w, const float* bias, float* y, void* scratch,
                          size_t scratch_bytes, int N, int H, int W, int C, int OC, int padding_same,
                          int relu, void* stream) {
    FS_CHECK(x && w && y, "fs_conv3x3_tc_forward: NULL argument");
    Conv3x3TcArgs a;
    memset(&a, 0, sizeof(a));
    FS_TRY(tc_scratch(scratch, scratch_bytes, (size_t)N * H * W * C, (size_t)9 * C * OC, &a.x, &a.w));
    FS_TRY(split_bf16(x, a.x, (long long)N * H * W * C, S(stream)));
    FS_TRY(pack_w3x3_tc(w, a.w, C, OC, 0, S(stream)));
    a.N = N; a.H = H; a.W = W; a.C = C; a.OC = OC;
    a.pad = padding_same ? 1 : 0;
    a.OH = padding_same ? H : H - 2; a.OW = padding_same ? W : W - 2;
    a.bias = bias; a.relu = relu; a.out_f32 = y;
    return launch_conv3x3_tc(a, S(stream));
}

int fs_conv3x3_tc_dgrad(const float* dy, const float* w, float* dx, void* scratch, size_t scratch_bytes,
                        int N, int H, int W, int C, int OC, int padding_same, void* stream) {
    FS_CHECK(dy && w && dx, "fs_conv3x3_tc_dgrad: NULL argument");
    const int OH = padding_same ? H : H - 2, OW = padding_same ? W : W - 2;
    Conv3x3TcArgs a;
    memset(&a, 0, sizeof(a));
    FS_TRY(tc_scratch(scratch, scratch_bytes, (size_t)N * OH * OW * OC, (size_t)9 * C * OC, &a.x, &a.w));
    FS_TRY(split_bf16(dy, a.x, (long long)N * OH * OW * OC, S(stream)));
    FS_TRY(pack_w3x3_tc(w, a.w, C, OC, 1, S(stream)));
    a.N = N; a.H = OH; a.W = OW; a.C = OC;          // gathered tensor = dy
    a.OH = H; a.OW = W; a.OC = C;                   // produces dx
    a.pad = padding_same ? 1 : 2;                   // k-1-pad
    a.out_f32 = dx;
    return launch_conv3x3_tc(a, S(stream));
}

/* test entry: weight gradient of a 3x3 stride-1 64->64 conv on the tensor path */
int fs_wgrad3x3_tc(const float* x, const float* dy, float* dw, void* scratch, size_t scratch_bytes, int N, int H,
                   int W, int padding_same, void* stream) {
    FS_CHECK(x && dy && dw && scratch && ((uintptr_t)scratch & 1023) == 0, "fs_wgrad3x3_tc: bad argument");
    const int OH = padding_same ? H : H - 2, OW = padding_same ? W : W - 2;
    size_t xa = al1k((size_t)N * H * W * 64 * 2), da = al1k((size_t)N * OH * OW * 64 * 2);
    size_t pa = (size_t)wgrad3x3_tc_partial_floats() * 4;
    FS_CHECK(scratch_bytes >= 2 * xa + 2 * da + pa, "fs_wgrad3x3_tc: scratch too small");
    char* p = (char*)scratch;
    SplitPtr xs{(__nv_bfloat16*)p, (__nv_bfloat16*)(p + xa)};
    SplitPtr ds{(__nv_bfloat16*)(p + 2 * xa), (__nv_bfloat16*)(p + 2 * xa + da)};
    float* partial = (float*)(p + 2 * xa + 2 * da);
    FS_TRY(split_bf16(x, xs, (long long)N * H * W * 64, S(stream)));
    FS_TRY(split_bf16(dy, ds, (long long)N * OH * OW * 64, S(stream)));
    return launch_wgrad3x3_tc(xs, ds, dw, partial, wgrad3x3_tc_partial_floats(), N, H, W, OH, OW, padding_same ? 1 : 0, S(stream));
}
size_t fs_wgrad3x3_tc_scratch_bytes(int N, int H, int W) {
    return 4 * al1k((size_t)N * H * W * 64 * 2) + (size_t)wgrad3x3_tc_partial_floats() * 4 + 1024;
}

/* test entry: Gram matrix on the tensor path (the kernel the engine uses for style taps) */
size_t fs_gram_tc_scratch_bytes(int N, int H, int W, int C) {
    return 2 * al1k((size_t)N * H * W * C * 2) + (size_t)gram_tc_partial_floats(N, H * W, C) * 4 + 1024;
}
int fs_gram_tc_forward(const float* f, float* g, void* scratch, size_t scratch_bytes, int N, int H, int W, int C,
                       void* stream) {
    FS_CHECK(f && g && scratch && ((uintptr_t)scratch & 1023) == 0, "fs_gram_tc_forward: bad argument");
    FS_CHECK(C % 64 == 0 && C >= 64, "fs_gram_tc_forward: C must be a multiple of 64");
    FS_CHECK(scratch_bytes >= fs_gram_tc_scratch_bytes(N, H, W, C), "fs_gram_tc_forward: scratch too small");
    size_t fa = al1k((size_t)N * H * W * C * 2);
    char* p = (char*)scratch;
    SplitPtr fs_{(__nv_bfloat16*)p, (__nv_bfloat16*)(p + fa)};
    float* partial = (float*)(p + 2 * fa);
    FS_TRY(split_bf16(f, fs_, (long long)N * H * W * C, S(stream)));
    return launch_gram_tc(fs_, g, partial, gram_tc_partial_floats(N, H * W, C), N, H * W, C,
                          (float)(1.0 / ((double)H * W * C)), S(stream));
}

}  // extern "C"
