// Whole-path composites: transform net forward/backward, VGG16 perceptual loss
// forward/backward, and the fused train step.  Host-side orchestration only;
// the arithmetic is in igemm_f32.cu / ops.cu / (tensor path) conv3x3_tc.cu.
//
// Reference call sites being replaced (one Session.run each in the reference):
//   transform fwd      stylize_image.py:75, train.py:161
//   train step         train.py:250-251 (content targets) + train.py:274-275
//   slow_style step    slow_style.py:170-176
#include "engine.cuh"
#include <algorithm>

namespace fs {

static const float IN_EPS = 1e-3f;       // reference im_transf_net.py:218
static bool s2_collapsed(const TConv& c);
// 9x9 stride-1 SAME layers with a 4x16 / 16x4 channel block run on the direct shared-memory kernels
static bool direct9(const TConv& c) { return c.k == 9 && c.stride == 1 && c.same && !c.upconv && c.cin_s * c.cout_s == 64; }

// ---------------------------------------------------------------- live per-kernel timing
// Optional CUDA-event instrumentation of the GEMM-class launches of a step (bench.py roofline).
int Engine::pbegin(int cat, double flops, cudaStream_t st) {
    if (!prof_on) return 0;
    cudaEvent_t a, b;
    FS_CUDA(cudaEventCreate(&a));
    FS_CUDA(cudaEventCreate(&b));
    FS_CUDA(cudaEventRecord(a, st));
    prof_ev.push_back(a); prof_ev.push_back(b); prof_cat.push_back(cat); prof_flops.push_back(flops);
    prof_bytes.push_back(next_bytes); next_bytes = 0.0;
    return 0;
}
int Engine::pend(cudaStream_t st) {
    if (!prof_on) return 0;
    FS_CUDA(cudaEventRecord(prof_ev.back(), st));
    return 0;
}
int Engine::prof_records(int max_rec, int* cat, float* ms, double* flops, int* count) {
    // per-launch records in issue order (does not reset; call before prof_read)
    int n = (int)prof_cat.size();
    *count = n;
    for (int i = 0; i < n && i < max_rec; ++i) {
        float t = 0.f;
        FS_CUDA(cudaEventSynchronize(prof_ev[2 * i + 1]));
        FS_CUDA(cudaEventElapsedTime(&t, prof_ev[2 * i], prof_ev[2 * i + 1]));
        cat[i] = prof_cat[i]; ms[i] = t; flops[i] = prof_flops[i];
    }
    return 0;
}
int Engine::prof_read(int ncat, float* ms, double* flops, int* launches) {
    for (int i = 0; i < ncat; ++i) { ms[i] = 0.f; flops[i] = 0.0; launches[i] = 0; }
    for (size_t i = 0; i < prof_cat.size(); ++i) {
        float t = 0.f;
        FS_CUDA(cudaEventSynchronize(prof_ev[2 * i + 1]));
        FS_CUDA(cudaEventElapsedTime(&t, prof_ev[2 * i], prof_ev[2 * i + 1]));
        int c = prof_cat[i];
        if (c >= 0 && c < ncat) { ms[c] += t; flops[c] += prof_flops[i]; launches[c] += 1; }
        cudaEventDestroy(prof_ev[2 * i]); cudaEventDestroy(prof_ev[2 * i + 1]);
    }
    prof_ev.clear(); prof_cat.clear(); prof_flops.clear(); prof_bytes.clear();
    return 0;
}
int Engine::prof_read_bytes(int ncat, double* bytes) {
    for (int i = 0; i < ncat; ++i) bytes[i] = 0.0;
    for (size_t i = 0; i < prof_cat.size(); ++i)
        if (prof_cat[i] >= 0 && prof_cat[i] < ncat) bytes[prof_cat[i]] += prof_bytes[i];
    return 0;
}
#define PROF(cat, flops, call)              \
    do {                                    \
        FS_TRY(pbegin((cat), (flops), st)); \
        FS_TRY(call);                       \
        FS_TRY(pend(st));                   \
    } while (0)
// algorithmic bytes of a tensor-path conv launch: every operand plane read once, every output plane written once
static double tc_bytes(const Conv3x3TcArgs& a) {
    const int taps = a.one_by_one ? 1 : (a.taps ? a.taps : 3);
    double b = 4.0 * a.N * a.H * a.W * (double)a.C;                                   // hi + lo input planes
    b += 4.0 * taps * taps * (double)a.C * a.OC * (a.one_by_one && a.per_sample_w ? a.N : 1);
    const double out = 4.0 * a.N * a.OH * a.OW * (double)a.OC;
    if (a.out_f32) b += out;
    if (a.out_split.hi) b += out;
    if (a.ref) b += out;
    if (a.ref_code) b += out / 4.0;
    if (a.out_code) b += out / 4.0;
    if (a.ctarget) b += out;
    if (a.pool_grad) b += out / 4.0;
    if (a.pool_split.hi) b += out / 4.0;
    if (a.addend) b += 4.0 * a.N * a.addH * a.addW * (double)a.OC;
    return b;
}
#define PROFB(cat, flops, bytes, call) \
    do {                               \
        next_bytes = prof_on ? (bytes) : 0.0; \
        PROF(cat, flops, call);        \
    } while (0)
static double igemm_flops(const IGemmArgs& a) { return 2.0 * a.N * a.OH * a.OW * (double)a.OC * a.KH * a.KW * a.C; }
static double tc_flops(const Conv3x3TcArgs& a) { return 2.0 * a.N * a.OH * a.OW * (double)a.OC * 9.0 * a.C; }
static double wgrad_flops(const WGradArgs& a) { return 2.0 * a.N * a.OH * a.OW * (double)a.OC * a.KH * a.KW * a.C; }

// ---------------------------------------------------------------- parameter table
void transform_param_table(TConv* tc) {
    auto set = [&](int i, int k, int s, int same, int cin, int cout, int up, int act) {
        tc[i].k = k; tc[i].stride = s; tc[i].same = same; tc[i].cin = cin; tc[i].cout = cout;
        tc[i].upconv = up; tc[i].act = act;
        tc[i].cin_s = (cin + 3) & ~3; tc[i].cout_s = (cout + 3) & ~3;
    };
    set(0, 9, 1, 1, 3, 16, 0, ACT_RELU);
    set(1, 3, 2, 1, 16, 32, 0, ACT_RELU);
    set(2, 3, 2, 1, 32, 64, 0, ACT_RELU);
    for (int r = 0; r < 5; ++r) {
        set(3 + 2 * r, 3, 1, 0, 64, 64, 0, ACT_RELU);
        set(4 + 2 * r, 3, 1, 0, 64, 64, 0, ACT_NONE);
    }
    set(13, 3, 2, 1, 64, 32, 1, ACT_RELU);
    set(14, 3, 2, 1, 32, 16, 1, ACT_RELU);
    set(15, 9, 1, 1, 16, 3, 0, ACT_TANH255);
    // flat layout == byte-sorted checkpoint key order
    long long off = 0;
    auto wsz = [&](int i) { return (long long)tc[i].k * tc[i].k * tc[i].cin * tc[i].cout; };
    for (int i = 0; i < 3; ++i) {
        tc[i].offG = off; off += tc[i].cout;
        tc[i].offB = off; off += tc[i].cout;
        tc[i].offW = off; off += wsz(i);
    }
    for (int r = 0; r < 5; ++r) {
        int a = 3 + 2 * r, b = 4 + 2 * r;
        tc[a].offG = off; off += 64;      // INscale1
        tc[b].offG = off; off += 64;      // INscale2
        tc[a].offB = off; off += 64;      // INshift1
        tc[b].offB = off; off += 64;      // INshift2
        tc[a].offW = off; off += wsz(a);  // W1
        tc[b].offW = off; off += wsz(b);  // W2
    }
    for (int i = 13; i < 16; ++i) {
        tc[i].offG = off; off += tc[i].cout;
        tc[i].offB = off; off += tc[i].cout;
        tc[i].offW = off; off += wsz(i);
    }
}

// ---------------------------------------------------------------- VGG tables
static const int V_CIN[V_NCONV] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512};
static const int V_COUT[V_NCONV] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512};
static const int V_POOL[V_NCONV] = {0, 1, 0, 1, 0, 0, 1, 0, 0, 0};

long long vgg_flat_floats() {
    long long n = 0;
    for (int l = 0; l < V_NCONV; ++l) n += 9LL * V_CIN[l] * V_COUT[l] + V_COUT[l];
    return n;
}
static void vgg_offsets(VConv* vc) {
    long long off = 0;
    for (int l = 0; l < V_NCONV; ++l) {
        vc[l].cin = V_CIN[l]; vc[l].cout = V_COUT[l]; vc[l].cin_s = (V_CIN[l] + 3) & ~3;
        vc[l].pool_after = V_POOL[l];
        vc[l].offW = off; off += 9LL * vc[l].cin_s * vc[l].cout;
        vc[l].offB = off; off += vc[l].cout;
        vc[l].offWT = off; off += 9LL * vc[l].cin_s * vc[l].cout;
        vc[l].offTcF = vc[l].offTcD = -1;
        if (l >= 1) {       // two bf16 planes of 9*cin*cout elements = 9*cin*cout floats per packing
            vc[l].offTcF = off; off += 9LL * vc[l].cin * vc[l].cout;
            vc[l].offTcD = off; off += 9LL * vc[l].cin * vc[l].cout;
        }
    }
}

static SplitPtr vgg_tc_w(const float* packed, const VConv& v, int mode) {
    SplitPtr s;
    s.hi = reinterpret_cast<__nv_bfloat16*>(const_cast<float*>(packed + (mode == 0 ? v.offTcF : v.offTcD)));
    s.lo = s.hi + 9LL * v.cin * v.cout;
    return s;
}
long long vgg_packed_floats() {
    VConv vc[V_NCONV];
    vgg_offsets(vc);
    const VConv& v = vc[V_NCONV - 1];
    return v.offTcD + 9LL * v.cin * v.cout;
}

int vgg_pack(const float* flat, float* packed, cudaStream_t st) {
    VConv vc[V_NCONV];
    vgg_offsets(vc);
    long long src = 0;
    for (int l = 0; l < V_NCONV; ++l) {
        long long wn = 9LL * vc[l].cin * vc[l].cout;
        if (vc[l].cin != vc[l].cin_s)
            FS_TRY(pad_taps(flat + src, packed + vc[l].offW, 9, vc[l].cin, vc[l].cout, vc[l].cin_s, vc[l].cout, st));
        else
            FS_CUDA(cudaMemcpyAsync(packed + vc[l].offW, flat + src, wn * sizeof(float), cudaMemcpyDeviceToDevice, st));
        src += wn;
        FS_CUDA(cudaMemcpyAsync(packed + vc[l].offB, flat + src, vc[l].cout * sizeof(float), cudaMemcpyDeviceToDevice, st));
        src += vc[l].cout;
        if (l == 0)      // conv1_1 runs on the direct kernels: its data gradient is a conv with flipped weights
            FS_TRY(flip_transpose_taps(packed + vc[l].offW, packed + vc[l].offWT, 9, vc[l].cin_s, vc[l].cout, st));
        else
            FS_TRY(transpose_taps(packed + vc[l].offW, packed + vc[l].offWT, 9, vc[l].cin_s, vc[l].cout, st));
        if (l >= 1) {
            FS_TRY(pack_w3x3_tc(packed + vc[l].offW, vgg_tc_w(packed, vc[l], 0), vc[l].cin, vc[l].cout, 0, st));
            FS_TRY(pack_w3x3_tc(packed + vc[l].offW, vgg_tc_w(packed, vc[l], 1), vc[l].cin, vc[l].cout, 1, st));
        }
    }
    return 0;
}

// ---------------------------------------------------------------- planning
int Engine::plan() {
    FS_CHECK(N >= 1 && H >= 1 && W >= 1, "engine: bad dims N=%d H=%d W=%d", N, H, W);
    transform_param_table(tc);
    VH = H; VW = W;
    if (flags & ENG_TRANSFORM) {
        FS_CHECK(H > 40 && W > 40, "transform net needs H,W > 40 for the 40-px reflect pad (got %dx%d)", H, W);
        Hp = H + 80; Wp = W + 80;
        int h = Hp, w = Wp;
        for (int l = 0; l < T_NCONV; ++l) {
            TConv& c = tc[l];
            c.inH = h; c.inW = w;
            if (c.upconv) {
                c.outH = 2 * h; c.outW = 2 * w; c.pad_t = c.pad_l = 0;
            } else if (c.same) {
                tf_same(h, c.k, c.stride, &c.outH, &c.pad_t);
                tf_same(w, c.k, c.stride, &c.outW, &c.pad_l);
            } else {
                c.outH = h - c.k + 1; c.outW = w - c.k + 1; c.pad_t = c.pad_l = 0;
            }
            FS_CHECK(c.outH > 0 && c.outW > 0, "transform net: image too small (layer %d output %dx%d)", l, c.outH, c.outW);
            h = c.outH; w = c.outW;
        }
        OH = h; OW = w;
        VH = OH; VW = OW;
    }
    if (flags & ENG_VGG) {
        vgg_offsets(vc);
        int h = VH, w = VW;
        for (int l = 0; l < V_NCONV; ++l) {
            vc[l].H = h; vc[l].W = w;
            if (vc[l].pool_after) { h = (h + 1) / 2; w = (w + 1) / 2; }
        }
    }
    return 0;
}

static long long maxll(long long a, long long b) { return a > b ? a : b; }

void Engine::layout(Arena& a) {
    const bool tb_ = flags & ENG_TRANSFORM, tbw = flags & ENG_TRANSFORM_BWD;
    const bool vg = flags & ENG_VGG, vgb = flags & ENG_VGG_BWD;
    long long wgcap = 0;
    if (tb_) {
        xpad4 = a.take<float>((long long)N * Hp * Wp * 4);
        long long maxact = 0;
        for (int l = 0; l < T_NCONV; ++l) {
            TConv& c = tc[l];
            long long n = (long long)N * c.outH * c.outW * c.cout_s;
            maxact = maxll(maxact, n);
            tb[l].raw = a.take<float>(n);
            tb[l].act = l == T_NCONV - 1 ? nullptr : a.take<float>(n);
            tb[l].mean = a.take<float>((long long)N * c.cout_s);
            tb[l].rstd = a.take<float>((long long)N * c.cout_s);
            weff[l] = nullptr; wefft[l] = nullptr;
            if (l == 0 || l == 15) weff[l] = a.take<float>((long long)c.k * c.k * c.cin_s * c.cout_s);
            if (c.upconv) weff[l] = a.take<float>(16LL * c.cin * c.cout);
            if (tbw && l > 0) {
                long long n2 = (c.upconv || c.stride == 2) ? 16LL * c.cin * c.cout : (long long)c.k * c.k * c.cin_s * c.cout_s;
                wefft[l] = a.take<float>(n2);
                }
            if (tbw) {
                int K = c.upconv ? 4 * c.cin : c.k * c.k * c.cin_s;
                int OC = c.upconv ? 4 * c.cout : c.cout_s;
                wgcap = maxll(wgcap, wgrad_partial_floats(K, OC, 1));
                if (c.k == 9) wgcap = maxll(wgcap, maxll(wgrad9x9_partial_floats(N, c.inH, c.inW), wgrad9_x8_partial_floats()));
                if ((flags & ENG_DECONV) && c.upconv) wgcap = maxll(wgcap, wgrad_partial_floats(9 * c.cout, c.cin, 1));
                if (l >= 3 && l <= 12) wgcap = maxll(wgcap, wgrad3x3_tc_partial_floats());
                if (tc2_layout(l)) wgcap = maxll(wgcap, wgrad2x2_tc_partial_floats());
            }
        }
        for (int l = 0; l < T_NCONV; ++l) {
            tsplit[l].hi = tsplit[l].lo = nullptr; tw_f[l].hi = tw_f[l].lo = nullptr; tw_d[l].hi = tw_d[l].lo = nullptr;
            rg[l].hi = rg[l].lo = nullptr;
            if (tbw && l >= 3 && l <= 12) {
                long long n = (long long)N * tc[l].outH * tc[l].outW * 64;
                rg[l].hi = a.take<__nv_bfloat16>(n); rg[l].lo = a.take<__nv_bfloat16>(n);
            }
            if (l >= 3 && l <= 12) {
                long long n = (long long)N * tc[l].inH * tc[l].inW * 64;
                tsplit[l].hi = a.take<__nv_bfloat16>(n); tsplit[l].lo = a.take<__nv_bfloat16>(n);
                tw_f[l].hi = a.take<__nv_bfloat16>(9 * 64 * 64); tw_f[l].lo = a.take<__nv_bfloat16>(9 * 64 * 64);
                if (tbw) { tw_d[l].hi = a.take<__nv_bfloat16>(9 * 64 * 64); tw_d[l].lo = a.take<__nv_bfloat16>(9 * 64 * 64); }
            }
            if (tc2_layout(l)) {
                long long n = (long long)N * tc[l].inH * tc[l].inW * tc[l].cin;
                const long long wn = 4LL * 128 * 64;                                    // 4 taps x K x N, all four layers
                tsplit[l].hi = a.take<__nv_bfloat16>(n); tsplit[l].lo = a.take<__nv_bfloat16>(n);
                tw_f[l].hi = a.take<__nv_bfloat16>(wn); tw_f[l].lo = a.take<__nv_bfloat16>(wn);
                if (tbw) { tw_d[l].hi = a.take<__nv_bfloat16>(wn); tw_d[l].lo = a.take<__nv_bfloat16>(wn); }
            }
        }
        w2f = a.take<float>(4LL * 128 * 64);
        wpair = a.take<float>(4LL * 128 * 64);
        {   // 9x9 layers on the tensor path (tc9): zero-margined input planes + expanded / packed weights
            // (the 4-channel planes hold either form: x16 [H, W + 16, 4] or x8 [H, W / 8, 64])
            const long long n9[3] = {(long long)N * Hp * maxll((Wp + 16) * 4, Wp * 8), (long long)N * OH * (OW + 16) * 16,
                                     (long long)N * OH * maxll((OW + 16) * 4, OW * 8)};
            for (int i = 0; i < 3; ++i) {
                const bool need = i < 2 || tbw;
                p9[i].hi = need ? a.take<__nv_bfloat16>(n9[i]) : nullptr;
                p9[i].lo = need ? a.take<__nv_bfloat16>(n9[i]) : nullptr;
                tw9[i].hi = need ? a.take<__nv_bfloat16>(18LL * 64 * 256) : nullptr;
                tw9[i].lo = need ? a.take<__nv_bfloat16>(18LL * 64 * 256) : nullptr;
            }
        }
        if (tbw) {         // plain split planes for the tensor-path 9x9 weight gradients (Engine::wg9)
            const long long n0 = (long long)N * Hp * Wp * 16, n14 = (long long)N * OH * OW * 16;
            g0split.hi = a.take<__nv_bfloat16>(n0); g0split.lo = a.take<__nv_bfloat16>(n0);
            a14split.hi = a.take<__nv_bfloat16>(n14); a14split.lo = a.take<__nv_bfloat16>(n14);
        }
        w2f_b = a.take<float>(4LL * 128 * 64);
        for (int i = 0; i < 4; ++i) wpair_x[i] = a.take<float>(4LL * 128 * 64);
        for (int l = 0; l < T_NCONV; ++l) wgs[l] = nullptr;
        if (tbw) {
            for (int l : {0, 15}) wgs[l] = a.take<float>(81LL * 64);
            for (int l : {1, 2, 13, 14}) wgs[l] = a.take<float>(4LL * 128 * 64);
            wgu[0] = a.take<float>(4LL * 128 * 64);
            wgu[1] = a.take<float>(4LL * 128 * 64);
        }
        y3 = a.take<float>((long long)N * OH * OW * 3);
        in_partial = a.take<double>((long long)N * 64 * 64 * 2);
        sums_n = (long long)STATS_REPLICAS * N * 64 * 2;
        for (int i = 0; i < 2; ++i) in_sums2[i] = a.take<double>(sums_n);
        for (int i = 0; i < 2; ++i) bw_sums2[i] = tbw ? a.take<double>(sums_n) : nullptr;
        in15 = a.take<float>(8);
        wtmp15 = a.take<float>(81 * 64);
        if (tbw) {
            m12 = a.take<float>((long long)N * 64 * 2);
            for (int i = 0; i < 3; ++i) tgrad[i] = a.take<float>(maxact);
            {   // split companions hold dRaw of the tensor-path convs: residual (3..12), initconv_2, upsample_0
                long long nres = (long long)N * tc[3].outH * tc[3].outW * 64;
                for (int l : {1, 2, 13, 14}) nres = maxll(nres, (long long)N * tc[l].outH * tc[l].outW * tc[l].cout);
                for (int i = 0; i < 3; ++i) { tgsplit[i].hi = a.take<__nv_bfloat16>(nres); tgsplit[i].lo = a.take<__nv_bfloat16>(nres); }
            }
            wg_tmp = a.take<float>(81LL * 16 * 4 + 16LL * 64 * 32 + 1024);
            wg_tmp2 = a.take<float>(16LL * 64 * 32);
            gb_tmp = a.take<float>(8);
        }
    }
    if (vg) {
        v_in4 = a.take<float>((long long)N * VH * VW * 4);
        long long maxact = 0;
        for (int l = 0; l < V_NCONV; ++l) {
            long long n = (long long)N * vc[l].H * vc[l].W * vc[l].cout;
            maxact = maxll(maxact, n);
            vact[l] = a.take<float>(n);
            vpool[l] = vc[l].pool_after ? a.take<float>((long long)N * ((vc[l].H + 1) / 2) * ((vc[l].W + 1) / 2) * vc[l].cout) : nullptr;
            bool st_ = (style_mask >> l) & 1, ct_ = (content_mask >> l) & 1;
            gram[l] = st_ ? a.take<float>((long long)N * vc[l].cout * vc[l].cout) : nullptr;
            gramS[l] = st_ ? a.take<float>((long long)N * vc[l].cout * vc[l].cout) : nullptr;
            ctarget[l] = ct_ ? a.take<float>(n) : nullptr;
            if (st_) wgcap = maxll(wgcap, wgrad_partial_floats(vc[l].cout, vc[l].cout, N));
            vtsplit[l].hi = vtsplit[l].lo = gsS[l].hi = gsS[l].lo = nullptr;
            if (st_ && l >= 1) {
                vtsplit[l].hi = a.take<__nv_bfloat16>(n); vtsplit[l].lo = a.take<__nv_bfloat16>(n);
                long long ncc = (long long)N * vc[l].cout * vc[l].cout;
                gsS[l].hi = a.take<__nv_bfloat16>(ncc); gsS[l].lo = a.take<__nv_bfloat16>(ncc);
                wgcap = maxll(wgcap, gram_tc_partial_floats(N, vc[l].H * vc[l].W, vc[l].cout));
            }
            vsplit[l].hi = vsplit[l].lo = nullptr;
            if (l >= 1) {
                long long nin = (long long)N * vc[l].H * vc[l].W * vc[l].cin;
                vsplit[l].hi = a.take<__nv_bfloat16>(nin); vsplit[l].lo = a.take<__nv_bfloat16>(nin);
            }
        }
        loss_acc = a.take<double>(4);
        for (int l = 0; l < V_NCONV; ++l)
            vcode[l] = vgb ? a.take<unsigned char>((long long)N * vc[l].H * vc[l].W * vc[l].cout) : nullptr;
        if (vgb) {
            vgrad_floats = maxact;
            for (int i = 0; i < 4; ++i) {
                vgrad[i] = a.take<float>(maxact);
                vgsplit[i].hi = a.take<__nv_bfloat16>(maxact); vgsplit[i].lo = a.take<__nv_bfloat16>(maxact);
            }
            dY4 = a.take<float>((long long)N * VH * VW * 4);
        }
    }
    wg_partial_cap = wgcap;
    wg_partial = wgcap ? a.take<float>(wgcap) : nullptr;
}

int Engine::bind(void* ws, size_t bytes) {
    FS_CHECK(bytes >= ws_bytes, "engine: workspace too small (%zu < %zu bytes)", bytes, ws_bytes);
    FS_CHECK(((uintptr_t)ws & 255) == 0, "engine: workspace must be 256-byte aligned");
    Arena a; a.base = (char*)ws; a.cap = bytes;
    layout(a);
    for (double* sp : {in_sums2[0], in_sums2[1], bw_sums2[0], bw_sums2[1]})
        if (sp) FS_CUDA(cudaMemset(sp, 0, (size_t)sums_n * sizeof(double)));
    FS_CUDA(cudaDeviceSynchronize());
    if (flags & ENG_TRANSFORM) {         // zero margins of the x16 planes (never written afterwards)
        const size_t n9[3] = {(size_t)N * Hp * (size_t)maxll((Wp + 16) * 4, Wp * 8), (size_t)N * OH * (OW + 16) * 16,
                              (size_t)N * OH * (size_t)maxll((OW + 16) * 4, OW * 8)};
        for (int i = 0; i < 3; ++i)
            if (p9[i].hi) {
                FS_CUDA(cudaMemset(p9[i].hi, 0, n9[i] * sizeof(__nv_bfloat16)));
                FS_CUDA(cudaMemset(p9[i].lo, 0, n9[i] * sizeof(__nv_bfloat16)));
            }
        FS_CUDA(cudaDeviceSynchronize());
    }
    bound = true;
    return 0;
}

// ---------------------------------------------------------------- transform net
// the table-driven path covers the standard configuration: resize upsampling, tensor path on, all four stride-2 /
// resize layers in their tcgen05 2x2 forms (even sizes); everything else keeps one launch per transform
bool Engine::use_fast_prep() const {
    return fast_prep && use_tc && !(flags & ENG_DECONV) && tc2(1) && tc2(2) && tc2(13) && tc2(14);
}

// 9x9 layers on the tensor path: standard (resize) model, table-driven preparation, widths in whole 16-pixel groups
bool Engine::tc9() const {
    return tc9_on && use_fast_prep() && Wp % 16 == 0 && OW % 16 == 0 && tc[0].outW == Wp && tc[15].inW == OW;
}

// which: 0 = initconv_0 forward (x = xpad4), 1 = upsample_2 forward (x = activation of upsample_1),
//        2 = upsample_2 data gradient (x = gradient w.r.t. its raw output)
int Engine::tc9_conv(int which, const float* src_f32, float* out, bool stats, cudaStream_t st, bool planes_ready) {
    const int Hh = which == 0 ? Hp : OH, Ww = which == 0 ? Wp : OW;
    const int cin = which == 1 ? 16 : 4;                 // channels per pixel on the K side
    const int cout_px = which == 1 ? 4 : 16;             // channels per pixel on the N side
    const bool x8 = tc9_x8 && which != 1;                // 4-channel input side: 8-pixel output groups, one horizontal tap
    if (which != 0 && !planes_ready)      // (initconv_0's planes come straight from the reflect-pad kernel)
        PROF(PC_POINTWISE, 0.0, split_pad_x16(src_f32, p9[which].hi, p9[which].lo, (long long)N * Hh, Ww, cin, st, x8 ? 1 : 0));
    Conv3x3TcArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.x = p9[which]; ta.w = tw9[which];
    ta.N = N; ta.H = Hh; ta.W = x8 ? Ww / 8 : (Ww + 16) / 16; ta.C = 16 * cin;
    ta.OH = Hh; ta.OW = x8 ? Ww / 8 : Ww / 16; ta.OC = (x8 ? 8 : 16) * cout_px;
    ta.taps = 9; ta.taps_w = x8 ? 1 : 2; ta.pad = 4; ta.pad_x = 0;
    ta.out_f32 = out;
    if (stats && in_epi) { ta.stats = in_sums2[which == 0 ? 0 : 1]; ta.stats_c = cout_px; }      // layer 0 / layer 15
    const double fl = 2.0 * N * Hh * Ww * 81.0 * (which == 1 ? 16 * 3 : 3 * 16);       // algorithmic: real channel counts
    PROFB(which == 2 ? PC_TC9_DGRAD : PC_TC9_FWD, fl, tc_bytes(ta), launch_conv3x3_tc(ta, st));
    return 0;
}

int Engine::prep_transform_weights_table(const float* params, bool need_bwd, cudaStream_t st) {
    PrepPlan pl;
    auto W = [&](int l) { return params + tc[l].offW; };
    // ---- phase 0: everything that reads the raw parameters
    FS_TRY(pl.add(0, pj(PJ_PAD, W(0), weff[0], 81, 3, 16, 4, 16)));
    FS_TRY(pl.add(0, pj(PJ_PAD, W(15), weff[15], 81, 16, 3, 16, 4)));
    FS_TRY(pl.add(0, pj(PJ_UPCONV_COLLAPSE, W(13), weff[13], tc[13].cin, tc[13].cout)));
    FS_TRY(pl.add(0, pj(PJ_UPCONV_COLLAPSE, W(14), weff[14], tc[14].cin, tc[14].cout)));
    {
        PrepJob j = pj(PJ_IN15, params + tc[15].offG, in15);
        j.src2 = params + tc[15].offB;
        FS_TRY(pl.add(0, j));
    }
    for (int l = 3; l <= 12; ++l) FS_TRY(pl.add(0, pj_pack(PJ_PACK_W3X3, W(l), tw_f[l].hi, tw_f[l].lo, 64, 64, 0)));
    if (need_bwd)
        for (int l = 3; l <= 12; ++l) FS_TRY(pl.add(0, pj_pack(PJ_PACK_W3X3, W(l), tw_d[l].hi, tw_d[l].lo, 64, 64, 1)));
    FS_TRY(pl.add(0, pj(PJ_S2_FWD_COLLAPSE, W(1), w2f, tc[1].cin, tc[1].cout)));          // [4][64][32]
    FS_TRY(pl.add(0, pj(PJ_S2_FWD_COLLAPSE, W(2), w2f_b, tc[2].cin, tc[2].cout)));        // [4][128][64]
    if (need_bwd) {
        FS_TRY(pl.add(0, pj(PJ_S2_DGRAD_COLLAPSE, W(1), wefft[1], tc[1].cin, tc[1].cout)));
        FS_TRY(pl.add(0, pj(PJ_S2_DGRAD_COLLAPSE, W(2), wefft[2], tc[2].cin, tc[2].cout)));
    }
    if (tc9()) {       // Toeplitz expansion straight into the packed split-bf16 layout: [18][K/64][N][64]
        auto x16 = [&](int which, const float* src, int A, int B, int KP, int NP, int mode) {
            PrepJob j = pj(PJ_X16, src, nullptr, A, B, KP, NP, mode);
            j.hi = tw9[which].hi; j.lo = tw9[which].lo;
            return j;
        };
        const int x8 = tc9_x8 ? 2 : 0;                                       // (bit 1 of the mode word: the x8 form)
        FS_TRY(pl.add(0, x16(0, W(0), 3, 16, 4, 16, 0 | x8)));               // K = 64, N = 256 (x8: 128)
        FS_TRY(pl.add(0, x16(1, W(15), 16, 3, 16, 4, 0)));                   // K = 256, N = 64
        if (need_bwd) FS_TRY(pl.add(0, x16(2, W(15), 16, 3, 4, 16, 1 | x8)));   // K = 64, N = 256 (x8: 128)
    }
    // ---- phase 1
    FS_TRY(pl.add(1, pj(PJ_PAIR, w2f, wpair_x[0], 4 * tc[1].cin, tc[1].cout, 1, 0)));                       // -> [4][128][64]
    FS_TRY(pl.add(1, pj_pack(PJ_PACK_TAPS, w2f_b, tw_f[2].hi, tw_f[2].lo, 4, 4 * tc[2].cin, tc[2].cout, 0)));
    FS_TRY(pl.add(1, pj_pack(PJ_PACK_TAPS, weff[13], tw_f[13].hi, tw_f[13].lo, 4, tc[13].cin, 4 * tc[13].cout, 0)));
    FS_TRY(pl.add(1, pj(PJ_PAIR, weff[14], wpair_x[1], tc[14].cin, 4 * tc[14].cout, 0, 0)));                // -> [4][64][128]
    if (need_bwd) {
        FS_TRY(pl.add(1, pj(PJ_FLIP_TRANSPOSE, weff[15], wefft[15], 81, tc[15].cin_s, tc[15].cout_s)));
        FS_TRY(pl.add(1, pj(PJ_TRANSPOSE, weff[13], wefft[13], 4, tc[13].cin, 4 * tc[13].cout)));
        FS_TRY(pl.add(1, pj(PJ_TRANSPOSE, weff[14], wefft[14], 4, tc[14].cin, 4 * tc[14].cout)));
        FS_TRY(pl.add(1, pj(PJ_PAIR, wefft[1], wpair_x[2], tc[1].cout, 4 * tc[1].cin, 0, 1)));              // -> [4][64][128]
        FS_TRY(pl.add(1, pj_pack(PJ_PACK_TAPS, wefft[2], tw_d[2].hi, tw_d[2].lo, 4, tc[2].cout, 4 * tc[2].cin, 1)));
    }
    // ---- phase 2
    FS_TRY(pl.add(2, pj_pack(PJ_PACK_TAPS, wpair_x[0], tw_f[1].hi, tw_f[1].lo, 4, 128, 64, 0)));
    FS_TRY(pl.add(2, pj_pack(PJ_PACK_TAPS, wpair_x[1], tw_f[14].hi, tw_f[14].lo, 4, 64, 128, 0)));
    if (need_bwd) {
        FS_TRY(pl.add(2, pj_pack(PJ_PACK_TAPS, wpair_x[2], tw_d[1].hi, tw_d[1].lo, 4, 64, 128, 1)));
        FS_TRY(pl.add(2, pj_pack(PJ_PACK_TAPS, wefft[13], tw_d[13].hi, tw_d[13].lo, 4, 4 * tc[13].cout, tc[13].cin, 1)));
        FS_TRY(pl.add(2, pj(PJ_PAIR, wefft[14], wpair_x[3], 4 * tc[14].cout, tc[14].cin, 1, 1)));           // -> [4][128][64]
        // ---- phase 3
        FS_TRY(pl.add(3, pj_pack(PJ_PACK_TAPS, wpair_x[3], tw_d[14].hi, tw_d[14].lo, 4, 128, 64, 1)));
    }
    PROF(PC_PREP, 0.0, pl.run(st));
    return 0;
}

// weight gradients of the layers computed in a transformed layout (staged in wgs[l] by transform_backward):
// back through the pairing / collapse / padding adjoints into the flat gradient buffer, two launches
int Engine::finish_weight_grads_table(float* grads, cudaStream_t st) {
    PrepPlan pl;
    auto G = [&](int l) { return grads + tc[l].offW; };
    if (!wg9()) {      // (the tensor-path 9x9 weight gradients write their [9,9,3|16,16|3] slots directly)
        FS_TRY(pl.add(0, pj(PJ_UNPAD, wgs[15], G(15), 81, 16, 3, 16, 4)));
        FS_TRY(pl.add(0, pj(PJ_UNPAD, wgs[0], G(0), 81, 3, 16, 4, 16)));
    }
    FS_TRY(pl.add(0, pj(PJ_UNPAIR, wgs[14], wgu[0], tc[14].cin, 4 * tc[14].cout, 0, 1)));
    FS_TRY(pl.add(0, pj(PJ_UPCONV_COLLAPSE_GRAD, wgs[13], G(13), tc[13].cin, tc[13].cout)));
    FS_TRY(pl.add(0, pj(PJ_S2_FWD_COLLAPSE_GRAD, wgs[2], G(2), tc[2].cin, tc[2].cout)));
    FS_TRY(pl.add(0, pj(PJ_UNPAIR, wgs[1], wgu[1], 4 * tc[1].cin, tc[1].cout, 1, 0)));
    FS_TRY(pl.add(0, pj(PJ_COPY, gb_tmp, grads + tc[15].offG, 3)));
    FS_TRY(pl.add(0, pj(PJ_COPY, gb_tmp + 4, grads + tc[15].offB, 3)));
    FS_TRY(pl.add(1, pj(PJ_UPCONV_COLLAPSE_GRAD, wgu[0], G(14), tc[14].cin, tc[14].cout)));
    FS_TRY(pl.add(1, pj(PJ_S2_FWD_COLLAPSE_GRAD, wgu[1], G(1), tc[1].cin, tc[1].cout)));
    PROF(PC_PREP, 0.0, pl.run(st));
    return 0;
}

int Engine::prep_transform_weights(const float* params, bool need_bwd, cudaStream_t st) {
    FS_CHECK(bound && (flags & ENG_TRANSFORM), "engine has no transform plan / workspace");
    if (need_bwd) FS_CHECK(flags & ENG_TRANSFORM_BWD, "engine was not created with a backward plan");
    if (use_fast_prep()) return prep_transform_weights_table(params, need_bwd, st);
    PROF(PC_PREP, 0.0, pad_taps(params + tc[0].offW, weff[0], 81, 3, 16, 4, 16, st));
    if (flags & ENG_DECONV) {
        // 'deconv' models (reference im_transf_net.py:57-63,158-190): W is [k,k,cout,cin] and the layer is
        // tf.nn.conv2d_transpose = the data gradient of a SAME conv.  Stride 2: the 4-phase 2x2 sub-pixel form
        // (the same machinery as the stride-2 data gradients); stride 1 (9x9): a conv with flipped weights.
        PROF(PC_PREP, 0.0, s2_dgrad_collapse(params + tc[13].offW, weff[13], tc[13].cout, tc[13].cin, st));
        PROF(PC_PREP, 0.0, s2_dgrad_collapse(params + tc[14].offW, weff[14], tc[14].cout, tc[14].cin, st));
        PROF(PC_PREP, 0.0, pad_taps(params + tc[15].offW, wtmp15, 81, 3, 16, 4, 16, st));
        PROF(PC_PREP, 0.0, flip_transpose_taps(wtmp15, weff[15], 81, 4, 16, st));
    } else {
        PROF(PC_PREP, 0.0, pad_taps(params + tc[15].offW, weff[15], 81, 16, 3, 16, 4, st));
        PROF(PC_PREP, 0.0, upconv_collapse(params + tc[13].offW, weff[13], tc[13].cin, tc[13].cout, st));
        PROF(PC_PREP, 0.0, upconv_collapse(params + tc[14].offW, weff[14], tc[14].cin, tc[14].cout, st));
    }
    // 4-channel staging of the last layer's IN scale/shift (its flat slots are 3 floats, unaligned)
    FS_TRY(fill_zero(in15, 8 * sizeof(float), st));
    FS_CUDA(cudaMemcpyAsync(in15, params + tc[15].offG, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    FS_CUDA(cudaMemcpyAsync(in15 + 4, params + tc[15].offB, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (use_tc) {
        const float* wsrc[10];
        for (int l = 3; l <= 12; ++l) wsrc[l - 3] = params + tc[l].offW;
        PROF(PC_PREP, 0.0, pack_w3x3_tc_batch(wsrc, &tw_f[3], 10, 64, 64, 0, st));
        if (need_bwd) PROF(PC_PREP, 0.0, pack_w3x3_tc_batch(wsrc, &tw_d[3], 10, 64, 64, 1, st));
        if (tc2(1)) {        // 3x3 s2 16->32: space-to-depth form [4][64][32], then pixel pairing -> [4][128][64]
            PROF(PC_PREP, 0.0, s2_fwd_collapse(params + tc[1].offW, w2f, tc[1].cin, tc[1].cout, st));
            PROF(PC_PREP, 0.0, pair_taps(w2f, wpair, 4 * tc[1].cin, tc[1].cout, 1, 0, st));
            PROF(PC_PREP, 0.0, pack_taps_tc(wpair, tw_f[1], 4, 128, 64, 0, st));
        }
        if (tc2(2)) {
            PROF(PC_PREP, 0.0, s2_fwd_collapse(params + tc[2].offW, w2f, tc[2].cin, tc[2].cout, st));
            PROF(PC_PREP, 0.0, pack_taps_tc(w2f, tw_f[2], 4, 4 * tc[2].cin, tc[2].cout, 0, st));
        }
        if (tc2(13)) PROF(PC_PREP, 0.0, pack_taps_tc(weff[13], tw_f[13], 4, tc[13].cin, 4 * tc[13].cout, 0, st));
        if (tc2(14)) {       // resize-conv 32->16: collapsed [4][32][64], paired -> [4][64][128]
            PROF(PC_PREP, 0.0, pair_taps(weff[14], wpair, tc[14].cin, 4 * tc[14].cout, 0, 0, st));
            PROF(PC_PREP, 0.0, pack_taps_tc(wpair, tw_f[14], 4, 64, 128, 0, st));
        }
    }
    if (need_bwd) {
        FS_CHECK(flags & ENG_TRANSFORM_BWD, "engine was not created with a backward plan");
        for (int l = 1; l < T_NCONV; ++l) {
            const TConv& c = tc[l];
            const float* src = weff[l] ? weff[l] : params + c.offW;
            if (use_tc && l >= 3 && l <= 12) continue;          // tensor path: packed above, no fp32 transpose
            if ((flags & ENG_DECONV) && l >= 13) continue;      // conv2d_transpose layers: backward uses W as stored
            if (direct9(c)) PROF(PC_PREP, 0.0, flip_transpose_taps(src, wefft[l], c.k * c.k, c.cin_s, c.cout_s, st));
            else if (c.upconv) PROF(PC_PREP, 0.0, transpose_taps(src, wefft[l], 4, c.cin, 4 * c.cout, st));
            else if (s2_collapsed(c)) PROF(PC_PREP, 0.0, s2_dgrad_collapse(src, wefft[l], c.cin, c.cout, st));
            else PROF(PC_PREP, 0.0, transpose_taps(src, wefft[l], c.k * c.k, c.cin_s, c.cout_s, st));
        }
        // tensor-path data gradients of the collapsed stride-2 layers: the gather forms correlate with the flipped taps
        if (tc2(1)) {        // wefft[1] = [4][32][4*16] gather form, paired -> [4][64][128]
            PROF(PC_PREP, 0.0, pair_taps(wefft[1], wpair, tc[1].cout, 4 * tc[1].cin, 0, 1, st));
            PROF(PC_PREP, 0.0, pack_taps_tc(wpair, tw_d[1], 4, 64, 128, 1, st));
        }
        if (tc2(2)) PROF(PC_PREP, 0.0, pack_taps_tc(wefft[2], tw_d[2], 4, tc[2].cout, 4 * tc[2].cin, 1, st));
        if (tc2(13)) PROF(PC_PREP, 0.0, pack_taps_tc(wefft[13], tw_d[13], 4, 4 * tc[13].cout, tc[13].cin, 1, st));
        if (tc2(14)) {       // wefft[14] = [4][4*16][32] gather form over the space-to-depth view, paired -> [4][128][64]
            PROF(PC_PREP, 0.0, pair_taps(wefft[14], wpair, 4 * tc[14].cout, tc[14].cin, 1, 1, st));
            PROF(PC_PREP, 0.0, pack_taps_tc(wpair, tw_d[14], 4, 128, 64, 1, st));
        }
    }
    return 0;
}

// stride-2 3x3 SAME convs on even inputs (TF pad 0/1): data gradient runs in the 4-phase 2x2 form
static bool s2_collapsed(const TConv& c) {
    return !c.upconv && c.k == 3 && c.stride == 2 && c.pad_t == 0 && c.pad_l == 0 && c.inH % 2 == 0 && c.inW % 2 == 0;
}

bool Engine::tc2_layout(int l) const {
    if (l == 2) return s2_collapsed(tc[2]);
    if (l == 13) return !(flags & ENG_DECONV);
    if (l == 1) return s2_collapsed(tc[1]) && tc[1].outW % 2 == 0;           // paired along x
    if (l == 14) return !(flags & ENG_DECONV) && tc[14].inW % 2 == 0;
    return false;
}
bool Engine::tc2(int l) const { return use_tc && l >= 1 && l < T_NCONV && tc2_layout(l); }

// Arguments of the tensor-path launch for transform conv l (1, 2, 13, 14) in its 2x2 form.
//   forward : x = tsplit[l] (planes of the layer input),  out = raw conv output
//   backward: x = tgsplit[ri] (planes of dRaw),            out = gradient w.r.t. the layer input
// Two shapes occur (with or without pixel pairing): "s2d in -> plain out" (C = 128, OC = 64) and
// "plain in -> depth-to-space out" (C = 64, OC = 128).
int Engine::tc2_args(int l, bool bwd, int ri, float* out, Conv3x3TcArgs& ta) const {
    const TConv& c = tc[l];
    memset(&ta, 0, sizeof(ta));
    ta.N = N; ta.taps = 2; ta.pad = bwd ? 1 : 0; ta.out_f32 = out;
    ta.x = bwd ? tgsplit[ri] : tsplit[l];
    ta.w = bwd ? tw_d[l] : tw_f[l];
    const bool paired = (l == 1 || l == 14);
    const int pw = paired ? 2 : 1;                       // pixels per GEMM-space pixel
    // does this launch read a space-to-depth view?  stride-2 conv forward / resize-conv backward
    const bool s2d_in = (c.upconv != 0) == bwd;
    const int gh = c.upconv ? c.inH : c.outH, gw = (c.upconv ? c.inW : c.outW) / pw;   // GEMM-space dims
    ta.H = gh; ta.W = gw; ta.OH = gh; ta.OW = gw;
    if (s2d_in) { ta.in_s2d = 1; ta.C = 128; ta.OC = 64; }
    else { ta.C = 64; ta.OC = 128; ta.out_d2s = paired ? 2 : 1; }
    return 0;
}

// 9x9 layers: algorithmic FLOPs with the real channel counts (the padded RGB channel is skipped by the kernels)
static double conv9_flops(const TConv& c, int N) { return 2.0 * N * c.outH * c.outW * 81.0 * c.cin * c.cout; }
static double wgrad_flops_tc2(int N, int gh, int gw) { return 2.0 * N * gh * gw * 4.0 * 64.0 * 128.0; }
static double tc2_flops(const Conv3x3TcArgs& a) { return 2.0 * a.N * a.OH * a.OW * (double)a.OC * 4.0 * a.C; }

static void conv_fwd_args(const TConv& c, int N, const float* in, const float* w, float* out, IGemmArgs& a) {
    memset(&a, 0, sizeof(a));
    a.in = in; a.w = w; a.out = out; a.N = N;
    a.H = c.inH; a.W = c.inW;
    a.in_bs = (long long)c.inH * c.inW * c.cin_s;
    if (c.upconv) {
        a.C = c.cin; a.KH = a.KW = 2; a.stride = 1; a.pad_t = a.pad_l = 0;
        a.OH = c.inH; a.OW = c.inW; a.OC = 4 * c.cout; a.out_mode = 1;
        a.out_bs = (long long)c.outH * c.outW * c.cout;
    } else {
        a.C = c.cin_s; a.KH = a.KW = c.k; a.stride = c.stride; a.pad_t = c.pad_t; a.pad_l = c.pad_l;
        a.OH = c.outH; a.OW = c.outW; a.OC = c.cout_s; a.out_mode = 0;
        a.out_bs = (long long)c.outH * c.outW * c.cout_s;
    }
}

int Engine::transform_forward(const float* params, const float* x3, float* y3_out, cudaStream_t st) {
    FS_CHECK(bound && (flags & ENG_TRANSFORM), "engine has no transform plan / workspace");
    const bool t9 = tc9();
    PROF(PC_POINTWISE, 0.0, reflect_pad_c4(x3, xpad4, N, H, W, 40, st, t9 ? p9[0].hi : nullptr, t9 ? p9[0].lo : nullptr,
                                            (t9 && tc9_x8) ? 1 : 0));
    const float* cur = xpad4;
    for (int l = 0; l < T_NCONV; ++l) {
        const TConv& c = tc[l];
        const bool tcl = use_tc && l >= 3 && l <= 12;
        if (tcl) {
            Conv3x3TcArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.x = tsplit[l]; ta.w = tw_f[l];
            ta.N = N; ta.H = c.inH; ta.W = c.inW; ta.C = 64; ta.OH = c.outH; ta.OW = c.outW; ta.OC = 64; ta.pad = 0;
            ta.out_f32 = tb[l].raw;
            if (in_epi) { ta.stats = in_sums2[l & 1]; ta.stats_c = 64; }
            PROFB(PC_TC_RES_FWD, tc_flops(ta), tc_bytes(ta), launch_conv3x3_tc(ta, st));
        } else if (tc2(l)) {
            Conv3x3TcArgs ta;
            FS_TRY(tc2_args(l, false, -1, tb[l].raw, ta));
            if (in_epi) { ta.stats = in_sums2[l & 1]; ta.stats_c = c.cout; }
            PROFB(PC_TC_S2_FWD, tc2_flops(ta), tc_bytes(ta), launch_conv3x3_tc(ta, st));
        } else if (direct9(c) && tc9()) {
            FS_TRY(tc9_conv(l == 0 ? 0 : 1, cur, tb[l].raw, true, st));
        } else if (direct9(c)) {
            IGemmArgs a;
            conv_fwd_args(c, N, cur, weff[l], tb[l].raw, a);
            PROF(PC_FFMA_CONV, conv9_flops(c, N), launch_conv9x9(cur, weff[l], tb[l].raw, N, c.inH, c.inW, c.cin_s, c.cout_s, st));
        } else {
            IGemmArgs a;
            conv_fwd_args(c, N, cur, weff[l] ? weff[l] : params + c.offW, tb[l].raw, a);
            if (c.upconv && (flags & ENG_DECONV)) a.gather = 1;      // transposed conv: iy = oy - a
            PROF(PC_FFMA_CONV, igemm_flops(a), launch_igemm(a, st));
        }
        // statistics: from the conv epilogue's sums where one accumulated them (finalised inside the fused apply
        // below, or by a separate launch with FS_IN_FUSE=0), else a reduction pass over the raw output
        const bool from_sums = in_epi && (tcl || tc2(l) || (direct9(c) && tc9()));
        const bool fused_apply = from_sums && fuse_in;
        if (fused_apply) {
        } else if (from_sums)
            PROF(PC_IN_STATS, 0.0, instnorm_stats_from_sums(in_sums2[l & 1], tb[l].mean, tb[l].rstd, N, c.outH * c.outW, c.cout_s, IN_EPS, STATS_REPLICAS, st));
        else
            PROF(PC_IN_STATS, 0.0, instnorm_stats(tb[l].raw, tb[l].mean, tb[l].rstd, N, c.outH * c.outW, c.cout_s, IN_EPS, in_partial, st));
        // the other accumulator set is zeroed by every fused apply; a layer off that path does it here
        if (in_epi && fuse_in && !fused_apply) FS_TRY(fill_zero(in_sums2[(l + 1) & 1], (size_t)sums_n * sizeof(double), st));
        const float* skip = nullptr;
        if (l >= 4 && l <= 12 && (l & 1) == 0) skip = tb[l - 2].act;       // residual: block input
        const bool last = l == T_NCONV - 1;
        const float* g = last ? in15 : params + c.offG;
        const float* b = last ? in15 + 4 : params + c.offB;
        float* out = last ? (y3_out ? y3_out : y3) : tb[l].act;
        const bool next_tc = (use_tc && (l + 1) >= 3 && (l + 1) <= 12) || tc2(l + 1);   // next conv consumes split planes
        // training: the input of upsample_2 also as plain split planes for its tensor-path weight gradient
        const bool planes14 = l == 14 && (flags & ENG_TRANSFORM_BWD) && wg9() && a14split.hi;
        __nv_bfloat16* const sp_hi = next_tc ? tsplit[l + 1].hi : (planes14 ? a14split.hi : nullptr);
        __nv_bfloat16* const sp_lo = next_tc ? tsplit[l + 1].lo : (planes14 ? a14split.lo : nullptr);
        // The fp32 copy of the activation is dead when its only reader is a tensor-path conv (split planes): it
        // survives where something else reads it - the residual skip (block inputs 2,4,..,10 and the block
        // outputs' own skip source), the FFMA 9x9 layer (input of 15) and its weight gradient, igemm fallbacks.
        if (!last && next_tc && !keep_acts) {
            const bool skip_src = l >= 2 && l <= 10 && (l & 1) == 0;       // read again as the next block's skip addend
            if (!skip_src) out = nullptr;
        }
        if (fused_apply)
            PROF(PC_IN_APPLY, 0.0, instnorm_apply_from_sums(tb[l].raw, in_sums2[l & 1], STATS_REPLICAS, IN_EPS, tb[l].mean, tb[l].rstd,
                                      in_sums2[(l + 1) & 1], sums_n, g, b, skip, out, N, c.outH, c.outW, c.cout_s, c.act,
                                      last ? 1 : 0, st, sp_hi, sp_lo));
        else
        PROF(PC_IN_APPLY, 0.0, instnorm_apply(tb[l].raw, tb[l].mean, tb[l].rstd, g, b, skip, out, N, c.outH, c.outW,
                                  c.cout_s, c.act, last ? 1 : 0, st, sp_hi, sp_lo));
        cur = tb[l].act;
    }
    return 0;
}

int Engine::transform_backward(const float* params, const float* dY4_in, float* grads, cudaStream_t st) {
    FS_CHECK(bound && (flags & ENG_TRANSFORM_BWD), "engine has no transform backward plan");
    // Three rotating gradient buffers.  cur = buffer holding dAct (gradient w.r.t. the layer's
    // post-activation output; -1 while it is the caller's dY4), held = buffer holding the gradient
    // of the enclosing residual block's output (needed again for the skip path).
    auto pick2 = [](int a, int b) { for (int i = 0; i < 3; ++i) if (i != a && i != b) return i; return -1; };
    const float* dAct = dY4_in;
    int cur = -1, held = -1;
    const float* resid_dOut = nullptr;
    int resid_H = 0, resid_W = 0;
    // table-driven path: gradients computed in a transformed layout are staged per layer and mapped back by
    // finish_weight_grads_table after the loop (2 launches instead of ~12)
    const bool fastp = use_fast_prep();
    for (int l = T_NCONV - 1; l >= 0; --l) {
        const TConv& c = tc[l];
        const bool last = l == T_NCONV - 1;
        const float* g = last ? in15 : params + c.offG;
        const float* b = last ? in15 + 4 : params + c.offB;
        float* dg = last ? gb_tmp : grads + c.offG;
        float* db = last ? gb_tmp + 4 : grads + c.offB;
        const bool second_of_block = (l >= 4 && l <= 12 && (l & 1) == 0);
        if (second_of_block) { resid_dOut = dAct; resid_H = c.outH; resid_W = c.outW; held = cur; }
        const int ri = pick2(cur, held);
        float* dRaw = tgrad[ri];
        const bool tcl = use_tc && l >= 3 && l <= 12;
        const bool tcd = tcl || (l > 0 && tc2(l));           // the data gradient reads dRaw's split planes
        const bool defer_wg = tcl && batch_wgrad;            // residual convs: planes kept in rg[l], weight gradient batched
        const bool wg9_l = direct9(c) && wg9() && !(flags & ENG_DECONV);    // this layer's weight gradient on tcgen05 (x8 form)
        const SplitPtr dsp = defer_wg ? rg[l] : ((wg9_l && l == 0) ? g0split : tgsplit[ri]);
        // dRaw in fp32 is dead when both its readers (data gradient, weight gradient) are tensor-path kernels
        float* dRaw_f32 = dRaw;
        if (!keep_acts && (defer_wg || (tcl && use_tc) || (l > 0 && tc2(l)) || (wg9_l && l == 0))) dRaw_f32 = nullptr;
        const bool tcd_planes = tcd || (wg9_l && l == 0);
        if (fuse_in)
            PROF(PC_IN_BWD, 0.0, instnorm_bwd_sums(dAct, tb[l].raw, tb[l].mean, tb[l].rstd, g, b, dRaw_f32, dg, db, N,
                                    c.outH * c.outW, c.cout_s, c.act, bw_sums2[l & 1], STATS_REPLICAS, bw_sums2[(l + 1) & 1],
                                    sums_n, st, tcd_planes ? dsp.hi : nullptr, tcd_planes ? dsp.lo : nullptr));
        else
        PROF(PC_IN_BWD, 0.0, instnorm_bwd(dAct, tb[l].raw, tb[l].mean, tb[l].rstd, g, b, dRaw_f32, dg, db, N,
                                c.outH * c.outW, c.cout_s, c.act, in_partial, m12, st,
                                tcd_planes ? dsp.hi : nullptr, tcd_planes ? dsp.lo : nullptr));
        if (last && !fastp) {
            FS_CUDA(cudaMemcpyAsync(grads + c.offG, gb_tmp, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
            FS_CUDA(cudaMemcpyAsync(grads + c.offB, gb_tmp + 4, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
        }
        // ---- weight gradient
        const float* in_act = l == 0 ? xpad4 : tb[l - 1].act;
        WGradArgs wa;
        memset(&wa, 0, sizeof(wa));
        wa.in = in_act; wa.dy = dRaw; wa.partial = wg_partial; wa.partial_cap = wg_partial_cap;
        wa.H = c.inH; wa.W = c.inW; wa.in_bs = (long long)c.inH * c.inW * c.cin_s;
        wa.N = N; wa.per_sample = 0; wa.scale = 1.f;
        const bool dcv = (flags & ENG_DECONV) && l >= 13;     // conv2d_transpose layers (im_transf_net.py:57-63)
        bool planes15_ready = false;
        if (wg9_l) {
            const double fl9 = conv9_flops(c, N);
            if (l == 0) {            // windowed planes of the padded input (forward pass) x plain planes of dRaw
                PROF(PC_WGRAD, fl9, launch_wgrad9_x8_tc(p9[0], g0split, grads + c.offW, wg_partial, wg_partial_cap, N, Hp, Wp / 8,
                                                        3, 0, st));
            } else {                 // windowed planes of dRaw (also the data gradient's input) x plain planes of the layer input
                PROF(PC_POINTWISE, 0.0, split_pad_x16(dRaw, p9[2].hi, p9[2].lo, (long long)N * OH, OW, 4, st, 1));
                planes15_ready = true;
                PROF(PC_WGRAD, fl9, launch_wgrad9_x8_tc(p9[2], a14split, grads + c.offW, wg_partial, wg_partial_cap, N, OH, OW / 8,
                                                        3, 1, st));
            }
        } else
        if (dcv && c.upconv) {
            // y = conv2d_transpose(x, W[3,3,cout,cin], s2 SAME) is the data gradient of the SAME conv
            // C: [2H,2W,cout] -> [H,W,cin] with HWIO filter W.  Hence dW = weight gradient of C with
            // dRaw in the role of C's input and x in the role of C's output gradient.
            wa.in = dRaw; wa.H = c.outH; wa.W = c.outW; wa.C = c.cout; wa.in_bs = (long long)c.outH * c.outW * c.cout;
            wa.KH = wa.KW = 3; wa.stride = 2; wa.pad_t = wa.pad_l = 0;     // even input: TF SAME pads 0 / 1
            wa.dy = in_act; wa.OH = c.inH; wa.OW = c.inW; wa.OC = c.cin; wa.dy_mode = 0;
            wa.dy_bs = (long long)c.inH * c.inW * c.cin;
            wa.out = grads + c.offW;
            PROF(PC_WGRAD, wgrad_flops(wa), launch_wgrad(wa, st));
        } else if (dcv) {
            // 9x9 stride-1 conv2d_transpose 16 -> 3(4): dW[k, co, ci] = sum_p dRaw[p + k - 4, co] * x[p, ci]
            const double fl = conv9_flops(c, N);
            PROF(PC_WGRAD, fl, launch_wgrad9x9(dRaw, in_act, wg_tmp, wg_partial, wg_partial_cap, N, c.inH, c.inW,
                                               c.cout_s, c.cin_s, st));
            PROF(PC_PREP, 0.0, unpad_taps(wg_tmp, grads + c.offW, 81, c.cout, c.cin, c.cout_s, c.cin_s, st));
        } else if (tc2(l)) {
            // weight gradient of the 2x2 form on tcgen05: X = the layer's input planes, dY = dRaw's planes, the
            // 128-channel side through its space-to-depth view; then back through the pairing / collapse adjoints
            const bool paired = (l == 1 || l == 14);
            const int gh = c.upconv ? c.inH : c.outH, gw = (c.upconv ? c.inW : c.outW) / (paired ? 2 : 1);
            PROF(PC_WGRAD, wgrad_flops_tc2(N, gh, gw), launch_wgrad2x2_tc(tsplit[l], c.upconv ? 0 : 1, tgsplit[ri], c.upconv ? 1 : 0,
                                                                          fastp ? wgs[l] : wg_tmp, wg_partial, wg_partial_cap, N, gh, gw, st));
            const float* wsrc = wg_tmp;
            if (fastp) {
                // adjoints run after the loop
            } else if (paired) {
                if (c.upconv) PROF(PC_PREP, 0.0, unpair_taps(wg_tmp, wg_tmp2, c.cin, 4 * c.cout, 0, 1, st));   // dY side: paired s2d view
                else PROF(PC_PREP, 0.0, unpair_taps(wg_tmp, wg_tmp2, 4 * c.cin, c.cout, 1, 0, st));
                wsrc = wg_tmp2;
            }
            if (fastp) {
            } else if (c.upconv) PROF(PC_PREP, 0.0, upconv_collapse_grad(wsrc, grads + c.offW, c.cin, c.cout, st));
            else PROF(PC_PREP, 0.0, s2_fwd_collapse_grad(wsrc, grads + c.offW, c.cin, c.cout, st));
        } else if (c.upconv) {
            wa.C = c.cin; wa.KH = wa.KW = 2; wa.stride = 1;
            wa.OH = c.inH; wa.OW = c.inW; wa.OC = 4 * c.cout; wa.dy_mode = 1;
            wa.dy_bs = (long long)c.outH * c.outW * c.cout;
            wa.out = wg_tmp;
            PROF(PC_WGRAD, wgrad_flops(wa), launch_wgrad(wa, st));
            PROF(PC_PREP, 0.0, upconv_collapse_grad(wg_tmp, grads + c.offW, c.cin, c.cout, st));
        } else {
            wa.C = c.cin_s; wa.KH = wa.KW = c.k; wa.stride = c.stride; wa.pad_t = c.pad_t; wa.pad_l = c.pad_l;
            wa.OH = c.outH; wa.OW = c.outW; wa.OC = c.cout_s; wa.dy_mode = 0;
            wa.dy_bs = (long long)c.outH * c.outW * c.cout_s;
            const bool padded = (c.cin != c.cin_s) || (c.cout != c.cout_s);
            wa.out = padded ? ((fastp && wgs[l]) ? wgs[l] : wg_tmp) : grads + c.offW;
            if (defer_wg) {
                // after the sweep (launch_wgrad3x3_tc_multi)
            } else if (tcl)
                PROF(PC_WGRAD, wgrad_flops(wa), launch_wgrad3x3_tc(tsplit[l], tgsplit[ri], wa.out, wg_partial, wg_partial_cap,
                                                                   N, c.inH, c.inW, c.outH, c.outW, 0, st));
            else if (c.k == 9 && c.stride == 1 && c.same && c.cin_s * c.cout_s == 64)
                PROF(PC_WGRAD, conv9_flops(c, N), launch_wgrad9x9(in_act, dRaw, wa.out, wg_partial, wg_partial_cap, N,
                                                                c.inH, c.inW, c.cin_s, c.cout_s, st));
            else
                PROF(PC_WGRAD, wgrad_flops(wa), launch_wgrad(wa, st));
            if (padded && !(fastp && wgs[l])) PROF(PC_PREP, 0.0, unpad_taps(wg_tmp, grads + c.offW, c.k * c.k, c.cin, c.cout, c.cin_s, c.cout_s, st));
        }
        if (l == 0) break;                       // no gradient w.r.t. the input image (train.py:198-204)
        // ---- data gradient -> gradient w.r.t. the previous layer's activation
        const int pidx = pick2(ri, held);
        float* dPrev = tgrad[pidx];
        const bool first_of_block = (l >= 3 && l <= 11 && (l & 1) == 1);
        if (tcl) {
            Conv3x3TcArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.x = dsp; ta.w = tw_d[l];
            ta.N = N; ta.H = c.outH; ta.W = c.outW; ta.C = 64; ta.OH = c.inH; ta.OW = c.inW; ta.OC = 64;
            ta.pad = 2;                              // VALID conv: data gradient is the "full" correlation
            if (first_of_block) { ta.addend = resid_dOut; ta.add_crop = 2; ta.addH = resid_H; ta.addW = resid_W; }
            ta.out_f32 = dPrev;
            PROFB(PC_TC_RES_DGRAD, tc_flops(ta), tc_bytes(ta), launch_conv3x3_tc(ta, st));
            dAct = dPrev; cur = pidx;
            if (first_of_block) { held = -1; resid_dOut = nullptr; }
            continue;
        }
        if (tc2(l)) {
            Conv3x3TcArgs ta;
            FS_TRY(tc2_args(l, true, ri, dPrev, ta));
            PROFB(PC_TC_S2_DGRAD, tc2_flops(ta), tc_bytes(ta), launch_conv3x3_tc(ta, st));
            dAct = dPrev; cur = pidx;
            continue;
        }
        if (dcv) {                               // data gradient of conv2d_transpose = the SAME conv itself
            if (c.upconv) {
                IGemmArgs a;
                memset(&a, 0, sizeof(a));
                a.in = dRaw; a.w = params + c.offW; a.out = dPrev; a.N = N;
                a.H = c.outH; a.W = c.outW; a.C = c.cout; a.in_bs = (long long)c.outH * c.outW * c.cout;
                a.KH = a.KW = 3; a.stride = 2; a.pad_t = a.pad_l = 0;
                a.OH = c.inH; a.OW = c.inW; a.OC = c.cin; a.out_bs = (long long)c.inH * c.inW * c.cin;
                PROF(PC_FFMA_CONV, igemm_flops(a), launch_igemm(a, st));
            } else {                             // wtmp15 = W padded to [81, 4, 16]
                const double fl = conv9_flops(c, N);
                PROF(PC_FFMA_CONV, fl, launch_conv9x9(dRaw, wtmp15, dPrev, N, c.inH, c.inW, c.cout_s, c.cin_s, st));
            }
            dAct = dPrev; cur = pidx;
            continue;
        }
        if (direct9(c) && tc9()) {               // upsample_2: 9x9 data gradient on the tensor path
            FS_TRY(tc9_conv(2, dRaw, dPrev, false, st, planes15_ready));
            dAct = dPrev; cur = pidx;
            continue;
        }
        if (direct9(c)) {                        // data gradient = forward 9x9 conv with flipped weights
            const double fl = conv9_flops(c, N);
            PROF(PC_FFMA_CONV, fl, launch_conv9x9(dRaw, wefft[l], dPrev, N, c.inH, c.inW, c.cout_s, c.cin_s, st));
            dAct = dPrev; cur = pidx;
            continue;
        }
        IGemmArgs a;
        memset(&a, 0, sizeof(a));
        a.in = dRaw; a.w = wefft[l]; a.out = dPrev; a.N = N; a.gather = 1;
        if (c.upconv) {
            a.H = c.inH; a.W = c.inW; a.C = 4 * c.cout; a.in_mode = 1;
            a.in_bs = (long long)c.outH * c.outW * c.cout;
            a.KH = a.KW = 2; a.stride = 1; a.pad_t = a.pad_l = 0;
        } else if (s2_collapsed(c)) {
            a.H = c.outH; a.W = c.outW; a.C = c.cout_s; a.in_mode = 0;
            a.in_bs = (long long)c.outH * c.outW * c.cout_s;
            a.KH = a.KW = 2; a.stride = 1; a.pad_t = a.pad_l = 0;
        } else {
            a.H = c.outH; a.W = c.outW; a.C = c.cout_s; a.in_mode = 0;
            a.in_bs = (long long)c.outH * c.outW * c.cout_s;
            a.KH = a.KW = c.k; a.stride = c.stride; a.pad_t = c.pad_t; a.pad_l = c.pad_l;
        }
        a.OH = c.inH; a.OW = c.inW; a.OC = c.cin_s;
        a.out_bs = (long long)c.inH * c.inW * c.cin_s;
        if (!c.upconv && s2_collapsed(c)) {          // 4 phases as channels, depth-to-space store
            a.OH = c.inH / 2; a.OW = c.inW / 2; a.OC = 4 * c.cin_s; a.out_mode = 1;
        }
        if (first_of_block) {                    // add the skip-path gradient, zero-padded by 2 px
            a.addend = resid_dOut; a.add_crop = 2; a.addH = resid_H; a.addW = resid_W;
            a.add_bs = (long long)resid_H * resid_W * 64;
        }
        PROF(PC_FFMA_CONV, igemm_flops(a), launch_igemm(a, st));
        dAct = dPrev; cur = pidx;
        if (first_of_block) { held = -1; resid_dOut = nullptr; }
    }
    if (use_tc && batch_wgrad) {
        WgradMultiItem items[WGRAD_MULTI_MAX];
        double fl = 0.0;
        for (int l = 3; l <= 12; ++l) {
            WgradMultiItem& it = items[l - 3];
            it.x = tsplit[l]; it.dy = rg[l]; it.out = grads + tc[l].offW;
            it.H = tc[l].inH; it.W = tc[l].inW; it.OH = tc[l].outH; it.OW = tc[l].outW; it.pad = 0;
            fl += 2.0 * N * tc[l].outH * tc[l].outW * 9.0 * 64 * 64;
        }
        PROF(PC_WGRAD, fl, launch_wgrad3x3_tc_multi(items, 10, wg_partial, wg_partial_cap, N, st));
    }
    if (fastp) FS_TRY(finish_weight_grads_table(grads, st));
    return 0;
}

// ---------------------------------------------------------------- VGG
static void vgg_conv_args(const VConv& v, int N, const float* packed, const float* in, float* out, IGemmArgs& a) {
    memset(&a, 0, sizeof(a));
    a.in = in; a.w = packed + v.offW; a.out = out; a.N = N;
    a.H = v.H; a.W = v.W; a.C = v.cin_s; a.in_bs = (long long)v.H * v.W * v.cin_s;
    a.KH = a.KW = 3; a.stride = 1; a.pad_t = a.pad_l = 1;
    a.OH = v.H; a.OW = v.W; a.OC = v.cout; a.out_bs = (long long)v.H * v.W * v.cout;
    a.bias = packed + v.offB; a.relu = 1;
}

int Engine::vgg_forward(const float* packed, const float* img3, int upto, float* const* act_override,
                        cudaStream_t st, unsigned code_mask) {
    FS_CHECK(bound && (flags & ENG_VGG), "engine has no VGG plan / workspace");
    FS_CHECK(upto >= 0 && upto < V_NCONV, "vgg_forward: bad layer %d", upto);
    PROF(PC_POINTWISE, 0.0, vgg_preprocess_c4(img3, v_in4, (long long)N * VH * VW, st));
    const float* cur = v_in4;
    for (int l = 0; l <= upto; ++l) {
        float* out = (act_override && act_override[l]) ? act_override[l] : vact[l];
        const bool pool_next = vc[l].pool_after && l < upto;
        bool pooled_in_epilogue = false;
        if (use_tc && l >= 1) {
            Conv3x3TcArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.x = vsplit[l]; ta.w = vgg_tc_w(packed, vc[l], 0);
            ta.N = N; ta.H = vc[l].H; ta.W = vc[l].W; ta.C = vc[l].cin;
            ta.OH = vc[l].H; ta.OW = vc[l].W; ta.OC = vc[l].cout; ta.pad = 1;
            ta.bias = packed + vc[l].offB; ta.relu = 1;
            // the following max-pool runs in this conv's epilogue (pooled split planes for the next conv)
            const bool fpool = pool_next && fuse_pool;
            // content-target pass: only the targets and the last layer (and un-fused pooled layers) need an fp32 copy
            const bool coded = !act_override && ((code_mask >> l) & 1u) && vcode[l] && (!pool_next || fpool);
            const bool need_f32 = (!act_override && !coded) || (pool_next && !fpool) || (act_override && (act_override[l] || l == upto));
            ta.out_f32 = need_f32 ? out : nullptr;
            if (coded) ta.out_code = vcode[l];
            if (l < upto && !pool_next) ta.out_split = vsplit[l + 1];     // next conv reads split planes
            else if (vtsplit[l].hi && !act_override) ta.out_split = vtsplit[l];   // style tap: Gram kernels read them
            if (fpool) ta.pool_split = vsplit[l + 1];
            pooled_in_epilogue = fpool;
            PROFB(PC_TC_VGG_FWD, tc_flops(ta), tc_bytes(ta), launch_conv3x3_tc(ta, st));
        } else if (l == 0) {             // conv1_1 (Cin = 3): direct shared-memory kernel, exact fp32
            const bool sp = use_tc && l < upto && !pool_next;
            // content-target pass: conv1_1's fp32 copy is dead unless it is a target itself (conv1_2 reads the planes)
            const bool coded0 = sp && !act_override && (code_mask & 1u) && vcode[0];
            float* out0 = ((sp && act_override && !act_override[0]) || coded0) ? nullptr : out;
            if (use_tc && conv11_tc)
                PROF(PC_TC_C11_FWD, 2.0 * N * vc[0].H * vc[0].W * 9.0 * 3 * 64,
                     launch_conv1_1_tc(cur, packed + vc[0].offW, packed + vc[0].offB, out0, sp ? vsplit[1].hi : nullptr,
                                       sp ? vsplit[1].lo : nullptr, coded0 ? vcode[0] : nullptr, N, vc[0].H, vc[0].W, st));
            else
            PROF(PC_FFMA_CONV, 2.0 * N * vc[0].H * vc[0].W * 9.0 * 3 * 64,
                 launch_conv3x3_c4_fwd(cur, packed + vc[0].offW, packed + vc[0].offB, out0, sp ? vsplit[1].hi : nullptr,
                                       sp ? vsplit[1].lo : nullptr, N, vc[0].H, vc[0].W, st, coded0 ? vcode[0] : nullptr));
        } else {
            IGemmArgs a;
            vgg_conv_args(vc[l], N, packed, cur, out, a);
            PROF(PC_FFMA_CONV, igemm_flops(a), launch_igemm(a, st));
        }
        cur = out;
        if (pool_next && pooled_in_epilogue) {
            cur = vpool[l];                  // not written on this path (the next conv reads vsplit[l + 1])
        } else if (pool_next) {
            const bool sp = use_tc != 0;
            PROF(PC_POINTWISE, 0.0, maxpool2x2_fwd(cur, vpool[l], N, vc[l].H, vc[l].W, vc[l].cout, st,
                                  sp ? vsplit[l + 1].hi : nullptr, sp ? vsplit[l + 1].lo : nullptr));
            cur = vpool[l];
        }
    }
    return 0;
}

int Engine::vgg_content_targets(const float* packed, const float* img3, const LossConfig& lc, cudaStream_t st) {
    if (lc.n_content == 0) return 0;
    float* ov[V_NCONV] = {nullptr};
    int top = 0;
    for (int i = 0; i < lc.n_content; ++i) {
        int l = lc.content_layer[i];
        FS_CHECK(l >= 0 && l < V_NCONV && ctarget[l], "content layer %d not planned in this engine", l);
        ov[l] = ctarget[l];
        top = std::max(top, l);
    }
    return vgg_forward(packed, img3, top, ov, st);
}

int Engine::vgg_loss_backward(const float* packed, const float* img3, const LossConfig& lc,
                              const float* const* target_grams, bool need_grad, cudaStream_t st) {
    FS_CHECK(bound && (flags & ENG_VGG), "engine has no VGG plan / workspace");
    FS_CHECK(!need_grad || (flags & ENG_VGG_BWD), "engine has no VGG backward plan");
    float sw[V_NCONV] = {0}, cw[V_NCONV] = {0};
    const float* tg[V_NCONV] = {nullptr};
    int top = -1;
    for (int i = 0; i < lc.n_style; ++i) {
        int l = lc.style_layer[i];
        FS_CHECK(l >= 0 && l < V_NCONV && gram[l], "style layer %d not planned in this engine", l);
        FS_CHECK(tg[l] == nullptr, "style layer %d listed twice", l);
        sw[l] = lc.style_w[i]; tg[l] = target_grams[i]; top = std::max(top, l);
    }
    bool has_c[V_NCONV] = {false};
    for (int i = 0; i < lc.n_content; ++i) {
        int l = lc.content_layer[i];
        FS_CHECK(l >= 0 && l < V_NCONV && ctarget[l], "content layer %d not planned in this engine", l);
        FS_CHECK(!has_c[l], "content layer %d listed twice", l);
        cw[l] = lc.content_w[i]; has_c[l] = true; top = std::max(top, l);
    }
    FS_CHECK(top >= 0, "no loss layers configured");
    // Layers whose fp32 activation nobody reads in this pass store ReLU / arg-max code bytes instead: every tensor-path
    // layer (and conv1_1 feeding one) that is no content target; a layer followed by a pool must be a style tap on
    // the tensor path (its pooling gradient is routed inside the Gram-backward epilogue).
    unsigned code_mask = 0;
    if (need_grad && relu_codes && use_tc && !keep_acts && fuse_pool && fold_pool && top >= 1)
        for (int l = 0; l <= top; ++l) {
            const bool pool_follow = vc[l].pool_after && l < top;
            if (has_c[l]) continue;
            if (l == 0 && (tg[0] || pool_follow)) continue;
            if (pool_follow && !tg[l]) continue;
            code_mask |= 1u << l;
        }
    FS_TRY(vgg_forward(packed, img3, top, nullptr, st, code_mask));

    // split-bf16 planes of activation l after vgg_forward(top): the next conv's input planes, or the
    // dedicated planes of a style tap that is followed by a pool / is the top layer
    auto act_planes = [&](int l, int top_) -> SplitPtr {
        SplitPtr none = {nullptr, nullptr};
        if (!use_tc || l < 1) return none;
        if (l < top_ && !vc[l].pool_after) return vsplit[l + 1];
        return vtsplit[l];
    };
    // ---- losses (+ the Gram-space gradient S = coef*(G-T))
    for (int l = 0; l <= top; ++l) {
        const VConv& v = vc[l];
        const double hwc = (double)v.H * v.W * v.cout;
        if (tg[l]) {
            WGradArgs wa;
            memset(&wa, 0, sizeof(wa));
            wa.in = vact[l]; wa.dy = vact[l]; wa.out = gram[l]; wa.partial = wg_partial; wa.partial_cap = wg_partial_cap;
            wa.H = v.H; wa.W = v.W; wa.C = v.cout; wa.in_bs = (long long)v.H * v.W * v.cout;
            wa.KH = wa.KW = 1; wa.stride = 1;
            wa.OH = v.H; wa.OW = v.W; wa.OC = v.cout; wa.dy_bs = wa.in_bs;
            wa.N = N; wa.per_sample = 1; wa.scale = (float)(1.0 / hwc);
            const SplitPtr fp = act_planes(l, top);
            if (fp.hi)
                PROF(PC_GRAM_FWD, wgrad_flops(wa), launch_gram_tc(fp, gram[l], wg_partial, wg_partial_cap, N, v.H * v.W,
                                                                  v.cout, wa.scale, st));
            else
                PROF(PC_GRAM_FWD, wgrad_flops(wa), launch_wgrad(wa, st));
            const double cc = (double)v.cout * v.cout;
            PROF(PC_LOSS, 0.0, style_loss_grad(gram[l], tg[l], gramS[l], N, v.cout * v.cout,
                                   (float)(4.0 * sw[l] / (cc * hwc)), sw[l] / cc, loss_acc + 1, st));
            if (need_grad && fp.hi) PROF(PC_PREP, 0.0, pack_gemm_b_tc(gramS[l], gsS[l], N, v.cout, st));
        }
        if (has_c[l])
            PROF(PC_LOSS, 0.0, sqdiff_sum(vact[l], ctarget[l], (long long)N * v.H * v.W * v.cout, cw[l] / hwc, loss_acc + 0, st));
    }
    if (!need_grad) return 0;

    // ---- backward: P_l = dLoss/d(pre-activation of conv l), top-down
    auto pick = [](int a, int b, int c) { for (int i = 0; i < 4; ++i) if (i != a && i != b && i != c) return i; return -1; };
    // data gradient of VGG conv lsrc applied to the masked gradient held in vgrad[pidx]
    // the ReLU reference of layer l: its code bytes when it stored them, else its fp32 activation
    auto refc = [&](int l) -> const unsigned char* { return ((code_mask >> l) & 1u) ? vcode[l] : nullptr; };
    auto reff = [&](int l) -> const float* { return ((code_mask >> l) & 1u) ? nullptr : vact[l]; };
    auto dgrad = [&](int lsrc, int pidx, float* out, SplitPtr out_split, const float* addend, const float* ref,
                     const unsigned char* ref_code) -> int {
        const VConv& v = vc[lsrc];
        const float* P = vgrad[pidx];
        if (use_tc && lsrc >= 1) {
            Conv3x3TcArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.x = vgsplit[pidx]; ta.w = vgg_tc_w(packed, v, 1);
            ta.N = N; ta.H = v.H; ta.W = v.W; ta.C = v.cout; ta.OH = v.H; ta.OW = v.W; ta.OC = v.cin; ta.pad = 1;
            ta.addend = addend; ta.addH = v.H; ta.addW = v.W; ta.ref = ref; ta.ref_code = ref_code;
            ta.out_f32 = out; ta.out_split = out_split;
            PROFB(PC_TC_VGG_DGRAD, tc_flops(ta), tc_bytes(ta), launch_conv3x3_tc(ta, st));
            return 0;
        }
        if (lsrc == 0 && !addend && !ref) {      // conv1_1 data gradient: direct kernel (64 -> 4 channels)
            PROF(PC_FFMA_CONV, 2.0 * N * v.H * v.W * 9.0 * 3 * 64, launch_dgrad3x3_c4(P, packed + v.offWT, out, N, v.H, v.W, st));
            return 0;
        }
        FS_CHECK(lsrc != 0, "conv1_1 data gradient with an epilogue is not supported");
        FS_CHECK(!ref_code, "code-byte references need the tensor path");
        IGemmArgs a;
        memset(&a, 0, sizeof(a));
        a.in = P; a.w = packed + v.offWT; a.out = out; a.N = N; a.gather = 1;
        a.H = v.H; a.W = v.W; a.C = v.cout; a.in_bs = (long long)v.H * v.W * v.cout;
        a.KH = a.KW = 3; a.stride = 1; a.pad_t = a.pad_l = 1;
        a.OH = v.H; a.OW = v.W; a.OC = v.cin_s; a.out_bs = (long long)v.H * v.W * v.cin_s;
        a.addend = addend; a.addH = v.H; a.addW = v.W; a.add_bs = a.out_bs; a.ref = ref;
        PROF(PC_FFMA_CONV, igemm_flops(a), launch_igemm(a, st));
        return 0;
    };
    const SplitPtr no_split = {nullptr, nullptr};
    // make the split planes of P_l (held in vgrad[idx]) valid when the consumer is a tensor-path conv
    auto ensure_split = [&](int l, int idx) -> int {
        if (!(use_tc && l >= 1)) return 0;
        return split_bf16(vgrad[idx], vgsplit[idx], (long long)N * vc[l].H * vc[l].W * vc[l].cout, st);
    };
    // pool_g / ct / cw2c: gradient w.r.t. the max-pool of this activation, content target and its coefficient -
    // folded into the tensor-path epilogue (no pool_bwd_combine pass); the FFMA path takes them pre-combined in addend
    auto gram_bwd = [&](int l, const float* addend, const float* ref, const unsigned char* ref_code, float* out,
                        SplitPtr out_split, const float* pool_g, const float* ctt, float cw2c) -> int {
        const VConv& v = vc[l];
        const SplitPtr fp = act_planes(l, top);
        if (fp.hi) {                     // dF = F * S as a per-sample 1x1 tensor-path GEMM
            Conv3x3TcArgs ta;
            memset(&ta, 0, sizeof(ta));
            ta.x = fp; ta.w = gsS[l]; ta.one_by_one = 1; ta.per_sample_w = 1;
            ta.N = N; ta.H = v.H; ta.W = v.W; ta.C = v.cout; ta.OH = v.H; ta.OW = v.W; ta.OC = v.cout; ta.pad = 0;
            ta.addend = addend; ta.addH = v.H; ta.addW = v.W; ta.ref = ref; ta.ref_code = ref_code;
            ta.pool_grad = pool_g; ta.ctarget = ctt; ta.cw2 = cw2c;
            ta.out_f32 = out; ta.out_split = out_split;
            PROFB(PC_GRAM_BWD, tc_flops(ta) / 9.0, tc_bytes(ta), launch_conv3x3_tc(ta, st));
            return 0;
        }
        FS_CHECK(!pool_g && !ctt && !ref_code, "gram backward: fused pool / content terms need the tensor path");
        IGemmArgs a;
        memset(&a, 0, sizeof(a));
        a.in = vact[l]; a.w = gramS[l]; a.w_bs = (long long)v.cout * v.cout; a.out = out; a.N = N;
        a.H = v.H; a.W = v.W; a.C = v.cout; a.in_bs = (long long)v.H * v.W * v.cout;
        a.KH = a.KW = 1; a.stride = 1;
        a.OH = v.H; a.OW = v.W; a.OC = v.cout; a.out_bs = a.in_bs;
        a.addend = addend; a.addH = v.H; a.addW = v.W; a.add_bs = a.in_bs; a.ref = ref;
        PROF(PC_GRAM_BWD, igemm_flops(a), launch_igemm(a, st));
        return 0;
    };
    int pi = -1;
    for (int l = top; l >= 0; --l) {
        const VConv& v = vc[l];
        const bool pool_follow = v.pool_after && l < top;
        const float* ct = has_c[l] ? ctarget[l] : nullptr;
        const float cw2 = (float)(2.0 * cw[l] / ((double)v.H * v.W * v.cout));
        if (l == top || pool_follow) {
            int gi = -1; const float* gp = nullptr;
            if (pool_follow) {
                gi = pick(pi, -1, -1);
                FS_TRY(dgrad(l + 1, pi, vgrad[gi], no_split, nullptr, nullptr, nullptr));
                gp = vgrad[gi];
            }
            if (tg[l]) {
                int ti = -1; const float* T = nullptr;
                const bool tcg = act_planes(l, top).hi != nullptr;
                const bool fold = tcg && fold_pool;          // pool routing + content term inside the Gram-backward epilogue
                if ((gp || ct) && !fold) {
                    ti = pick(gi, -1, -1);
                    PROF(PC_POINTWISE, 0.0, pool_bwd_combine(vact[l], gp, ct, cw2, 0, vgrad[ti], N, v.H, v.W, v.cout, st));
                    T = vgrad[ti];
                }
                int oi = pick(gi, ti, -1);
                // P_l feeds the tensor-path data gradient of conv l: only its split planes are needed
                FS_TRY(gram_bwd(l, T, reff(l), refc(l), tcg ? nullptr : vgrad[oi], tcg ? vgsplit[oi] : no_split,
                                fold ? gp : nullptr, fold ? ct : nullptr, cw2));
                if (!tcg) PROF(PC_POINTWISE, 0.0, ensure_split(l, oi));
                pi = oi;
            } else {
                int oi = pick(gi, -1, -1);
                PROF(PC_POINTWISE, 0.0, pool_bwd_combine(vact[l], gp, ct, cw2, 1, vgrad[oi], N, v.H, v.W, v.cout, st));
                PROF(PC_POINTWISE, 0.0, ensure_split(l, oi));
                pi = oi;
            }
        } else {
            int ai = -1; const float* A = nullptr;
            if (tg[l]) {
                int ti = -1; const float* T = nullptr;
                if (ct) {
                    ti = pick(pi, -1, -1);
                    PROF(PC_POINTWISE, 0.0, pool_bwd_combine(vact[l], nullptr, ct, cw2, 0, vgrad[ti], N, v.H, v.W, v.cout, st));
                    T = vgrad[ti];
                }
                ai = pick(pi, ti, -1);
                FS_TRY(gram_bwd(l, T, nullptr, nullptr, vgrad[ai], no_split, nullptr, nullptr, 0.f));
                A = vgrad[ai];
            } else if (ct) {
                ai = pick(pi, -1, -1);
                PROF(PC_POINTWISE, 0.0, pool_bwd_combine(vact[l], nullptr, ct, cw2, 0, vgrad[ai], N, v.H, v.W, v.cout, st));
                A = vgrad[ai];
            }
            int oi = pick(pi, ai, -1);
            FS_TRY(dgrad(l + 1, pi, (use_tc && l >= 1) ? nullptr : vgrad[oi], (use_tc && l >= 1) ? vgsplit[oi] : no_split, A,
                         reff(l), refc(l)));
            pi = oi;
        }
    }
    FS_TRY(dgrad(0, pi, dY4, no_split, nullptr, nullptr, nullptr));
    return 0;
}

int Engine::train_fwd_bwd(const float* params, const float* packed, const float* x3, const LossConfig& lc,
                          const float* const* target_grams, float* grads, float* losses4, float* y3_out,
                          cudaStream_t st) {
    const int need = ENG_TRANSFORM | ENG_TRANSFORM_BWD | ENG_VGG | ENG_VGG_BWD;
    FS_CHECK(bound && (flags & need) == need, "engine was not created for training");
    FS_CHECK(OH == H && OW == W, "train step needs (H+80)%%4==0 and (W+80)%%4==0 so that the stylised "
             "image matches the content targets (got %dx%d -> %dx%d)", H, W, OH, OW);
    float* Y = y3_out ? y3_out : y3;
    FS_TRY(prep_transform_weights(params, true, st));
    weights_prepared = false;                      // the caller's optimiser step changes the parameters
    FS_TRY(fill_zero(loss_acc, 4 * sizeof(double), st));
    FS_TRY(vgg_content_targets(packed, x3, lc, st));                  // train.py:250-251
    FS_TRY(transform_forward(params, x3, Y, st));                     // train.py:161
    FS_TRY(vgg_loss_backward(packed, Y, lc, target_grams, true, st)); // train.py:165-184 + autodiff
    if (lc.beta != 0.f) PROF(PC_LOSS, 0.0, tv_loss_grad(Y, dY4, N, VH, VW, lc.beta, loss_acc + 2, st));
    FS_TRY(transform_backward(params, dY4, grads, st));
    FS_TRY(finalize_losses(loss_acc, losses4, st));
    return 0;
}

}  // namespace fs
