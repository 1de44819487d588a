"""Whole-path parity through the C-ABI composites against the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ckpt as ockpt  # noqa: E402
from oracle import restate as R  # noqa: E402

STYLE = ("conv1_2", "conv2_2", "conv3_3", "conv4_3")


def _relerr(got, want):
    want = torch.as_tensor(want).double()
    return float((torch.as_tensor(got).double().cpu() - want).abs().max() / max(want.abs().max().item(), 1e-30))


@pytest.fixture(scope="module")
def starry(golden_dir):
    return ockpt.load(os.path.join(golden_dir, "starry_final.ckpt"))


def test_transform_forward_matches_oracle(built_lib, starry):
    """create_net forward (im_transf_net.py:14-75), B=2 96x80 uniform-noise input.
    Tolerance: max-abs <= 1e-3 on the [0,1]-normalised pixel scale (0.255 levels on 0..255),
    the SURVEY 7.3 statement of north_star's pixel tolerance; fp32 FFMA lands ~1e-5."""
    from faststyle_b200.engine import Engine, params_to_device
    rng = np.random.RandomState(0)
    x = rng.randint(0, 256, (2, 96, 80, 3)).astype(np.float32)
    with torch.no_grad():
        yo = R.create_net(x, starry, "resize", torch.float64)
    eng = Engine(2, 96, 80, transform=True)
    y = eng.transform_forward(params_to_device(starry, "cuda"), x)
    torch.cuda.synchronize()
    assert tuple(y.shape) == tuple(yo.shape)
    err = float((y.double().cpu() - yo).abs().max()) / 255.0
    assert err <= 1e-3, err
    assert err <= 2e-4, "split-bf16/fp32 path should be far inside the tolerance (%g)" % err
    eng.set_tensor_path(False)                      # exact-fp32 FFMA path
    y32 = eng.transform_forward(params_to_device(starry, "cuda"), x)
    torch.cuda.synchronize()
    assert float((y32.double().cpu() - yo).abs().max()) / 255.0 <= 5e-5


@pytest.mark.parametrize("shape", [(1, 64, 72), (2, 52, 100), (1, 50, 61), (3, 41, 44)])
def test_transform_intermediates(built_lib, starry, shape):
    """Per-layer activations of create_net on the tensor path.  (64,72) / (52,100): initconv_2 sees an even
    input and runs in its 2x2 space-to-depth form on tcgen05 (ragged 33x45 tiles for the second); (50,61) and
    (41,44): odd inputs keep the generic stride-2 gather (TF SAME pads 1/1 there); upsample_0 is the 4-phase
    tensor-path resize-conv in every case."""
    from faststyle_b200.engine import Engine, params_to_device
    N, H, W = shape
    rng = np.random.RandomState(1 + H)
    x = rng.randint(0, 256, (N, H, W, 3)).astype(np.float32)
    taps = {}
    with torch.no_grad():
        R.create_net(x, starry, "resize", torch.float64, taps=taps)
    eng = Engine(N, H, W, transform=True)
    eng.keep_activations(True)                  # fp32 copies of the activations only tensor-path convs read
    eng.transform_forward(params_to_device(starry, "cuda"), x)
    torch.cuda.synchronize()
    names = {0: "initconv_0", 1: "initconv_1", 2: "initconv_2", 4: "resblock_0", 12: "resblock_4",
             13: "upsample_0", 14: "upsample_1"}
    for idx, nm in names.items():
        got = eng.transform_activation(idx, 1)
        assert _relerr(got, taps[nm]) < 2e-4, (nm, _relerr(got, taps[nm]))
    raw15 = eng.transform_activation(15, 0)[..., :3]
    assert _relerr(raw15, taps["upsample_2/conv"]) < 2e-4


@pytest.mark.parametrize("tensor_path", [False, True])
def test_golden_chicago(built_lib, starry, golden_dir, tensor_path):
    """Reference golden pair: results/chicago.jpg -> results/starry_chicago.jpg
    (README.md:5-18,59-61) through the CUDA path, JPEG q95 like cv2.imwrite.
    The exact-fp32 path must match as well as the oracle does (>= 98.5 % identical decoded
    sub-pixels); the split-bf16 tensor path perturbs pre-rounding pixels by ~1e-2 levels, which
    flips ~1 % of uint8 roundings before the JPEG encoder, so its floor is 97 %."""
    import cv2
    from faststyle_b200.engine import Engine, params_to_device
    img = cv2.cvtColor(cv2.imread(os.path.join(golden_dir, "chicago.jpg")), cv2.COLOR_BGR2RGB)
    eng = Engine(1, img.shape[0], img.shape[1], transform=True)
    eng.set_tensor_path(tensor_path)
    y = eng.transform_forward(params_to_device(starry, "cuda"), img[None]).cpu().numpy()[0]
    assert y.shape == (476, 712, 3)
    bgr = cv2.cvtColor(np.clip(np.rint(y), 0, 255).astype(np.uint8), cv2.COLOR_RGB2BGR)
    ok, enc = cv2.imencode(".jpg", bgr)
    dec = cv2.imdecode(enc, cv2.IMREAD_COLOR).astype(int)
    gold = cv2.imread(os.path.join(golden_dir, "starry_chicago.jpg")).astype(int)
    d = np.abs(dec - gold)
    print("golden chicago tensor_path=%s: identical %.4f mae %.4f max %d" % (tensor_path, (d == 0).mean(), d.mean(), d.max()))
    floor = 0.97 if tensor_path else 0.985
    assert (d == 0).mean() >= floor and d.mean() <= 0.05 and d.max() <= 10


def _setup_loss(N, H, W, seed=0):
    from faststyle_b200 import synth
    vggw = synth.synthetic_vgg_weights(7)
    rng = np.random.RandomState(seed)
    style = rng.randint(0, 256, (1, 40, 56, 3)).astype(np.float32)
    tg = R.style_target_grams(style, vggw, STYLE, torch.float64)
    return vggw, style, tg, rng


def _cos(a, b):
    a, b = a.double().flatten().cpu(), torch.as_tensor(b).double().flatten()
    return float((a @ b) / (a.norm() * b.norm()))


# tolerances per path: (loss scalars relative, gradient max-abs relative to max, 1 - cosine).  The per-tensor
# gradient bound of the tensor path is 3x the worst value measured with the trained checkpoint over these cases
# (5.2e-3; the exact-fp32 path measures 9e-4); the untrained deconv case has its own bound below.
TOL = {False: (2e-4, 2e-3, 1e-7), True: (1e-3, 1.5e-2, 1e-5)}


@pytest.mark.parametrize("tensor_path", [False, True])
def test_perceptual_loss_and_pixel_gradient(built_lib, tensor_path):
    """slow_style.py:116-154 loss + dLoss/dX vs oracle autograd (fp64).  Parity is
    UNPINNED by the reference (no goldens for the loss path) - oracle self-consistency only.
    Loss scalars: <= 1e-3 relative on the tensor path (north_star's loss tolerance), 2e-4 on the
    exact-fp32 path.  Gradients are ill-conditioned sums: fp32-vs-fp64 of the oracle itself
    differs by 1e-4 relative-to-max, the 16-bit-mantissa split operands by ~100x that."""
    from faststyle_b200.engine import Engine, make_loss_config, pack_vgg
    N, H, W = 1, 48, 40
    vggw, style, tg, rng = _setup_loss(N, H, W)
    content = rng.randint(0, 256, (N, H, W, 3)).astype(np.float32)
    xvar = rng.uniform(0, 255, (N, H, W, 3)).astype(np.float32)
    with torch.no_grad():
        ct = [R.vgg16_layers(content, vggw, "conv3_3", torch.float64)["conv3_3"]]
    ref = R.slow_style_grads(xvar, ct, vggw, tg, beta=1e-4, dtype=torch.float64)

    packed = pack_vgg(vggw, "cuda")
    cfg = make_loss_config(["conv3_3"], [1.0], STYLE, [5.0] * 4, 1e-4)
    ltol, gtol, ctol = TOL[tensor_path]
    eng = Engine(N, H, W, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE)
    seng = Engine(1, 40, 56, vgg=True, style_layers=STYLE)
    eng.set_tensor_path(tensor_path); seng.set_tensor_path(tensor_path)
    grams = seng.vgg_grams(packed, style, STYLE)
    for g, t in zip(grams, tg):
        assert _relerr(g, t) < (1e-3 if tensor_path else 1e-5)
    eng.set_content_targets(packed, content, cfg)
    losses, grad = eng.perceptual_loss(packed, xvar, cfg, grams)
    torch.cuda.synchronize()
    L = losses.cpu().double()
    errs = [abs(float(got) - float(want)) / max(abs(float(want)), 1e-30)
            for got, want in zip(L, [ref["content"], ref["style"], ref["tv"], ref["loss"]])]
    print("perceptual tensor_path=%s: loss rel errs %s, grad relmax %.3g, 1-cos %.3g" %
          (tensor_path, ["%.2g" % e for e in errs], _relerr(grad, ref["grad"]), 1 - _cos(grad, ref["grad"])))
    assert max(errs) <= ltol
    assert _relerr(grad, ref["grad"]) < gtol
    assert 1 - _cos(grad, ref["grad"]) < ctol


@pytest.mark.parametrize("tensor_path", [False, True])
def test_train_step_gradients(built_lib, starry, tensor_path):
    """train.py:158-204 step: losses + all 48 variable gradients vs oracle autograd (fp64)."""
    from faststyle_b200.engine import Engine, make_loss_config, pack_vgg, params_to_device
    from faststyle_b200.layout import transform_offsets
    N, H, W = 2, 64, 48
    vggw, style, tg, rng = _setup_loss(N, H, W, seed=3)
    x = rng.randint(0, 256, (N, H, W, 3)).astype(np.float32)
    ref = R.train_grads(x, starry, vggw, tg, dtype=torch.float64, beta=1e-4)

    packed = pack_vgg(vggw, "cuda")
    cfg = make_loss_config(["conv3_3"], [1.0], STYLE, [5.0] * 4, 1e-4)
    ltol, gtol, ctol = TOL[tensor_path]
    eng = Engine(N, H, W, transform_bwd=True, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE)
    eng.set_tensor_path(tensor_path)
    tgd = [t.float().cuda().contiguous() for t in tg]
    y = torch.empty((N, H, W, 3), device="cuda")
    grads, losses = eng.train_fwd_bwd(params_to_device(starry, "cuda"), packed, x, cfg, tgd, y=y)
    torch.cuda.synchronize()
    assert float((y.double().cpu() - ref["Y"]).abs().max()) / 255.0 < 2e-4
    L = losses.cpu().double()
    for got, want in zip(L, [ref["content"], ref["style"], ref["tv"], ref["loss"]]):
        assert abs(float(got) - float(want)) <= ltol * abs(float(want)) + 1e-12, (float(got), float(want))
    g = grads.cpu().double()
    flat_ref = torch.cat([ref["grads"][n].flatten() for n in transform_offsets()])
    print("train step tensor_path=%s: 1-cos(grad) %.3g, rel L2 %.3g" %
          (tensor_path, 1 - _cos(g, flat_ref), float((g - flat_ref).norm() / flat_ref.norm())))
    assert 1 - _cos(g, flat_ref) < ctol
    worst = 0.0
    for name, (off, shape) in transform_offsets().items():
        want = ref["grads"][name[len("img_t_net/"):]] if name not in ref["grads"] else ref["grads"][name]
        got = g[off:off + want.numel()].view(want.shape)
        e = float((got - want).abs().max() / max(want.abs().max().item(), 1e-30))
        worst = max(worst, e)
        assert e < gtol, (name, e)
    print("worst relative gradient error", worst)


def _deconv_params(rng):
    from faststyle_b200.layout import TRANSFORM_VARS_DECONV
    params = {}
    for name, shape in TRANSFORM_VARS_DECONV:
        leaf = name.rsplit("/", 1)[1]
        if leaf.startswith("INscale"):
            params[name] = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif leaf.startswith("INshift"):
            params[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        else:
            params[name] = (rng.standard_normal(shape) * (0.3 if "upsample" in name else 0.1)).astype(np.float32)
    return params


@pytest.mark.parametrize("tensor_path", [False, True])
def test_train_step_gradients_deconv(built_lib, tensor_path):
    """training the 'deconv' upsampling variant (im_transf_net.py:57-63,158-190;
    train.py --upsample_method deconv).  conv2d_transpose backward = the SAME conv (data gradient) and its
    weight gradient with the roles of input / output-gradient swapped; checked against oracle autograd
    (fp64).  Unpinned like the rest of the training path: the reference ships no deconv checkpoint."""
    from faststyle_b200.engine import Engine, make_loss_config, pack_vgg, params_to_device
    from faststyle_b200.layout import transform_offsets
    N, H, W = 2, 48, 64
    vggw, style, tg, rng = _setup_loss(N, H, W, seed=11)
    params = _deconv_params(rng)
    x = rng.randint(0, 256, (N, H, W, 3)).astype(np.float32)
    ref = R.train_grads(x, params, vggw, tg, dtype=torch.float64, beta=1e-4, upsample_method="deconv")

    packed = pack_vgg(vggw, "cuda")
    cfg = make_loss_config(["conv3_3"], [1.0], STYLE, [5.0] * 4, 1e-4)
    ltol, gtol, ctol = TOL[tensor_path]
    eng = Engine(N, H, W, transform_bwd=True, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE,
                 deconv=True)
    eng.set_tensor_path(tensor_path)
    tgd = [t.float().cuda().contiguous() for t in tg]
    y = torch.empty((N, H, W, 3), device="cuda")
    grads, losses = eng.train_fwd_bwd(params_to_device(params, "cuda", "deconv"), packed, x, cfg, tgd, y=y)
    torch.cuda.synchronize()
    assert float((y.double().cpu() - ref["Y"]).abs().max()) / 255.0 < 2e-4
    L = losses.cpu().double()
    for got, want in zip(L, [ref["content"], ref["style"], ref["tv"], ref["loss"]]):
        assert abs(float(got) - float(want)) <= ltol * abs(float(want)) + 1e-12, (float(got), float(want))
    g = grads.cpu().double()
    offs = transform_offsets("deconv")
    flat_ref = torch.cat([ref["grads"][n].flatten() for n in offs])
    print("deconv train step tensor_path=%s: 1-cos(grad) %.3g" % (tensor_path, 1 - _cos(g, flat_ref)))
    # random (untrained) weights with the N(0, 0.3^2) upsample filters of this test drive activations ~10x larger
    # than a trained net's; on the split-bf16 path the direction error of the 424102-vector sits at ~1e-5 here
    # (2.8e-8 with the trained starry weights above), so this case gets 3e-5 instead of 1e-5
    assert 1 - _cos(g, flat_ref) < (3e-5 if tensor_path else ctol)
    # same reason for the per-tensor bound: 3.7e-2 measured on the tensor path for the smallest InstanceNorm-shift
    # gradients of this untrained net (the trained checkpoint sits at 1.7e-3, see TOL)
    worst = 0.0
    for name, (off, shape) in offs.items():
        want = ref["grads"][name]
        got = g[off:off + want.numel()].view(want.shape)
        e = float((got - want).abs().max() / max(want.abs().max().item(), 1e-30))
        worst = max(worst, e)
        assert e < (6e-2 if tensor_path else gtol), (name, e)
    print("deconv worst relative gradient error", worst)


def test_train_step_batch_additivity_full_size(built_lib, starry):
    """BASELINE configs 3/4 at full size (per-GPU batch 8, 256x256) through a size-independent property: the
    reference's losses SUM over the batch (losses.py:32-37,63-64) and InstanceNorm is per sample, so
    grads(batch of 8) == grads(images 0-3) + grads(images 4-7) and likewise for the loss scalars - the identity
    the data-parallel all-reduce(SUM) relies on.  fp32 summation order is the only difference."""
    from faststyle_b200.engine import Engine, make_loss_config, pack_vgg, params_to_device
    N, H, W = 8, 256, 256
    vggw, style, tg, rng = _setup_loss(N, H, W, seed=21)
    x = rng.randint(0, 256, (N, H, W, 3)).astype(np.float32)
    packed = pack_vgg(vggw, "cuda")
    cfg = make_loss_config(["conv3_3"], [1.0], STYLE, [5.0] * 4, 1e-4)
    tgd = [t.float().cuda().contiguous() for t in tg]
    p = params_to_device(starry, "cuda")
    e8 = Engine(N, H, W, transform_bwd=True, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE)
    g8, l8 = e8.train_fwd_bwd(p, packed, x, cfg, tgd)
    g8, l8 = g8.double().cpu(), l8.double().cpu()
    del e8
    e4 = Engine(4, H, W, transform_bwd=True, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE)
    ga, la = e4.train_fwd_bwd(p, packed, x[:4], cfg, tgd)
    ga, la = ga.double().cpu(), la.double().cpu()
    gb, lb = e4.train_fwd_bwd(p, packed, x[4:], cfg, tgd)
    gb, lb = gb.double().cpu(), lb.double().cpu()
    torch.cuda.synchronize()
    assert torch.isfinite(g8).all() and float(g8.abs().max()) > 0
    rel = float((g8 - (ga + gb)).abs().max() / g8.abs().max())
    print("batch additivity at 8x256x256: grad rel-to-max %.3g, losses %s vs %s" % (rel, l8.tolist(), (la + lb).tolist()))
    assert rel < 1e-4
    # the fused InstanceNorm statistics are fp32 partial sums whose grouping follows the tile -> CTA assignment, which
    # differs between a batch-8 and a batch-4 plan: the two sides differ by fp32 summation order in every layer
    # (each side agrees with the fp64 oracle to 1 - cos = 6e-9 at this size, test_gpu_fullsize.py)
    assert 1 - _cos(g8, ga + gb) < 2e-8
    for a, b in zip(l8, la + lb):
        assert abs(float(a) - float(b)) <= 1e-5 * abs(float(b)) + 1e-12


def test_slow_style_step_1024(built_lib):
    """BASELINE config 5 at full size: slow_style.py:116-154 loss + pixel gradient of a 1024x1024 variable image
    (VGG16 + Gram path only) against the oracle's autograd in fp32 on the host cores (~20 s), plus the Gram
    symmetry property on the largest taps.  Tolerances as test_perceptual_loss_and_pixel_gradient (tensor path)."""
    from faststyle_b200.engine import Engine, make_loss_config, pack_vgg
    N, H, W = 1, 1024, 1024
    vggw, style, tg, _ = _setup_loss(N, H, W)
    g = torch.Generator().manual_seed(0)
    content = torch.randint(0, 256, (N, H, W, 3), generator=g).float().numpy()
    g2 = torch.Generator().manual_seed(2)
    xvar = (torch.rand((N, H, W, 3), generator=g2) * 255.0).numpy()
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    with torch.no_grad():
        ct = [R.vgg16_layers(content, vggw, "conv3_3", torch.float32)["conv3_3"]]
    ref = R.slow_style_grads(xvar, ct, vggw, tg, beta=1e-4, dtype=torch.float32)

    packed = pack_vgg(vggw, "cuda")
    cfg = make_loss_config(["conv3_3"], [1.0], STYLE, [5.0] * 4, 1e-4)
    eng = Engine(N, H, W, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE)
    eng.set_content_targets(packed, content, cfg)
    tgd = [t.float().cuda().contiguous() for t in tg]
    losses, grad = eng.perceptual_loss(packed, xvar, cfg, tgd)
    torch.cuda.synchronize()
    L = losses.cpu().double()
    errs = [abs(float(got) - float(want)) / max(abs(float(want)), 1e-30)
            for got, want in zip(L, [ref["content"], ref["style"], ref["tv"], ref["loss"]])]
    print("slow_style 1024^2: loss rel errs %s, grad relmax %.3g, 1-cos %.3g" %
          (["%.2g" % e for e in errs], _relerr(grad, ref["grad"]), 1 - _cos(grad, ref["grad"])))
    assert max(errs) <= 1e-3
    assert 1 - _cos(grad, ref["grad"]) < 1e-5
    grams = eng.vgg_grams(packed, xvar, STYLE)
    torch.cuda.synchronize()
    for G in grams:
        Gc = G.double().cpu()
        assert float((Gc - Gc.transpose(1, 2)).abs().max() / Gc.abs().max()) < 2e-5
