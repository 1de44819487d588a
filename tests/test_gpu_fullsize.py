"""Oracle parity at BASELINE.json's own sizes, through the C-ABI composites.

config 2: transform-net inference, batch 32, 256x256 - seed-0 uniform noise AND chicago area-resized to 256x256
          tiled x32 (natural-image statistics: flat sky regions stress InstanceNorm's variance), starry ckpt,
          vs the fp64 oracle `create_net` (im_transf_net.py:14-75).
config 3: train.py step, batch 4, 256x256, seed-0 input, style starry_night_crop.jpg (640x938): losses and all 48
          gradients vs oracle autograd (train.py:158-204), with the starry ckpt AND the TF initialisers from seed 1.
These are the sizes at which the 16-row tiles, multi-wave persistent loops and large split-K reductions of the
tensor-path kernels actually run (the small cases in test_gpu_paths.py mostly take the 8-row single-wave paths).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ckpt as ockpt  # noqa: E402
from oracle import restate as R  # noqa: E402

STYLE = ("conv1_2", "conv2_2", "conv3_3", "conv4_3")

# Tolerances at full size on the tensor path (split-bf16 x3, fp32 accumulate), vs the fp64 oracle.
PIX_TOL = 2e-4            # stylised pixels, max-abs on the [0,1] scale (north_star bound: 1e-3)
LOSS_TOL = 1e-3           # loss scalars, relative (north_star's loss tolerance); measured ~1e-4
GRAD_TOL = 1e-2           # per-tensor gradient max-abs relative to the tensor's max (VERDICT r1 #1a)
COS_TOL = 1e-6            # 1 - cosine of the whole 424102-vector


def _threads():
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))


@pytest.fixture(scope="module")
def starry(golden_dir):
    return ockpt.load(os.path.join(golden_dir, "starry_final.ckpt"))


def _noise(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, 256, 256, 3), generator=g).float().numpy()      # SURVEY 8(d) config 2/3 input


def _chicago256(golden_dir):
    import cv2
    img = cv2.cvtColor(cv2.imread(os.path.join(golden_dir, "chicago.jpg")), cv2.COLOR_BGR2RGB)
    return cv2.resize(img, (256, 256), interpolation=cv2.INTER_AREA).astype(np.float32)


@pytest.mark.parametrize("kind", ["noise", "chicago"])
def test_config2_transform_b32(built_lib, starry, golden_dir, kind):
    """BASELINE config 2 at full size: batch 32, 256x256, pixel tolerance vs the fp64 oracle."""
    from faststyle_b200.engine import Engine, params_to_device
    _threads()
    if kind == "noise":
        x = _noise(32)
    else:
        # 32 tiles of the same natural image, each with its own small brightness offset so that a per-sample
        # indexing slip cannot hide behind identical samples
        base = _chicago256(golden_dir)
        x = np.stack([np.clip(base + (i % 4), 0, 255) for i in range(32)]).astype(np.float32)
    eng = Engine(32, 256, 256, transform=True)
    p = params_to_device(starry, "cuda")
    y = eng.transform_forward(p, x)
    torch.cuda.synchronize()
    assert tuple(y.shape) == (32, 256, 256, 3)
    yc = y.double().cpu()
    worst = 0.0
    with torch.no_grad():
        if kind == "chicago":                  # four distinct samples, each repeated 8 times across the batch
            yo4 = R.create_net(x[:4], starry, "resize", torch.float64)
            for i in range(32):
                worst = max(worst, float((yc[i] - yo4[i % 4]).abs().max()) / 255.0)
        else:
            for i in range(0, 32, 8):          # the oracle in slices of 8 (memory)
                yo = R.create_net(x[i:i + 8], starry, "resize", torch.float64)
                worst = max(worst, float((yc[i:i + 8] - yo).abs().max()) / 255.0)
    print("config 2 (%s) B=32 256^2: pixel max-abs/255 = %.3g" % (kind, worst))
    assert worst <= PIX_TOL, worst
    # frozen-weights inference (weights prepared once) must give the same bytes as the per-call preparation
    eng.set_frozen_weights(True)
    y1 = eng.transform_forward(p, x)
    y2 = eng.transform_forward(p, x)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2)
    assert float((y1 - y).abs().max()) / 255.0 <= 2e-6      # fp64 atomics of the fused statistics are order-dependent


def _tf_init_params(seed):
    """TF-1.0 initialisers of the transform net (im_transf_net.py:114,149,233-236): conv W ~ N(0, 0.1^2),
    upconv W ~ N(0, 1), InstanceNorm scale 1 / shift 0."""
    from faststyle_b200.layout import TRANSFORM_VARS
    rng = np.random.RandomState(seed)
    params = {}
    for name, shape in TRANSFORM_VARS:
        leaf = name.rsplit("/", 1)[1]
        if leaf.startswith("INscale"):
            params[name] = np.ones(shape, np.float32)
        elif leaf.startswith("INshift"):
            params[name] = np.zeros(shape, np.float32)
        else:
            std = 1.0 if ("upsample_0" in name or "upsample_1" in name) else 0.1
            params[name] = (rng.standard_normal(shape) * std).astype(np.float32)
    return params


@pytest.mark.parametrize("weights", ["starry", "tf_init"])
def test_config3_train_step_b4(built_lib, starry, golden_dir, weights):
    """BASELINE config 3 at full size: batch 4, 256x256, losses + all 48 gradients vs oracle autograd."""
    import cv2
    from faststyle_b200 import synth
    from faststyle_b200.engine import Engine, make_loss_config, pack_vgg, params_to_device
    from faststyle_b200.layout import transform_offsets
    _threads()
    N, H, W = 4, 256, 256
    vggw = synth.synthetic_vgg_weights(7)
    style = cv2.cvtColor(cv2.imread(os.path.join(golden_dir, "starry_night_crop.jpg")), cv2.COLOR_BGR2RGB)
    style = style.astype(np.float32)[None]                               # train.py:135-137 (no resize: scale 1.0)
    params = starry if weights == "starry" else _tf_init_params(1)
    x = _noise(N)
    with torch.no_grad():
        tg = R.style_target_grams(style, vggw, STYLE, torch.float32)     # 640x938 style image: fp32 on the host
    ref = R.train_grads(x, params, vggw, tg, dtype=torch.float64, beta=0.0)

    packed = pack_vgg(vggw, "cuda")
    cfg = make_loss_config(["conv3_3"], [1.0], STYLE, [5.0] * 4, 0.0)    # train.py defaults: cw 1, sw 5, beta 0
    eng = Engine(N, H, W, transform_bwd=True, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE)
    # the style Grams through the CUDA path too (train.py:144-151), checked against the oracle's
    seng = Engine(1, style.shape[1], style.shape[2], vgg=True, style_layers=STYLE)
    grams = seng.vgg_grams(packed, style, STYLE)
    for g, t in zip(grams, tg):
        e = float((g.double().cpu() - t.double()).abs().max() / t.double().abs().max())
        assert e < 1e-3, e
    del seng
    tgd = [t.float().cuda().contiguous() for t in tg]
    y = torch.empty((N, H, W, 3), device="cuda")
    grads, losses = eng.train_fwd_bwd(params_to_device(params, "cuda"), packed, x, cfg, tgd, y=y)
    torch.cuda.synchronize()
    pix = float((y.double().cpu() - ref["Y"]).abs().max()) / 255.0
    L = losses.cpu().double()
    lerr = [abs(float(got) - float(want)) / max(abs(float(want)), 1e-30)
            for got, want in zip(L[[0, 1, 3]], [ref["content"], ref["style"], ref["loss"]])]
    g = grads.cpu().double()
    offs = transform_offsets()
    flat_ref = torch.cat([ref["grads"][n].flatten() for n in offs])
    cos = float((g @ flat_ref) / (g.norm() * flat_ref.norm()))
    worst, worst_name = 0.0, ""
    for name, (off, shape) in offs.items():
        want = ref["grads"][name]
        got = g[off:off + want.numel()].view(want.shape)
        e = float((got - want).abs().max() / max(want.abs().max().item(), 1e-30))
        if e > worst:
            worst, worst_name = e, name
    print("config 3 (%s) B=4 256^2: pixel %.3g, loss rel errs %s, 1-cos %.3g, worst per-tensor grad %.3g (%s)" %
          (weights, pix, ["%.2g" % e for e in lerr], 1 - cos, worst, worst_name))
    assert pix <= PIX_TOL, pix
    assert max(lerr) <= LOSS_TOL, lerr
    # Untrained TF-initialiser weights (unit-variance resize-conv filters) make the gradient ill-conditioned: the
    # oracle ITSELF in fp32 vs fp64 differs by 1 - cos = 1.3e-6 / 3.4e-3 per tensor on this case (measured on the
    # host at batch 1), against 1e-10-class numbers for the trained checkpoint.  The 16-mantissa-bit tensor path
    # sits a few times above that floor (measured 5.9e-6 / 2.0e-2).
    cos_tol, grad_tol = (COS_TOL, GRAD_TOL) if weights == "starry" else (2e-5, 5e-2)
    assert 1 - cos <= cos_tol, 1 - cos
    assert worst <= grad_tol, (worst_name, worst)
    if weights == "tf_init":
        # the same case on the exact-fp32 (FFMA) path must sit AT the fp32 floor: this separates operand precision
        # from arithmetic slips in the pipeline
        eng.set_tensor_path(False)
        g32, _ = eng.train_fwd_bwd(params_to_device(params, "cuda"), packed, x, cfg, tgd, y=y)
        g32 = g32.cpu().double()
        cos32 = float((g32 @ flat_ref) / (g32.norm() * flat_ref.norm()))
        worst32 = max(float((g32[off:off + int(np.prod(shape))].view(shape) - ref["grads"][name]).abs().max()
                            / max(ref["grads"][name].abs().max().item(), 1e-30)) for name, (off, shape) in offs.items())
        print("   exact-fp32 path on the same case: 1-cos %.3g, worst per-tensor %.3g" % (1 - cos32, worst32))
        assert 1 - cos32 <= 5e-6 and worst32 <= 1e-2
