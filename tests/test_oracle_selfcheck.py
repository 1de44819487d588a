"""Self-consistency checks of the ORACLE itself (CPU): finite differences of its loss path and hand-computed
known answers of the TF-1.0 bicubic restatement.  These do not pin the oracle to TensorFlow (nothing in this
container can: see oracle/__init__.py) - they rule out slips inside the restatement (a detached branch, a wrong
normaliser, a swapped weight) that autograd alone would faithfully differentiate."""
import numpy as np
import torch

from oracle import bicubic as OB
from oracle import restate as R

STYLE = ("conv1_2", "conv2_2", "conv3_3", "conv4_3")


def _tiny_setup(seed=0):
    from faststyle_b200 import synth
    from faststyle_b200.layout import TRANSFORM_VARS
    rng = np.random.RandomState(seed)
    params = {}
    for name, shape in TRANSFORM_VARS:
        leaf = name.rsplit("/", 1)[1]
        if leaf.startswith("INscale"):
            params[name] = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float64)
        elif leaf.startswith("INshift"):
            params[name] = (0.1 * rng.standard_normal(shape)).astype(np.float64)
        else:
            params[name] = (rng.standard_normal(shape) * 0.1).astype(np.float64)
    vggw = synth.synthetic_vgg_weights(7)
    style = rng.randint(0, 256, (1, 24, 28, 3)).astype(np.float64)
    tg = R.style_target_grams(style, vggw, STYLE, torch.float64)
    return params, vggw, tg, rng


def test_train_loss_gradient_finite_differences():
    """Directional derivatives of the train.py:178-184 loss w.r.t. the 48 variables: central differences in fp64
    vs the gradients the oracle hands to the parity tests."""
    params, vggw, tg, rng = _tiny_setup()
    x = rng.randint(0, 256, (1, 44, 48, 3)).astype(np.float64)
    kw = dict(beta=1e-3)
    ref = R.train_grads(x, params, vggw, tg, dtype=torch.float64, **kw)

    def loss_at(p):
        with torch.no_grad():
            return float(R.train_losses(x, {k: torch.from_numpy(v) for k, v in p.items()}, vggw, tg,
                                        dtype=torch.float64, **kw)[0])

    # one direction per variable family + one over everything
    families = ["initconv_0/W", "initconv_1/INscale", "resblock_2/W1", "resblock_4/INshift2", "upsample_0/W",
                "upsample_1/W", "upsample_2/W", "upsample_2/INshift", None]
    for fam in families:
        d = {k: (rng.standard_normal(v.shape) if (fam is None or k.endswith(fam)) else np.zeros_like(v))
             for k, v in params.items()}
        analytic = sum(float((ref["grads"][k] * torch.from_numpy(d[k])).sum()) for k in params)
        # ReLU / max-pool kinks make the loss only piecewise smooth (a step of 1e-6 already crosses enough of them
        # to be off by 1e-2): steps small enough to stay on one piece, where fp64 still resolves the difference
        best = None
        for eps in (1e-7, 1e-8):
            lp = loss_at({k: v + eps * d[k] for k, v in params.items()})
            lm = loss_at({k: v - eps * d[k] for k, v in params.items()})
            fd = (lp - lm) / (2 * eps)
            err = abs(fd - analytic) / max(abs(analytic), 1e-12)
            best = err if best is None else min(best, err)
        assert best < 1e-5, (fam, analytic, best)


def test_slow_style_pixel_gradient_finite_differences():
    """slow_style.py:140-154: dLoss/dX of the variable image vs central differences."""
    _, vggw, tg, rng = _tiny_setup(1)
    content = rng.randint(0, 256, (1, 32, 36, 3)).astype(np.float64)
    xvar = rng.uniform(0, 255, (1, 32, 36, 3))
    with torch.no_grad():
        ct = [R.vgg16_layers(content, vggw, "conv3_3", torch.float64)["conv3_3"]]
    ref = R.slow_style_grads(xvar, ct, vggw, tg, beta=1e-2, dtype=torch.float64)
    d = rng.standard_normal(xvar.shape)
    analytic = float((ref["grad"] * torch.from_numpy(d)).sum())

    def loss_at(xv):
        x = torch.from_numpy(xv).requires_grad_(False)
        L = R.vgg16_layers(x, vggw, "conv4_3", torch.float64)
        c = R.content_loss([L["conv3_3"]], ct, (1.0,))
        s = R.style_loss([R.gram(L[n]) for n in STYLE], tg, (5.0,) * 4)
        return float(c + s + 1e-2 * R.tv_loss(R.nhwc_to_nchw(x)))

    assert abs(loss_at(xvar) - float(ref["loss"])) <= 1e-9 * abs(float(ref["loss"]))
    best = min(abs((loss_at(xvar + e * d) - loss_at(xvar - e * d)) / (2 * e) - analytic) / abs(analytic)
               for e in (1e-6, 1e-7))
    assert best < 1e-5, (analytic, best)


def test_losses_hand_values():
    """losses.py:12-97 / utils.py:66-83 on hand-checkable tensors."""
    f = torch.arange(2 * 3 * 2 * 2, dtype=torch.float64).reshape(2, 3, 2, 2)       # NCHW: b=2, c=3, h=w=2
    G = R.gram(f)                                                                  # [2,3,3] = F^T F / (h w c)
    F0 = f[0].reshape(3, 4)
    assert torch.allclose(G[0], F0 @ F0.T / 12.0)
    t = torch.zeros_like(f)
    # content: w * sum_all (f - t)^2 / (h w c), the sum INCLUDES the batch, the normaliser does not
    assert abs(float(R.content_loss([f], [t], (2.0,))) - 2.0 * float((f ** 2).sum()) / 12.0) < 1e-9
    T = torch.zeros(1, 3, 3, dtype=torch.float64)
    assert abs(float(R.style_loss([G], [T], (3.0,))) - 3.0 * float((G ** 2).sum()) / 9.0) < 1e-12
    x = torch.tensor([[[[0.0, 1.0], [3.0, 6.0]]]], dtype=torch.float64)            # [1,1,2,2]
    # vertical diffs: (0-3)^2 + (1-6)^2 = 34; horizontal: (0-1)^2 + (3-6)^2 = 10
    assert abs(float(R.tv_loss(x)) - 44.0) < 1e-12


def test_tf_adam_rule_hand_value():
    """TF-Adam: epsilon is added to sqrt(v) without bias correction."""
    p = {"w": torch.tensor([1.0], dtype=torch.float64)}
    opt = R.TFAdam(p, lr=0.1)
    opt.step(p, {"w": torch.tensor([0.5], dtype=torch.float64)})
    m, v = 0.05, 0.00025
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert abs(float(p["w"]) - (1.0 - lr_t * m / (np.sqrt(v) + 1e-8))) < 1e-12


def test_bicubic_known_answers():
    """Keys cubic with A = -0.75: weights at delta 0 are (0,1,0,0), at delta 0.5 (-3/32, 19/32, 19/32, -3/32);
    TF-1.0 maps src = dst * in/out without half-pixel centres and clamps the border taps."""
    w, idx = OB.weights_and_indices(8, 16)                # scale 0.5: even outputs delta 0, odd outputs delta 0.5
    assert np.array_equal(w[0], np.float32([0, 1, 0, 0]))
    assert np.allclose(w[1], [-0.09375, 0.59375, 0.59375, -0.09375], atol=1e-7)
    assert np.allclose(w.sum(1), 1.0, atol=1e-6)
    assert idx[0].tolist() == [0, 0, 1, 2] and idx[15].tolist() == [6, 7, 7, 7]
    ramp = np.arange(8, dtype=np.float32).reshape(1, 8, 1).repeat(3, 0)
    out = OB.resize_bicubic_tf1(ramp, 3, 16)[1, :, 0]
    assert out[4] == 2.0 and abs(out[5] - 2.5) < 1e-6      # interior of a linear ramp is reproduced
    assert abs(out[15] - (-0.09375 * 6 + 0.59375 * 14 - 0.09375 * 7)) < 1e-5   # right border: taps 6,7,7,7
    img = np.random.RandomState(0).randint(0, 256, (9, 11, 3)).astype(np.uint8)
    assert np.array_equal(OB.resize_bicubic_tf1(img, 9, 11), img.astype(np.float32))          # identity size
    sub = OB.resize_bicubic_tf1(img[:8, :10], 4, 5)                                          # exact 2x: subsampling
    assert np.array_equal(sub, img[:8:2, :10:2].astype(np.float32))


def test_bicubic_oracle_vs_product_host_restatement():
    """The product's host-side restatement (faststyle_b200/datapipe.py) against the oracle's, written
    independently (different vectorisation; same TF arithmetic order): float32-rounding agreement."""
    from faststyle_b200 import datapipe
    rng = np.random.RandomState(3)
    for (h, w, oh, ow) in [(480, 640, 256, 256), (37, 53, 256, 256), (300, 200, 64, 96), (7, 5, 20, 30)]:
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        a, b = OB.resize_bicubic_tf1(img, oh, ow), datapipe.resize_bicubic_tf1(img, oh, ow)
        assert a.shape == b.shape == (oh, ow, 3)
        assert np.abs(a - b).max() <= 1e-4
