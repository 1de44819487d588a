"""G1/G2: the oracle restatement of create_net reproduces the reference's shipped golden
images (README.md:5-18,59-61: results/chicago.jpg -> results/{starry,candy}_chicago.jpg with
models/{starry,candy}_final.ckpt) after cv2's default JPEG q95 - this pins the oracle."""
import os

import cv2
import numpy as np
import pytest
import torch

from oracle import ckpt as ockpt
from oracle import restate as R


@pytest.mark.parametrize("model", ["starry", "candy"])
def test_golden_chicago(golden_dir, model):
    torch.set_num_threads(os.cpu_count() or 1)
    img = cv2.cvtColor(cv2.imread(os.path.join(golden_dir, "chicago.jpg")), cv2.COLOR_BGR2RGB)
    assert img.shape == (474, 712, 3)
    p = ockpt.load(os.path.join(golden_dir, model + "_final.ckpt"))
    with torch.no_grad():
        y = R.create_net(img[None].astype(np.float32), p, "resize").numpy()[0]
    assert y.shape == (476, 712, 3)                       # (H+80)%4 != 0 grows H
    out = np.clip(np.rint(y), 0, 255).astype(np.uint8)    # cv2.imwrite saturate-rounds float32
    ok, enc = cv2.imencode(".jpg", cv2.cvtColor(out, cv2.COLOR_RGB2BGR))
    dec = cv2.imdecode(enc, cv2.IMREAD_COLOR).astype(int)
    gold = cv2.imread(os.path.join(golden_dir, model + "_chicago.jpg")).astype(int)
    d = np.abs(dec - gold)
    assert (d == 0).mean() >= 0.985, (d == 0).mean()
    assert d.mean() <= 0.02 and d.max() <= 8
    assert abs(len(enc) - os.path.getsize(os.path.join(golden_dir, model + "_chicago.jpg"))) < 200


def test_same_padding_rule():
    assert R.same_pad(336, 3, 2) == (0, 1)
    assert R.same_pad(277, 3, 2) == (1, 1)
    assert R.same_pad(336, 9, 1) == (4, 4)


def test_resize_conv_collapse_identity():
    """NN x4 + 3x3 s2 SAME == 4-phase 2x2 conv on the un-resized source."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 5, 7, 6, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, 5, 4, generator=g, dtype=torch.float64)
    ref = R.upconv2d(x, w, 2)
    Rm = [torch.tensor([[1., 1, 1], [0, 0, 0]], dtype=torch.float64),
          torch.tensor([[1., 1, 0], [0, 0, 1]], dtype=torch.float64)]
    xp = torch.nn.functional.pad(x, (0, 1, 0, 1))
    out = torch.zeros_like(ref)
    for p in range(2):
        for q in range(2):
            wc = torch.einsum("ak,bl,klio->abio", Rm[p], Rm[q], w)
            out[:, :, p::2, q::2] = torch.nn.functional.conv2d(xp, wc.permute(3, 2, 0, 1))
    assert float((out - ref).abs().max()) < 1e-12


def test_oracle_fp32_vs_fp64_noise_floor(golden_dir):
    """G4: fp32-vs-fp64 restatement noise at 128^2 stays ~1e-5 on the unit scale."""
    p = ockpt.load(os.path.join(golden_dir, "starry_final.ckpt"))
    x = np.random.RandomState(0).randint(0, 256, (1, 128, 128, 3)).astype(np.float32)
    with torch.no_grad():
        a = R.create_net(x, p, "resize", torch.float32).double()
        b = R.create_net(x, p, "resize", torch.float64)
    assert float((a - b).abs().max()) / 255.0 < 1e-4


def test_tf_adam_rule():
    p = {"w": torch.tensor([1.0, -2.0], dtype=torch.float64)}
    opt = R.TFAdam(p, 0.1)
    g = {"w": torch.tensor([0.5, -0.25], dtype=torch.float64)}
    opt.step(p, g)
    # t=1: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1) -> step = lr*g/(|g| + eps*sqrt(1-b2)) ~ lr*sign(g)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    want = np.array([1.0, -2.0]) - lr_t * (0.1 * np.array([0.5, -0.25])) / (np.sqrt(0.001 * np.array([0.25, 0.0625])) + 1e-8)
    assert np.allclose(p["w"].numpy(), want, rtol=0, atol=1e-12)


def test_loss_definitions_small():
    """losses.py:12-97 on hand-checkable inputs."""
    l = torch.arange(8, dtype=torch.float64).view(1, 2, 2, 2)
    t = torch.zeros_like(l)
    assert float(R.content_loss([l], [t], [2.0])) == pytest.approx(2.0 * float((l ** 2).sum()) / 8)
    G = R.gram(l)
    f = l.view(1, 2, 4)
    assert torch.allclose(G, torch.bmm(f, f.transpose(1, 2)) / 8)
    assert float(R.style_loss([G], [torch.zeros(1, 2, 2, dtype=torch.float64)], [3.0])) == pytest.approx(3.0 * float((G ** 2).sum()) / 4)
    x = torch.tensor([[[[0., 1.], [3., 6.]]]], dtype=torch.float64)
    assert float(R.tv_loss(x)) == pytest.approx((1 ** 2 + 3 ** 2) + (3 ** 2 + 5 ** 2))


def test_oracle_deconv_is_conv_gradient():
    """tf.nn.conv2d_transpose(SAME) is DEFINED as the gradient of conv2d(SAME) w.r.t. its input
    (im_transf_net.py:158-190): the oracle's deconv2d must equal autograd of conv2d_tf."""
    g = torch.Generator().manual_seed(3)
    for stride, k, hw in ((2, 3, (6, 5)), (1, 9, (7, 8)), (2, 3, (5, 7))):
        x = torch.randn(2, 4, hw[0], hw[1], generator=g, dtype=torch.float64)          # transpose input: cin=4
        w = torch.randn(k, k, 3, 4, generator=g, dtype=torch.float64)                  # [k,k,cout=3,cin=4]
        big = torch.zeros(2, 3, hw[0] * stride, hw[1] * stride, dtype=torch.float64, requires_grad=True)
        y = R.conv2d_tf(big, w, stride, "SAME")                                        # conv: 3 -> 4 channels
        assert tuple(y.shape) == tuple(x.shape)
        (gin,) = torch.autograd.grad(y, big, x)
        assert float((R.deconv2d(x, w, stride) - gin).abs().max()) < 1e-10
