"""Consumes tests/golden/tf1/*.npz written by tools/dump_tf1_goldens.py (TensorFlow 1.x running the REFERENCE's own
graph).  The files cannot be produced in the build container (no TensorFlow), so every test here skips while they
are absent; once a TF-1.x owner drops them in, the CPU tests pin the oracle's loss / Gram / gradient / Adam /
bicubic restatements to the reference and the `-m gpu` tests pin the CUDA path directly."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
TF1 = os.path.join(HERE, "golden", "tf1")
STYLE = ("conv1_2", "conv2_2", "conv3_3", "conv4_3")


def _load(name):
    p = os.path.join(TF1, name)
    if not os.path.exists(p):
        pytest.skip("tests/golden/tf1/%s absent: run tools/dump_tf1_goldens.py under TensorFlow 1.x" % name)
    return np.load(p)


def _vgg_weights(meta):
    if bool(meta["real_vgg"]):
        path = os.environ.get("VGG16_WEIGHTS", os.path.join(HERE, "..", "libs", "vgg16_weights.npz"))
        if not os.path.exists(path):
            pytest.skip("goldens were made with the real vgg16_weights.npz; set VGG16_WEIGHTS to it")
        w = np.load(path)
        return {k: w[k] for k in w.keys() if "fc" not in k and not k.startswith("conv5")}
    from faststyle_b200 import synth
    return synth.synthetic_vgg_weights(7)


def _starry():
    from oracle import ckpt as ockpt
    return ockpt.load(os.path.join(HERE, "golden", "starry_final.ckpt"))


def _style_image():
    import cv2
    img = cv2.cvtColor(cv2.imread(os.path.join(HERE, "golden", "starry_night_crop.jpg")), cv2.COLOR_BGR2RGB)
    return img[None].astype(np.float32)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ------------------------------------------------------------------------------------------------ CPU: the oracle
def test_oracle_transform_forward_vs_tf1():
    from oracle import restate as R
    g = _load("transform_fwd.npz")
    x = np.random.RandomState(0).randint(0, 256, (2, 256, 256, 3)).astype(np.float32)
    with torch.no_grad():
        y = R.create_net(x, _starry(), "resize", torch.float64).numpy()
    assert np.abs(y - g["Y"]).max() / 255.0 <= 1e-4           # TF's own fp32 noise floor is ~2e-5


def test_oracle_style_grams_vs_tf1():
    from oracle import restate as R
    g = _load("style_grams.npz")
    tg = R.style_target_grams(_style_image(), _vgg_weights(g), STYLE, torch.float32)
    for n, t in zip(STYLE, tg):
        assert _rel(t.numpy(), g[n]) <= 1e-4, n


def test_oracle_train_step_vs_tf1():
    from oracle import restate as R
    g = _load("train_step.npz")
    sg = _load("style_grams.npz")
    vggw = _vgg_weights(g)
    x = np.random.RandomState(3).randint(0, 256, (2, 128, 128, 3)).astype(np.float32)
    tg = [torch.from_numpy(sg[n]) for n in STYLE]
    params = _starry()
    ref = R.train_grads(x, params, vggw, tg, dtype=torch.float64, beta=float(g["beta"]))
    assert np.abs(ref["Y"].numpy() - g["Y"]).max() / 255.0 <= 1e-4
    for k in ("content", "style", "loss"):
        assert abs(float(ref[k]) - float(g[k])) <= 1e-4 * abs(float(g[k])), k
    assert abs(float(ref["tv"]) - float(g["beta"]) * float(g["tv"])) <= 1e-4 * float(g["beta"]) * float(g["tv"])
    for name, want in ((k[len("grad/"):], g[k]) for k in g.files if k.startswith("grad/")):
        assert _rel(ref["grads"][name].numpy(), want) <= 2e-3, name     # TF computes these in fp32
    # one TF-Adam step
    p = {k: torch.from_numpy(v).double() for k, v in params.items()}
    R.TFAdam(p, 1e-3).step(p, ref["grads"])
    # the first Adam update is ~ lr * sign(g): compare where the golden gradient is clearly non-zero (a sign flip
    # of a ~0 gradient between fp32 and fp64 moves the variable by 2 lr)
    for name, want in ((k[len("after_adam/"):], g[k]) for k in g.files if k.startswith("after_adam/")):
        mask = np.abs(g["grad/" + name]) > 1e-5
        assert mask.any() and np.abs(p[name].numpy() - want)[mask].max() <= 2e-5, name       # |update| <= lr = 1e-3


def test_oracle_bicubic_vs_tf1():
    import cv2
    from oracle.bicubic import resize_bicubic_tf1
    g = _load("bicubic.npz")
    img = cv2.cvtColor(cv2.imread(os.path.join(HERE, "golden", "chicago.jpg")), cv2.COLOR_BGR2RGB)
    got = resize_bicubic_tf1(img, 256, 256)
    assert np.abs(got - g["resized"]).max() <= 1e-3


# ------------------------------------------------------------------------------------------------ GPU: the product
@pytest.mark.gpu
def test_cuda_train_step_vs_tf1(built_lib):
    from faststyle_b200.engine import Engine, make_loss_config, pack_vgg, params_to_device
    from faststyle_b200.layout import transform_offsets
    g = _load("train_step.npz")
    sg = _load("style_grams.npz")
    vggw = _vgg_weights(g)
    x = np.random.RandomState(3).randint(0, 256, (2, 128, 128, 3)).astype(np.float32)
    packed = pack_vgg(vggw, "cuda")
    cfg = make_loss_config(["conv3_3"], [1.0], STYLE, [5.0] * 4, float(g["beta"]))
    eng = Engine(2, 128, 128, transform_bwd=True, vgg_bwd=True, content_layers=["conv3_3"], style_layers=STYLE)
    tgd = [torch.from_numpy(sg[n]).float().cuda().contiguous() for n in STYLE]
    y = torch.empty((2, 128, 128, 3), device="cuda")
    grads, losses = eng.train_fwd_bwd(params_to_device(_starry(), "cuda"), packed, x, cfg, tgd, y=y)
    torch.cuda.synchronize()
    assert float(np.abs(y.cpu().numpy() - g["Y"]).max()) / 255.0 <= 1e-3          # north_star's pixel tolerance
    L = losses.cpu().numpy()
    assert abs(L[0] - float(g["content"])) <= 1e-3 * abs(float(g["content"]))      # north_star's loss tolerance
    assert abs(L[1] - float(g["style"])) <= 1e-3 * abs(float(g["style"]))
    gflat = grads.cpu().numpy()
    for name, (off, shape) in transform_offsets().items():
        want = g["grad/" + name]
        assert _rel(gflat[off:off + want.size].reshape(want.shape), want) <= 1e-2, name


@pytest.mark.gpu
def test_cuda_bicubic_vs_tf1(built_lib):
    import cv2
    from faststyle_b200.ops import resize_bicubic_tf1
    g = _load("bicubic.npz")
    img = cv2.cvtColor(cv2.imread(os.path.join(HERE, "golden", "chicago.jpg")), cv2.COLOR_BGR2RGB)
    got = resize_bicubic_tf1(img, 256, 256).cpu().numpy()
    assert np.abs(got - g["resized"]).max() <= 1e-3
