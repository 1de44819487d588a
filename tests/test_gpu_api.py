"""The Python mirror of the reference API and the three CLIs, on the GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ckpt as ockpt  # noqa: E402
from oracle import restate as R  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _relerr(got, want):
    want = torch.as_tensor(want).double()
    return float((torch.as_tensor(got).double().cpu() - want).abs().max() / max(want.abs().max().item(), 1e-30))


def test_stylize_image_cli_reproduces_golden(built_lib, golden_dir, tmp_path):
    """config 1: `stylize_image.py` on results/chicago.jpg with models/starry_final.ckpt
    (README.md:59-61) -> decoded output vs the reference's shipped results/starry_chicago.jpg."""
    import cv2
    out = str(tmp_path / "styled.jpg")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "stylize_image.py"),
                        "--input_img_path", os.path.join(golden_dir, "chicago.jpg"),
                        "--output_img_path", out,
                        "--model_path", os.path.join(golden_dir, "starry_final.ckpt")],
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    got = cv2.imread(out).astype(int)
    gold = cv2.imread(os.path.join(golden_dir, "starry_chicago.jpg")).astype(int)
    assert got.shape == gold.shape == (476, 712, 3)
    d = np.abs(got - gold)
    assert (d == 0).mean() >= 0.97 and d.mean() <= 0.05 and d.max() <= 10
    # wrong --upsample_method / checkpoint mismatch is an error, like the reference's Saver
    r = subprocess.run([sys.executable, os.path.join(ROOT, "stylize_image.py"), "--input_img_path",
                        os.path.join(golden_dir, "chicago.jpg"), "--model_path", str(tmp_path / "nope.ckpt")],
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode != 0 and "not found" in r.stderr


def test_mirror_ops_and_create_net(built_lib, golden_dir):
    from faststyle_b200 import im_transf_net as T
    from faststyle_b200 import variables as V
    p = ockpt.load(os.path.join(golden_dir, "candy_final.ckpt"))
    x = np.random.RandomState(2).randint(0, 256, (1, 72, 64, 3)).astype(np.float32)
    V.reset_default_graph()
    V.Saver().restore(None, os.path.join(golden_dir, "candy_final.ckpt"))
    with V.variable_scope('img_t_net'):
        y = T.create_net(x, 'resize')
        # op-by-op composition of the first layers (im_transf_net.py:34-40)
        h = T.reflect_pad(x, 40)
        with V.variable_scope('initconv_0'):
            h = T.relu(T.inst_norm(T.conv2d(h, 3, 16, 9, [1, 1, 1, 1])))
        with V.variable_scope('initconv_1'):
            h = T.relu(T.inst_norm(T.conv2d(h, 16, 32, 3, [1, 2, 2, 1])))
    taps = {}
    with torch.no_grad():
        yo = R.create_net(x, p, "resize", torch.float64, taps=taps)
    assert float((y.double().cpu() - yo).abs().max()) / 255.0 < 2e-4
    assert _relerr(h, taps["initconv_1"]) < 1e-5
    with pytest.raises(AssertionError):
        with V.variable_scope('img_t_net'):
            T.create_net(x, 'bilinear')
    V.reset_default_graph()


def test_deconv_variant_forward(built_lib):
    """upsample_method='deconv' (conv2d_transpose layers, im_transf_net.py:57-63,158-190): forward parity
    with the oracle restatement (unpinned: the reference ships no deconv checkpoint or golden)."""
    from faststyle_b200 import im_transf_net as T
    from faststyle_b200 import variables as V
    from faststyle_b200.layout import TRANSFORM_VARS_DECONV
    rng = np.random.RandomState(5)
    params = {}
    for name, shape in TRANSFORM_VARS_DECONV:
        leaf = name.rsplit("/", 1)[1]
        if leaf.startswith("INscale"):
            params[name] = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
        elif leaf.startswith("INshift"):
            params[name] = (0.1 * rng.standard_normal(shape)).astype(np.float32)
        else:
            params[name] = (rng.standard_normal(shape) * (0.3 if "upsample" in name else 0.1)).astype(np.float32)
    x = rng.randint(0, 256, (2, 56, 64, 3)).astype(np.float32)
    V.reset_default_graph()
    for k, v in params.items():
        V.set_variable(k, v)
    with V.variable_scope('img_t_net'):
        y = T.create_net(x, 'deconv')
    with torch.no_grad():
        yo = R.create_net(x, params, "deconv", torch.float64)
    assert tuple(y.shape) == tuple(yo.shape)
    assert float((y.double().cpu() - yo).abs().max()) / 255.0 < 2e-4
    # single transposed-conv op (stride 2, [k,k,cout,cin] weights)
    xs = rng.standard_normal((1, 9, 7, 8)).astype(np.float32)
    w = rng.standard_normal((3, 3, 4, 8)).astype(np.float32)
    V.reset_default_graph()
    V.set_variable('t/W', w)
    with V.variable_scope('t'):
        d = T.deconv2d(xs, 8, 4, 3, [1, 2, 2, 1])
    do = R.nchw_to_nhwc(R.deconv2d(R.nhwc_to_nchw(torch.from_numpy(xs).double()), torch.from_numpy(w).double(), 2))
    assert tuple(d.shape) == (1, 18, 14, 4) and _relerr(d, do) < 1e-5
    # a 'resize' checkpoint under --upsample_method deconv is a shape error, like the reference's Saver
    V.reset_default_graph()
    V.Saver().restore(None, os.path.join(os.path.dirname(__file__), "golden", "starry_final.ckpt"))
    with pytest.raises(ValueError):
        with V.variable_scope('img_t_net'):
            T.create_net(x, 'deconv')
    V.reset_default_graph()


def test_vgg_class_grams_and_losses(built_lib, tmp_path):
    from faststyle_b200 import losses, synth, utils
    from faststyle_b200 import variables as V
    from faststyle_b200.libs import vgg16
    w = synth.synthetic_vgg_weights(7)
    full = dict(w)
    full["conv5_1_W"] = np.zeros((3, 3, 512, 512), np.float32); full["conv5_1_b"] = np.zeros(512, np.float32)
    full["fc6_W"] = np.zeros((4, 4), np.float32); full["fc6_b"] = np.zeros(4, np.float32)
    np.savez(tmp_path / "vgg16_weights.npz", **full)
    x = np.random.RandomState(4).randint(0, 256, (2, 32, 48, 3)).astype(np.float32)
    V.reset_default_graph()
    with V.variable_scope('vgg'):
        net = vgg16.vgg16(x)
    net.load_weights(str(tmp_path / "vgg16_weights.npz"), None)     # sorted keys, stops at 'fc'
    names = ['vgg/conv1_2:0', 'vgg/conv3_3:0']
    layers = utils.get_layers(names)
    grams = utils.get_grams(names)
    L = R.vgg16_layers(x, w, "conv4_3", torch.float64)
    assert _relerr(layers[1], R.nchw_to_nhwc(L["conv3_3"])) < 2e-4
    assert _relerr(grams[0], R.gram(L["conv1_2"])) < 2e-4
    tgt = [torch.zeros_like(R.nchw_to_nhwc(L["conv3_3"]))]
    c = losses.content_loss([layers[1]], [t.float().cuda() for t in tgt], [2.0])
    assert abs(float(c) - float(R.content_loss([L["conv3_3"]], [torch.zeros_like(L["conv3_3"])], [2.0]))) <= 3e-4 * float(c)
    tg = [np.zeros((1, 64, 64), np.float32)]
    s = losses.style_loss([grams[0]], tg, [5.0])
    assert abs(float(s) - float(R.style_loss([R.gram(L["conv1_2"])], [torch.zeros(1, 64, 64, dtype=torch.float64)], [5.0]))) <= 3e-4 * float(s)
    tv = losses.tv_loss(x)
    assert abs(float(tv) - float(R.tv_loss(R.nhwc_to_nchw(torch.from_numpy(x).double())))) <= 1e-6 * float(tv)
    with pytest.raises(KeyError):
        utils.get_layers(['vgg/conv9_9:0'])
    V.reset_default_graph()


def test_train_cli_and_slow_style_cli(built_lib, golden_dir, tmp_path):
    """train.py for a few steps on synthetic data: checkpoints appear at the reference's paths,
    the final model loads back and stylizes; slow_style.py runs and lowers its loss."""
    import cv2
    env = dict(os.environ, VGG16_WEIGHTS="synthetic")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "--train_dir", "synthetic:64",
                        "--model_name", "tiny", "--style_img_path", os.path.join(golden_dir, "starry_night_crop.jpg"),
                        "--style_target_resize", "0.25", "--batch_size", "2", "--preprocess_size", "64", "64",
                        "--num_steps_break", "11", "--num_steps_ckpt", "10", "--num_pipe_buffer", "8"],
                       capture_output=True, text=True, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l.split() for l in r.stdout.splitlines() if l[:1].isdigit()]
    assert [int(l[0]) for l in lines] == [0, 10]                    # printed at ckpt / every-10 steps
    assert float(lines[1][1]) < float(lines[0][1])                  # loss decreases
    for f in ("training/tiny.ckpt-0.index", "training/tiny.ckpt-10.index", "models/tiny_final.ckpt.index",
              "models/tiny_final.ckpt.data-00000-of-00001"):
        assert os.path.exists(tmp_path / f), f
    assert any("tfevents" in f for f in os.listdir(tmp_path / "summaries/train/tiny0"))
    from faststyle_b200 import tf_bundle
    final = tf_bundle.read_checkpoint(str(tmp_path / "models/tiny_final.ckpt"))
    assert len(final) == 48                                          # trainable vars only (train.py:225)
    full = tf_bundle.read_checkpoint(str(tmp_path / "training/tiny.ckpt-10"))
    assert "global_step" in full and int(full["global_step"]) == 10 and "img_t_net/initconv_0/W/Adam_1" in full
    # the trained model is usable by the stylize CLI
    r = subprocess.run([sys.executable, os.path.join(ROOT, "stylize_image.py"), "--input_img_path",
                        os.path.join(golden_dir, "chicago.jpg"), "--content_target_resize", "0.25",
                        "--output_img_path", str(tmp_path / "o.jpg"), "--model_path", "models/tiny_final.ckpt"],
                       capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert cv2.imread(str(tmp_path / "o.jpg")) is not None
    # slow_style
    small = cv2.resize(cv2.imread(os.path.join(golden_dir, "chicago.jpg")), (96, 64))
    cv2.imwrite(str(tmp_path / "c.jpg"), small)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "slow_style.py"), "--style_img_path",
                        os.path.join(golden_dir, "starry_night_crop.jpg"), "--style_target_resize", "0.2",
                        "--cont_img_path", str(tmp_path / "c.jpg"), "--num_steps_break", "20",
                        "--output_img_path", str(tmp_path / "s.jpg")],
                       capture_output=True, text=True, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    ls = [l.split() for l in r.stdout.splitlines() if l[:1].isdigit()]
    assert [int(l[0]) for l in ls] == [0, 10, 20] and float(ls[-1][1]) < float(ls[0][1])
    assert cv2.imread(str(tmp_path / "s.jpg")).shape == (64, 96, 3)


def test_frame_stylizer_and_webcam_cli(built_lib, golden_dir, tmp_path):
    """the stylize_webcam.py loop body (:76-103).  A uint8 BGR frame is fed un-swapped, the
    float result is truncated (ndarray.astype(uint8)) and channels 0/2 are swapped; the CUDA-graph replay
    and the plain-launch form must agree bit for bit, and both with the oracle within one grey level."""
    import subprocess
    import sys
    import cv2
    from faststyle_b200.stream import FrameStylizer
    from oracle import ckpt as ockpt
    params = ockpt.load(os.path.join(golden_dir, "starry_final.ckpt"))
    rng = np.random.RandomState(2)
    H, W = 72, 88                                    # (H+80)%4 == 0: output size == input size
    frames = [rng.randint(0, 256, (H, W, 3)).astype(np.uint8) for _ in range(3)]
    fs_g = FrameStylizer(params, H, W, use_graph=True)
    fs_e = FrameStylizer(params, H, W, use_graph=False)
    assert fs_e.graph is None
    for f in frames:
        a, b = fs_g.stylize(f), fs_e.stylize(f)
        assert a.dtype == np.uint8 and a.shape == (H, W, 3)
        assert np.array_equal(a, b)
        with torch.no_grad():
            yo = R.create_net(f[np.newaxis].astype(np.float32), params, "resize", torch.float64)[0].numpy()
        want = np.clip(yo, 0, 255).astype(np.uint8)[..., ::-1]
        d = np.abs(a.astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and (d != 0).mean() < 0.05      # truncation flips only values within ~0.01 level of an integer
    print("frame stylizer: CUDA graph captured =", fs_g.graph is not None)

    # the CLI on a video file instead of camera 0
    src = str(tmp_path / "in.avi")
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*'MJPG'), 15.0, (W, H))
    assert wr.isOpened()
    for f in frames:
        wr.write(f)
    wr.release()
    out = str(tmp_path / "output.avi")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "stylize_webcam.py"), "--model_path",
                        os.path.join(golden_dir, "starry_final.ckpt"), "--source", src, "--output", out,
                        "--no_display"], capture_output=True, text=True, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Resolution is: %d by %d" % (W, H) in r.stdout and "Stylised 3 frames" in r.stdout
    cap = cv2.VideoCapture(out)
    n = 0
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        assert fr.shape == (H, W, 3)
        n += 1
    assert n == 3


@pytest.mark.parametrize("shape", [((480, 640), (256, 256)), ((97, 131), (256, 256)), ((256, 256), (256, 256)),
                                   ((333, 500), (128, 192)), ((5, 7), (16, 9))])
def test_gpu_bicubic_resize_matches_oracle(built_lib, shape):
    """datapipe.preprocessing (datapipe.py:14-26) on the device, against the independent restatement
    of TF-1.0's ResizeBicubic in oracle/bicubic.py (and the product's own host restatement).  The kernel follows
    TF's arithmetic order (four horizontal interpolations, then one vertical; table entries evaluated in double;
    no FMA contraction).  Tolerance stated: max-abs 1e-4 on the 0..255 scale (bit-exact in practice).
    Unpinned against TensorFlow itself (tools/dump_tf1_goldens.py writes the golden for a TF-1.x owner)."""
    from faststyle_b200 import datapipe
    from faststyle_b200.ops import resize_bicubic_tf1
    from oracle.bicubic import resize_bicubic_tf1 as oracle_bicubic
    (h, w), (oh, ow) = shape
    rng = np.random.RandomState(h * 7 + w)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    want = oracle_bicubic(img, oh, ow)
    got = resize_bicubic_tf1(img, oh, ow).cpu().numpy()
    d = np.abs(got - want)
    print("bicubic", shape, "max abs diff vs oracle", d.max(), "exact", bool((got == want).all()))
    assert got.shape == (oh, ow, 3) and d.max() <= 1e-4
    assert np.abs(got - datapipe.resize_bicubic_tf1(img, oh, ow)).max() <= 1e-4


def test_gpu_preprocessor_batch(built_lib):
    from faststyle_b200 import datapipe
    rng = np.random.RandomState(0)
    imgs = [rng.randint(0, 256, s).astype(np.uint8) for s in [(300, 400, 3), (256, 256, 3), (128, 500, 3)]]
    prep = datapipe.GpuPreprocessor(3, (256, 256))
    got = prep(imgs).cpu().numpy()
    want = np.stack([datapipe.resize_bicubic_tf1(i, 256, 256) for i in imgs])
    assert got.shape == (3, 256, 256, 3) and np.abs(got - want).max() <= 1e-4
    got2 = prep(imgs[::-1]).cpu().numpy()           # staging buffer reuse
    assert np.abs(got2 - want[::-1]).max() <= 1e-4


def test_batch_stylizer_pipeline(built_lib, golden_dir):
    """BatchStylizer (pipelined uint8 -> uint8 batch inference, the e2e leg of bench.py's config-2 block): results
    equal round-and-saturate of Engine.transform_forward, and the oracle within one grey level; two batches in
    flight come back in order."""
    import os
    from faststyle_b200.engine import Engine, params_to_device
    from faststyle_b200.stream import BatchStylizer
    from faststyle_b200.tf_bundle import read_checkpoint
    from oracle import restate as R
    params = read_checkpoint(os.path.join(golden_dir, "starry_final.ckpt"))
    rng = np.random.RandomState(5)
    B, H, W = 3, 64, 80
    xs = [rng.randint(0, 256, (B, H, W, 3)).astype(np.uint8) for _ in range(3)]
    st = BatchStylizer(params, B, H, W)
    eng = Engine(B, H, W, transform=True)
    p = params_to_device(params, "cuda")
    st.submit(xs[0]); st.submit(xs[1])
    outs = [st.fetch()]
    st.submit(xs[2])
    outs += [st.fetch(), st.fetch()]
    for x, got in zip(xs, outs):
        y = eng.transform_forward(p, x.astype(np.float32)).cpu().numpy()
        want = np.clip(np.rint(y), 0, 255).astype(np.uint8)
        assert got.shape == want.shape and got.dtype == np.uint8
        assert (got != want).mean() < 1e-3          # fused-statistics atomics: last-bit differences at .5 boundaries
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1
    with torch.no_grad():
        yo = R.create_net(xs[2].astype(np.float32), params, "resize", torch.float64).numpy()
    assert np.abs(outs[2].astype(float) - np.clip(yo, 0, 255)).max() <= 0.5 + 0.26      # rounding + 1e-3 of the unit scale
