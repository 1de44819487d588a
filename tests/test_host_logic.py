"""CPU tests of the host-side mirror: variable store / Saver, CLI flag parity, input
pipeline (TFRecord framing + tf.Example parsing, TF-1 bicubic, shuffle/shard/epochs),
TensorBoard summaries."""
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_variable_scope_and_saver_roundtrip(tmp_path, golden_dir):
    from faststyle_b200 import variables as V
    from faststyle_b200 import tf_bundle
    V.reset_default_graph()
    with V.variable_scope('img_t_net'):
        with V.variable_scope('initconv_0'):
            w = V.get_variable('W', [9, 9, 3, 16], lambda s: np.zeros(s, np.float32))
    assert 'img_t_net/initconv_0/W' in V.all_variables()
    # restoring a checkpoint whose variable has another shape fails like TF's Saver
    bad = {'img_t_net/initconv_0/W': np.zeros((9, 9, 16, 3), np.float32)}
    tf_bundle.write_checkpoint(str(tmp_path / 'bad.ckpt'), bad)
    with pytest.raises(ValueError, match='shapes of both tensors'):
        V.Saver().restore(None, str(tmp_path / 'bad.ckpt'))
    with pytest.raises(KeyError, match='not found in checkpoint'):
        V.Saver(['img_t_net/missing']).restore(None, str(tmp_path / 'bad.ckpt'))
    # restore the shipped model, save it again: byte-identical files
    V.reset_default_graph()
    V.Saver().restore(None, os.path.join(golden_dir, 'starry_final.ckpt'))
    assert len(V.all_variables('img_t_net')) == 48
    out = V.Saver().save(None, str(tmp_path / 'again.ckpt'))
    for ext in ('.index', '.data-00000-of-00001'):
        assert open(out + ext, 'rb').read() == open(os.path.join(golden_dir, 'starry_final.ckpt') + ext, 'rb').read()
    assert V.Saver().save(None, str(tmp_path / 'm.ckpt'), global_step=12).endswith('m.ckpt-12')
    V.reset_default_graph()


REF_FLAGS = {
    'stylize_image.py': {'input_img_path': None, 'output_img_path': './results/styled.jpg',
                         'model_path': './models/starry_final.ckpt', 'content_target_resize': 1.0,
                         'upsample_method': 'resize'},
    'train.py': {'train_dir': None, 'model_name': None, 'style_img_path': './style_images/starry_night_crop.jpg',
                 'learn_rate': 1e-3, 'batch_size': 4, 'n_epochs': 2, 'preprocess_size': [256, 256],
                 'run_name': None, 'loss_content_layers': ['conv3_3'],
                 'loss_style_layers': ['conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'], 'content_weights': [1.0],
                 'style_weights': [5.0, 5.0, 5.0, 5.0], 'num_steps_ckpt': 1000, 'num_pipe_buffer': 4000,
                 'num_steps_break': -1, 'beta': 0.0, 'style_target_resize': 1.0, 'upsample_method': 'resize'},
    'slow_style.py': {'style_img_path': None, 'cont_img_path': None, 'learn_rate': 10.0,
                      'loss_content_layers': ['conv3_3'],
                      'loss_style_layers': ['conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'], 'content_weights': [1.0],
                      'style_weights': [5.0, 5.0, 5.0, 5.0], 'num_steps_break': 500, 'beta': 1e-4,
                      'style_target_resize': 1.0, 'cont_target_resize': 1.0, 'output_img_path': './out.jpg'},
}


@pytest.mark.parametrize('script', sorted(REF_FLAGS))
def test_cli_flags_match_reference(script):
    """Flag names and defaults of the reference parsers (stylize_image.py:19-43, train.py:23-105,
    slow_style.py:17-67), transcribed into REF_FLAGS."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('cli_' + script[:-3], os.path.join(ROOT, script))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ns = vars(mod.setup_parser().parse_args([]))
    assert ns == REF_FLAGS[script]


def _example(jpeg: bytes) -> bytes:
    def ld(field, payload):
        n, out = len(payload), bytearray([field << 3 | 2])
        while True:
            b = n & 127; n >>= 7
            out.append(b | (128 if n else 0))
            if not n:
                break
        return bytes(out) + payload
    feature = ld(1, ld(1, jpeg))                         # Feature{bytes_list{value}}
    entry = ld(1, b'image/encoded') + ld(2, feature)     # map entry
    other = ld(1, b'image/height') + ld(2, bytes([0x1a, 0x02, 0x08, 0x20]))   # int64_list{32}
    return ld(1, ld(1, other) + ld(1, entry))            # Example{features{feature...}}


def test_tfrecord_pipeline(tmp_path):
    import cv2
    from faststyle_b200 import datapipe
    rng = np.random.RandomState(0)
    recs = []
    for i in range(7):
        img = rng.randint(0, 256, (40 + i, 50, 3)).astype(np.uint8)
        ok, enc = cv2.imencode('.jpg', img)
        recs.append(_example(enc.tobytes()))
    with open(tmp_path / 'train-00000-of-00001', 'wb') as f:
        for r in recs:
            f.write(struct.pack('<Q', len(r)) + b'\0\0\0\0' + r + b'\0\0\0\0')
    assert len(list(datapipe.iter_tfrecord(str(tmp_path / 'train-00000-of-00001')))) == 7
    assert datapipe.example_bytes_feature(recs[0], 'image/encoded')[:2] == b'\xff\xd8'
    it = datapipe.batcher(str(tmp_path), 2, [32, 48], num_epochs=2, min_after_dequeue=3, seed=1)
    n = 0
    with pytest.raises(datapipe.OutOfRangeError):
        for b in it:
            assert b.shape == (2, 32, 48, 3) and b.dtype == np.float32
            n += 1
    assert n == 7                      # 14 images over 2 epochs, batches of 2
    # sharding: two ranks see disjoint halves
    a = list(_drain(datapipe.batcher(str(tmp_path), 1, None, 1, 1, seed=5, shard=(0, 2))))
    b = list(_drain(datapipe.batcher(str(tmp_path), 1, None, 1, 1, seed=5, shard=(1, 2))))
    assert len(a) + len(b) == 7 and abs(len(a) - len(b)) <= 1


def _drain(it):
    from faststyle_b200 import datapipe
    try:
        for b in it:
            yield b
    except datapipe.OutOfRangeError:
        return


def test_bicubic_tf1_properties():
    from faststyle_b200 import datapipe
    img = np.full((30, 40, 3), 77.0, np.float32)
    assert np.allclose(datapipe.resize_bicubic_tf1(img, 16, 24), 77.0, atol=1e-3)      # partition of unity
    ramp = np.tile(np.arange(40, dtype=np.float32)[None, :, None], (10, 1, 1))
    out = datapipe.resize_bicubic_tf1(ramp, 10, 20)                                    # scale 2: src = 2*dst
    assert np.allclose(out[0, 1:18, 0], 2.0 * np.arange(1, 18), atol=1e-3)
    assert datapipe.resize_bicubic_tf1(ramp, 10, 40).shape == (10, 40, 1)


def test_summary_writer(tmp_path):
    from faststyle_b200.summary import FileWriter
    w = FileWriter(str(tmp_path))
    w.add_scalars(10, {'summaries/loss': 1.5, 'summaries/style_loss': 1.0})
    w.close()
    files = [f for f in os.listdir(tmp_path) if 'tfevents' in f]
    assert files and os.path.getsize(tmp_path / files[0]) > 40


def test_loss_config_validation():
    from faststyle_b200.engine import make_loss_config
    cfg = make_loss_config(['vgg/conv3_3:0'], [1.0], ['conv1_2', 'conv4_3'], [5.0, 2.0], 1e-4)
    assert cfg.n_content == 1 and cfg.content_layer[0] == 6 and cfg.style_layer[1] == 9
    with pytest.raises(AssertionError):
        make_loss_config(['conv3_3'], [1.0, 2.0], [], [], 0.0)      # losses.py:23
    with pytest.raises(ValueError):
        make_loss_config(['conv5_1'], [1.0], [], [], 0.0)


def test_webcam_cli_keeps_reference_flags():
    """stylize_webcam.py:17-38: --model_path, --upsample_method, --resolution with the reference's defaults;
    our parser only ADDS flags (source / output / headless operation)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('cli_webcam', os.path.join(ROOT, 'stylize_webcam.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ns = vars(mod.setup_parser().parse_args([]))
    ref = {'model_path': './models/starry_final.ckpt', 'upsample_method': 'resize', 'resolution': None}
    assert {k: ns[k] for k in ref} == ref
    assert vars(mod.setup_parser().parse_args(['--resolution', '640', '480']))['resolution'] == [640, 480]
    assert set(ns) - set(ref) == {'source', 'output', 'no_display', 'max_frames'}


def test_deconv_variant_host_plumbing():
    """'deconv' models (im_transf_net.py:57-63,173,183): transposed weight shapes at the same flat offsets, unit-stddev
    initialisers for all three deconv2d layers, flatten/unflatten round trip, and train.py accepting the flag."""
    import importlib.util
    from faststyle_b200 import synth
    from faststyle_b200.layout import (TRANSFORM_NPARAMS, flatten_transform, transform_offsets, unflatten_transform)
    p = synth.init_transform_params(seed=3, upsample_method='deconv')
    assert p['img_t_net/upsample_0/W'].shape == (3, 3, 32, 64)
    assert p['img_t_net/upsample_1/W'].shape == (3, 3, 16, 32)
    assert p['img_t_net/upsample_2/W'].shape == (9, 9, 3, 16)
    assert 0.8 < p['img_t_net/upsample_2/W'].std() < 1.2 and 0.05 < p['img_t_net/initconv_0/W'].std() < 0.15
    r = synth.init_transform_params(seed=3, upsample_method='resize')
    assert r['img_t_net/upsample_2/W'].shape == (9, 9, 16, 3) and r['img_t_net/upsample_2/W'].std() < 0.15
    flat = flatten_transform(p, 'deconv')
    assert flat.shape == (TRANSFORM_NPARAMS,)
    back = unflatten_transform(flat, 'deconv')
    assert all(np.array_equal(back[k], p[k]) for k in p)
    assert [o for o, _ in transform_offsets('deconv').values()] == [o for o, _ in transform_offsets('resize').values()]
    with pytest.raises(ValueError):
        flatten_transform(p, 'resize')              # wrong variant for these shapes, like the reference's Saver
    spec = importlib.util.spec_from_file_location('cli_train2', os.path.join(ROOT, 'train.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.setup_parser().parse_args(['--upsample_method', 'deconv']).upsample_method == 'deconv'


def test_batcher_raw_mode():
    """raw=True hands the decoded uint8 images (un-resized) to the device-side preprocessor; same order and
    shuffle as the float path for the same seed."""
    from faststyle_b200 import datapipe
    a = datapipe.batcher('synthetic:12', 4, (256, 256), 1, 4, seed=5)
    b = datapipe.batcher('synthetic:12', 4, (256, 256), 1, 4, seed=5, raw=True)
    fa, rb = next(a), next(b)
    assert isinstance(rb, list) and len(rb) == 4 and rb[0].dtype == np.uint8 and rb[0].shape == (256, 256, 3)
    assert fa.shape == (4, 256, 256, 3) and np.array_equal(fa, np.stack(rb).astype(np.float32))


def test_prefetcher_order_and_termination():
    """Background input thread (the reference's queue runners, train.py:241-242): same batches in the same order,
    bounded depth, and the pipeline's OutOfRangeError reaches the training loop."""
    import time
    from faststyle_b200 import datapipe
    ref = list(_take(datapipe.batcher('synthetic:10', 2, (256, 256), 1, 2, seed=9), 100))
    pf = datapipe.prefetch(datapipe.batcher('synthetic:10', 2, (256, 256), 1, 2, seed=9), depth=2)
    got = []
    with pytest.raises(datapipe.OutOfRangeError):
        while True:
            got.append(next(pf))
    assert len(got) == len(ref) == 5 and all(np.array_equal(a, b) for a, b in zip(got, ref))
    with pytest.raises(datapipe.OutOfRangeError):       # stays terminated
        next(pf)

    def slow():
        for i in range(4):
            time.sleep(0.05)
            yield i
    pf = datapipe.prefetch(slow(), depth=3)
    time.sleep(0.3)                                      # producer runs ahead while the consumer is busy
    t0 = time.time()
    assert [next(pf) for _ in range(4)] == [0, 1, 2, 3]
    assert time.time() - t0 < 0.1
    with pytest.raises(StopIteration):
        next(pf)
    pf.close()


def _take(it, n):
    from faststyle_b200 import datapipe
    try:
        for _ in range(n):
            yield next(it)
    except datapipe.OutOfRangeError:
        return
