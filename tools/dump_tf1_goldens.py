#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Pinning kit: dump TensorFlow-1.x outputs of the REFERENCE's own graph into tests/golden/tf1/.

The loss / Gram / gradient / Adam path and the TF-1.0 bicubic resize are "parity unpinned" in this repository:
TensorFlow cannot be installed in the build container, so nothing produced by the reference itself exists for
them.  Anyone with TensorFlow 1.x (1.0 - 1.15; Python 2.7 or 3.x) closes that gap with ONE command, run from a
checkout of ghwatson/faststyle (the script imports the reference's own im_transf_net / libs.vgg16 / utils / losses,
i.e. the golden vectors come from the reference code, not from a restatement):

    cd /path/to/faststyle
    python /path/to/this/repo/tools/dump_tf1_goldens.py --out /path/to/this/repo/tests/golden/tf1 \\
           [--vgg libs/vgg16_weights.npz]

and then `python -m pytest tests/test_tf1_goldens.py` in this repository (CPU part: the oracle against the
goldens; `-m gpu` part: the CUDA path against them).  Without --vgg (or when the file is absent) the same seeded
synthetic VGG weights as the rest of the test-suite are used, written with numpy only.

What is written (all inputs are regenerated from numpy seeds by the consuming test, so only outputs are stored):
  transform_fwd.npz   Y = create_net(X) for models/starry_final.ckpt, X = RandomState(0) uint8 [2,256,256,3]
  train_step.npz      train.py:158-204 graph at batch 2, 128x128: content/style/tv/total loss, the 4 Grams of the
                      stylised batch, all 48 gradients (tf.gradients of the loss w.r.t. the img_t_net variables),
                      and the variables after ONE AdamOptimizer(1e-3) step
  style_grams.npz     target Grams of style_images/starry_night_crop.jpg (train.py:143-151)
  bicubic.npz         tf.image.resize_images(results/chicago.jpg, [256,256], method=2) (datapipe.py:25)
The file is deliberately Python-2 compatible (the reference is Python 2)."""
from __future__ import print_function

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())

VGG_CONV = [("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128),
            ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256),
            ("conv4_1", 256, 512), ("conv4_2", 512, 512), ("conv4_3", 512, 512),
            ("conv5_1", 512, 512), ("conv5_2", 512, 512), ("conv5_3", 512, 512)]
STYLE_LAYERS = ["conv1_2", "conv2_2", "conv3_3", "conv4_3"]
CONTENT_LAYERS = ["conv3_3"]


def synthetic_vgg_npz(path, seed=7):
    """faststyle_b200.synth.synthetic_vgg_weights(7) (He-normal kernels, zero biases) in the npz layout the
    reference's loader expects; conv5_x get their own draws AFTER conv4_3 so conv1_1..conv4_3 match the suite."""
    rng = np.random.RandomState(seed)
    out = {}
    for name, cin, cout in VGG_CONV:
        std = np.sqrt(2.0 / (9.0 * cin))
        out[name + "_W"] = (rng.standard_normal((3, 3, cin, cout)) * std).astype(np.float32)
        out[name + "_b"] = np.zeros((cout,), np.float32)
    np.savez(path, **out)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--vgg", default="libs/vgg16_weights.npz")
    args = ap.parse_args()
    import tensorflow as tf
    if hasattr(tf, "compat") and hasattr(tf.compat, "v1") and not hasattr(tf, "placeholder"):
        tf = tf.compat.v1
        tf.disable_v2_behavior()
    import losses
    import utils
    from im_transf_net import create_net
    from libs import vgg16

    if not os.path.isdir(args.out):
        os.makedirs(args.out)
    vgg_path = args.vgg
    real_vgg = os.path.exists(vgg_path)
    if not real_vgg:
        vgg_path = synthetic_vgg_npz(os.path.join(args.out, "_synthetic_vgg16_weights.npz"))
    print("TensorFlow", tf.__version__, "| VGG weights:", vgg_path, "(real)" if real_vgg else "(synthetic, seed 7)")
    meta = dict(tf_version=str(tf.__version__), real_vgg=bool(real_vgg))

    # ---------------------------------------------------------------- transform forward (stylize_image.py:63-75)
    x = np.random.RandomState(0).randint(0, 256, (2, 256, 256, 3)).astype(np.float32)
    tf.reset_default_graph()
    with tf.variable_scope('img_t_net'):
        X = tf.placeholder(tf.float32, shape=x.shape, name='input')
        Y = create_net(X, 'resize')
    with tf.Session() as sess:
        tf.train.Saver().restore(sess, 'models/starry_final.ckpt')
        y = sess.run(Y, feed_dict={X: x})
    np.savez_compressed(os.path.join(args.out, "transform_fwd.npz"), Y=y, **meta)

    # ---------------------------------------------------------------- style target Grams (train.py:135-151)
    style_img = utils.imread('style_images/starry_night_crop.jpg')[np.newaxis, :].astype(np.float32)
    style_names = ['vgg/' + i + ':0' for i in STYLE_LAYERS]
    content_names = ['vgg/' + i + ':0' for i in CONTENT_LAYERS]
    tf.reset_default_graph()
    with tf.variable_scope('vgg'):
        X_vgg = tf.placeholder(tf.float32, shape=style_img.shape, name='input')
        vggnet = vgg16.vgg16(X_vgg)
    with tf.Session() as sess:
        vggnet.load_weights(vgg_path, sess)
        target_grams = sess.run(utils.get_grams(style_names), feed_dict={X_vgg: style_img})
    np.savez_compressed(os.path.join(args.out, "style_grams.npz"),
                        **dict(meta, **{n: g for n, g in zip(STYLE_LAYERS, target_grams)}))

    # ---------------------------------------------------------------- train step (train.py:158-204, 245-275)
    xb = np.random.RandomState(3).randint(0, 256, (2, 128, 128, 3)).astype(np.float32)
    beta_val = 1e-4
    tf.reset_default_graph()
    with tf.variable_scope('img_t_net'):
        X = tf.placeholder(tf.float32, shape=xb.shape, name='input')
        Y = create_net(X, 'resize')
    with tf.variable_scope('vgg'):
        vggnet = vgg16.vgg16(Y)
    input_img_grams = utils.get_grams(style_names)
    content_layers = utils.get_layers(content_names)
    content_targets = tuple(tf.placeholder(tf.float32, shape=l.get_shape()) for l in content_layers)
    cont_loss = losses.content_loss(content_layers, content_targets, [1.0])
    style_loss = losses.style_loss(input_img_grams, target_grams, [5.0] * 4)
    tv_loss = losses.tv_loss(Y)
    beta = tf.placeholder(tf.float32, shape=[])
    loss = cont_loss + style_loss + beta * tv_loss
    train_vars = tf.get_collection(tf.GraphKeys.TRAINABLE_VARIABLES, scope='img_t_net')
    grads = tf.gradients(loss, train_vars)
    global_step = tf.Variable(0, name='global_step', trainable=False)
    optimizer = tf.train.AdamOptimizer(1e-3).minimize(loss, global_step, train_vars)
    out = dict(meta)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        vggnet.load_weights(vgg_path, sess)
        tf.train.Saver(train_vars).restore(sess, 'models/starry_final.ckpt')
        content_data = sess.run(content_layers, feed_dict={Y: xb})            # train.py:250-251
        fd = {X: xb, beta: beta_val}
        fd.update({p: d for p, d in zip(content_targets, content_data)})
        vals = sess.run([loss, cont_loss, style_loss, tv_loss, Y] + list(input_img_grams) + list(grads), feed_dict=fd)
        out.update(loss=vals[0], content=vals[1], style=vals[2], tv=vals[3], Y=vals[4], beta=beta_val)
        for n, g in zip(STYLE_LAYERS, vals[5:5 + len(STYLE_LAYERS)]):
            out["gram/" + n] = g
        for v, g in zip(train_vars, vals[5 + len(STYLE_LAYERS):]):
            out["grad/" + v.name.replace(":0", "")] = g
        sess.run(optimizer, feed_dict=fd)
        for v, val in zip(train_vars, sess.run(train_vars)):
            out["after_adam/" + v.name.replace(":0", "")] = val
    np.savez_compressed(os.path.join(args.out, "train_step.npz"), **out)

    # ---------------------------------------------------------------- bicubic resize (datapipe.py:14-26)
    img = utils.imread('results/chicago.jpg')
    tf.reset_default_graph()
    I = tf.placeholder(tf.uint8, shape=img.shape)
    R = tf.image.resize_images(I, size=[256, 256], method=2)
    with tf.Session() as sess:
        r = sess.run(R, feed_dict={I: img})
    np.savez_compressed(os.path.join(args.out, "bicubic.npz"), resized=r, **meta)
    print("wrote", sorted(os.listdir(args.out)))


if __name__ == '__main__':
    main()
