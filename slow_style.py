#!/usr/bin/env python
"""Gatys-style optimisation of an image through the VGG16 perceptual loss on a B200
(drop-in for the reference's slow_style.py: same flags and defaults).  The pixel tensor is
the trainable variable; TF-Adam updates it directly."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def setup_parser():
    p = argparse.ArgumentParser(description='Optimise an image towards a style (Gatys et al.).')
    p.add_argument('--style_img_path', help='Style template image.')
    p.add_argument('--cont_img_path', help='Content template image.')
    p.add_argument('--learn_rate', default=1e1, type=float, help='Adam learning rate.')
    p.add_argument('--loss_content_layers', nargs='*', default=['conv3_3'], help='VGG layers of the content loss.')
    p.add_argument('--loss_style_layers', nargs='*', default=['conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'],
                   help='VGG layers of the style loss.')
    p.add_argument('--content_weights', nargs='*', default=[1.0], type=float, help='Content loss weights.')
    p.add_argument('--style_weights', nargs='*', default=[5.0, 5.0, 5.0, 5.0], type=float, help='Style loss weights.')
    p.add_argument('--num_steps_break', default=500, type=int, help='Number of optimiser steps.')
    p.add_argument('--beta', default=1.e-4, type=float, help='Total-variation weight.')
    p.add_argument('--style_target_resize', default=1.0, type=float, help='Scale factor for the style image.')
    p.add_argument('--cont_target_resize', default=1.0, type=float, help='Scale factor for the content image.')
    p.add_argument('--output_img_path', default='./out.jpg', help='Where to write the result.')
    return p


def main(args, seed=None):
    import torch
    from faststyle_b200 import utils
    from faststyle_b200.engine import Engine, TFAdam, make_loss_config, pack_vgg
    from train import load_vgg_weights

    style_img = utils.imresize(utils.imread(args.style_img_path), args.style_target_resize)
    style_img = style_img[np.newaxis, :].astype(np.float32)
    cont_img = utils.imresize(utils.imread(args.cont_img_path), args.cont_target_resize)
    cont_img = cont_img[np.newaxis, :].astype(np.float32)

    dev = torch.device('cuda', torch.cuda.current_device())
    packed = pack_vgg(load_vgg_weights(), dev)
    cfg = make_loss_config(args.loss_content_layers, args.content_weights, args.loss_style_layers,
                           args.style_weights, args.beta)
    print('Precomputing target style layers.')
    seng = Engine(1, style_img.shape[1], style_img.shape[2], vgg=True, style_layers=args.loss_style_layers, device=dev)
    target_grams = seng.vgg_grams(packed, style_img, args.loss_style_layers)
    torch.cuda.synchronize()
    del seng

    _, H, W, _ = cont_img.shape
    eng = Engine(1, H, W, vgg_bwd=True, content_layers=args.loss_content_layers,
                 style_layers=args.loss_style_layers, device=dev)
    print('Precomputing target content layers.')
    eng.set_content_targets(packed, cont_img, cfg)

    # white-noise start, U[0,255) (slow_style.py:116-122; the reference does not seed it)
    rng = np.random.RandomState(seed)
    X = torch.from_numpy((rng.rand(1, H, W, 3) * 255.0).astype(np.float32)).to(dev)
    opt = TFAdam(X.view(-1), args.learn_rate)

    # The reference reads global_step BEFORE each update (slow_style.py:160-176), so
    # num_steps_break + 1 updates run; reproduced here.
    global_step = 0
    current_step = 0
    while current_step < args.num_steps_break:
        current_step = global_step
        losses, grad = eng.perceptual_loss(packed, X, cfg, target_grams, need_grad=True)
        opt.step(grad.view(-1))
        global_step += 1
        if current_step % 10 == 0:
            print(current_step, float(losses[3]))
    img_out = np.squeeze(X.cpu().numpy())
    utils.imwrite(args.output_img_path, img_out)
    return img_out


if __name__ == '__main__':
    main(setup_parser().parse_args())
