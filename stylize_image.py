#!/usr/bin/env python
"""Stylize one image with a trained transform network on a B200 (drop-in for the
reference's stylize_image.py: same flags, same .ckpt format, same output convention).

    python stylize_image.py --input_img_path in.jpg --output_img_path out.jpg \
                            --model_path models/starry_final.ckpt
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def setup_parser():
    p = argparse.ArgumentParser(description="Apply a trained fast-style-transfer model to an image.")
    p.add_argument('--input_img_path', help='Content image to stylize.')
    p.add_argument('--output_img_path', default='./results/styled.jpg', help='Where to write the result.')
    p.add_argument('--model_path', default='./models/starry_final.ckpt',
                   help='Checkpoint prefix (<prefix>.index / .data-00000-of-00001).')
    p.add_argument('--content_target_resize', default=1.0, type=float,
                   help='Scale factor applied to the input before stylizing.')
    p.add_argument('--upsample_method', choices=['resize', 'deconv'], default='resize',
                   help='Upsampling variant the model was trained with.')
    return p


def main(args):
    from faststyle_b200 import utils
    from faststyle_b200 import variables as V
    from faststyle_b200.im_transf_net import create_net

    if not args.input_img_path:
        raise SystemExit("--input_img_path is required")
    img = utils.imread(args.input_img_path)
    img = utils.imresize(img, args.content_target_resize)
    img_4d = img[np.newaxis, :]

    V.reset_default_graph()
    saver = V.Saver()
    print('Loading up model...')
    saver.restore(None, args.model_path)
    print('Evaluating...')
    with V.variable_scope('img_t_net'):
        Y = create_net(img_4d, args.upsample_method)
    img_out = np.squeeze(Y.cpu().numpy())

    print('Saving image.')
    out_dir = os.path.dirname(args.output_img_path)
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
    utils.imwrite(args.output_img_path, img_out)
    print('Done.')


if __name__ == '__main__':
    main(setup_parser().parse_args())
