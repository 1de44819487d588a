#!/usr/bin/env python
"""Train a transform network against a style image with the VGG16 perceptual loss on B200s
(drop-in for the reference's train.py: same flags and defaults, same checkpoint files).

Single GPU:   python train.py --train_dir <dir> --model_name starry
Multi GPU:    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
                     train.py --train_dir <dir> --model_name starry --batch_size 64
``--batch_size`` is the GLOBAL batch; it is split evenly across ranks.
``--train_dir`` may hold the reference's TFRecord shards (train-*), plain image files, or be
the literal ``synthetic[:N]``.  VGG weights are read from libs/vgg16_weights.npz (override
with $VGG16_WEIGHTS; ``synthetic`` selects seeded synthetic weights).

Files written: ``models/<name>_final.ckpt`` holds exactly the reference's 48 variables (train.py:225,286);
the periodic ``training/<name>.ckpt-<step>`` files hold those plus the Adam slots and an int32 ``global_step``
under TF's names, but NOT the frozen ``vgg/*`` variables the reference's all-variables Saver also dumps
(train.py:224,259) - no code path restores a training checkpoint.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def setup_parser():
    p = argparse.ArgumentParser(description='Train a style transfer net.')
    p.add_argument('--train_dir', help='Training data: TFRecord shards, image files, or "synthetic[:N]".')
    p.add_argument('--model_name', help='Name of the model being trained.')
    p.add_argument('--style_img_path', default='./style_images/starry_night_crop.jpg', help='Style target image.')
    p.add_argument('--learn_rate', default=1e-3, type=float, help='Adam learning rate.')
    p.add_argument('--batch_size', default=4, type=int, help='(Global) batch size.')
    p.add_argument('--n_epochs', default=2, type=int, help='Number of passes over the data.')
    p.add_argument('--preprocess_size', default=[256, 256], nargs=2, type=int,
                   help='Training images are resized to this H W.')
    p.add_argument('--run_name', default=None, help='TensorBoard run directory under ./summaries/train.')
    p.add_argument('--loss_content_layers', nargs='*', default=['conv3_3'], help='VGG layers of the content loss.')
    p.add_argument('--loss_style_layers', nargs='*', default=['conv1_2', 'conv2_2', 'conv3_3', 'conv4_3'],
                   help='VGG layers of the style loss.')
    p.add_argument('--content_weights', nargs='*', default=[1.0], type=float, help='Content loss weights.')
    p.add_argument('--style_weights', nargs='*', default=[5.0, 5.0, 5.0, 5.0], type=float, help='Style loss weights.')
    p.add_argument('--num_steps_ckpt', default=1000, type=int, help='Checkpoint cadence in steps.')
    p.add_argument('--num_pipe_buffer', default=4000, type=int, help='Shuffle buffer size (images).')
    p.add_argument('--num_steps_break', default=-1, type=int, help='Stop after this step (-1: run all epochs).')
    p.add_argument('--beta', default=0.0, type=float, help='Total-variation weight.')
    p.add_argument('--style_target_resize', default=1.0, type=float, help='Scale factor for the style image.')
    p.add_argument('--upsample_method', choices=['deconv', 'resize'], default='resize', help='Upsampling variant.')
    return p


def load_vgg_weights():
    path = os.environ.get('VGG16_WEIGHTS', 'libs/vgg16_weights.npz')
    if path == 'synthetic':
        from faststyle_b200 import synth
        print('Using seeded SYNTHETIC VGG16 weights (no perceptual meaning).')
        return synth.synthetic_vgg_weights(7)
    if not os.path.exists(path):
        raise SystemExit("VGG16 weights %r not found. Fetch vgg16_weights.npz as the reference's "
                         "libs/get_vgg16_weights.sh does, or set VGG16_WEIGHTS=synthetic." % path)
    w = np.load(path)
    return {k: w[k] for k in w.keys() if 'fc' not in k}


def main(args):
    import torch
    from faststyle_b200 import datapipe, synth, utils
    from faststyle_b200 import variables as V
    from faststyle_b200.layout import TRANSFORM_VARS
    from faststyle_b200.summary import FileWriter
    from faststyle_b200.trainer import Trainer

    if not args.train_dir or not args.model_name:
        raise SystemExit("--train_dir and --model_name are required")

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    pg = vote_pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        pg = dist.group.WORLD
        # host-side group for the per-step end-of-data vote (datapipe.next_batch_collective): CPU tensors, so the
        # vote never touches the GPU queue
        vote_pg = dist.new_group(backend='gloo')
    if args.batch_size % world:
        raise SystemExit("--batch_size %d is not divisible by the %d ranks" % (args.batch_size, world))
    local_batch = args.batch_size // world

    style_img = utils.imresize(utils.imread(args.style_img_path), args.style_target_resize)
    style_img = style_img[np.newaxis, :].astype(np.float32)

    # transform-net variables with the reference's initialisers (identical on every rank)
    params = synth.init_transform_params(seed=1, upsample_method=args.upsample_method)
    if rank == 0:
        print('Precomputing target style layers.')
    trainer = Trainer(params, load_vgg_weights(), style_img, local_batch, args.preprocess_size,
                      args.loss_content_layers, args.loss_style_layers, args.content_weights,
                      args.style_weights, args.beta, args.learn_rate,
                      device='cuda:%d' % local_rank, process_group=pg, upsample_method=args.upsample_method)

    gpu_prep = os.environ.get('FS_GPU_PREPROCESS', '0') == '1'      # resize + float cast on the device (same arithmetic)
    batches = datapipe.batcher(args.train_dir, local_batch, args.preprocess_size, args.n_epochs,
                               max(args.num_pipe_buffer // world, local_batch), seed=1234, shard=(rank, world),
                               raw=gpu_prep)
    prep = datapipe.GpuPreprocessor(local_batch, args.preprocess_size, 'cuda:%d' % local_rank) if gpu_prep else None
    # the reference feeds through queue-runner threads with room for 3 batches beyond the shuffle buffer
    # (train.py:241-242, datapipe.py:72-77): decode / resize / shuffle run beside the training loop here too
    batches = datapipe.prefetch(batches, 3)

    run_name = args.run_name
    writer = None
    if rank == 0:
        os.makedirs('./summaries/train/', exist_ok=True)
        if run_name is None:
            existing = [d for d in os.listdir('./summaries/train/') if os.path.isdir('./summaries/train/' + d)]
            count = 0
            while args.model_name + str(count) in existing:
                count += 1
            run_name = args.model_name + str(count)
        os.makedirs('./training', exist_ok=True)
        os.makedirs('./models', exist_ok=True)
        writer = FileWriter('./summaries/train/' + run_name)

    def save(prefix, with_slots):
        V.reset_default_graph()
        for k, v in trainer.variables().items():
            V.set_variable(k, v)
        if with_slots:
            for k, v in trainer.optimizer_slots().items():
                V.set_variable(k, v)
            V.set_variable('global_step', np.array(trainer.global_step, np.int32))     # tf.Variable(0): int32
        return V.Saver().save(None, prefix)

    if rank == 0:
        print('Starting training...')
    try:
        while True:
            current_step = trainer.global_step
            batch = datapipe.next_batch_collective(batches, vote_pg)     # all ranks stop at the same step
            if prep is not None:
                batch = prep(batch)
            want_log = (current_step % args.num_steps_ckpt == 0) or (current_step % 10 == 0)
            if rank == 0 and current_step % args.num_steps_ckpt == 0:
                save('training/%s.ckpt-%d' % (args.model_name, current_step), with_slots=True)
            out = trainer.step(batch, fetch_losses=want_log)
            if want_log and rank == 0:
                content, style, tv, loss = [float(v) for v in out]
                writer.add_scalars(current_step, {'summaries/loss': loss, 'summaries/style_loss': style,
                                                  'summaries/content_loss': content, 'summaries/tv_loss': tv})
                print(current_step, loss)
            if current_step == args.num_steps_break:
                if rank == 0:
                    print('Done training.')
                break
    except (datapipe.OutOfRangeError, StopIteration):
        if rank == 0:
            print('Done training.')
    finally:
        # always leave the trained transform network behind (train.py:283-288)
        batches.close()
        if rank == 0:
            save('models/%s_final.ckpt' % args.model_name, with_slots=False)
            writer.close()
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    assert len(TRANSFORM_VARS) == 48


if __name__ == '__main__':
    main(setup_parser().parse_args())
