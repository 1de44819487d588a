#!/usr/bin/env python
"""Stylize a video feed frame by frame on a B200 (drop-in for the reference's stylize_webcam.py:
same flags, same .ckpt format, same frame conventions, same output.avi).

    python stylize_webcam.py --model_path models/starry_final.ckpt --resolution 640 480

The reference opens camera 0, shows every stylised frame in a window and appends it to
./output.avi (XVID, 15 fps) until 'q' is pressed (stylize_webcam.py:52-108).  Extra flags for
machines without a camera or a display (none of them changes the frame arithmetic):
``--source`` (camera index, or a video file / image-sequence pattern OpenCV can open),
``--output``, ``--no_display``, ``--max_frames``.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def setup_parser():
    p = argparse.ArgumentParser(description="Use a trained fast style transfer model to filter webcam feed.")
    p.add_argument('--model_path', default='./models/starry_final.ckpt', help='Path to .ckpt for the trained model.')
    p.add_argument('--upsample_method', choices=['resize', 'deconv'], default='resize',
                   help='The upsample method that was used to construct the model being loaded.')
    p.add_argument('--resolution', nargs=2, type=int, default=None,
                   help='Dimensions for webcam (x y). Leave blank to use the default resolution.')
    # --- additions (not in the reference)
    p.add_argument('--source', default='0', help='Camera index (default 0) or a video file path.')
    p.add_argument('--output', default='output.avi', help='Where to write the stylised video.')
    p.add_argument('--no_display', action='store_true', help='Do not open a window (headless boxes).')
    p.add_argument('--max_frames', type=int, default=0, help='Stop after this many frames (0 = until q / end of feed).')
    return p


def open_capture(source, resolution):
    import cv2
    cap = cv2.VideoCapture(int(source) if str(source).isdigit() else source)
    if not cap.isOpened():
        raise SystemExit("cannot open video source %r" % (source,))
    if resolution is not None:
        x_length, y_length = resolution
        cap.set(3, x_length)      # 3 and 4 are OpenCV property IDs (stylize_webcam.py:58-59)
        cap.set(4, y_length)
    return cap, int(cap.get(3)), int(cap.get(4))


def main(args):
    import cv2
    from faststyle_b200 import variables as V
    from faststyle_b200.layout import transform_vars
    from faststyle_b200.stream import FrameStylizer

    cap, x_new, y_new = open_capture(args.source, args.resolution)
    print('Resolution is: {0} by {1}'.format(x_new, y_new))

    # "Create the graph" + restore (stylize_webcam.py:62-78): every variable must be in the checkpoint
    V.reset_default_graph()
    print('Loading up model...')
    V.Saver().restore(None, args.model_path)
    params = {}
    with V.variable_scope('img_t_net'):
        for full, shape in transform_vars(args.upsample_method):
            params[full] = V.get_variable(full[len('img_t_net/'):], shape)
    fs = FrameStylizer(params, y_new, x_new, args.upsample_method, swap_rb=True)

    fourcc = cv2.VideoWriter_fourcc(*'XVID')
    out = cv2.VideoWriter(args.output, fourcc, 15.0, (x_new, y_new))
    print('Begin filtering...')
    n, t0 = 0, time.time()
    try:
        while True:
            ret, frame = cap.read()
            if not ret or frame is None:
                break
            # the BGR frame is fed as-is and channels are swapped afterwards (stylize_webcam.py:82-90)
            img_out = fs.stylize(np.ascontiguousarray(frame))
            # the writer only accepts frames of the size it was opened with; the net's output is
            # (OH, OW) != (y, x) unless (y+80) % 4 == 0 - the reference silently drops
            # such frames inside cv2; keep them by cropping/padding to the declared size
            if img_out.shape[:2] != (y_new, x_new):
                fixed = np.zeros((y_new, x_new, 3), np.uint8)
                h, w = min(y_new, img_out.shape[0]), min(x_new, img_out.shape[1])
                fixed[:h, :w] = img_out[:h, :w]
                img_out = fixed
            out.write(img_out)
            n += 1
            if not args.no_display:
                cv2.imshow('frame', img_out)
                if cv2.waitKey(1) & 0xFF == ord('q'):
                    break
            if args.max_frames and n >= args.max_frames:
                break
    finally:
        cap.release()
        out.release()
        if not args.no_display:
            cv2.destroyAllWindows()
    dt = time.time() - t0
    print('Stylised %d frames in %.2f s (%.1f fps, CUDA graph: %s).' % (n, dt, n / max(dt, 1e-9), fs.graph is not None))


if __name__ == '__main__':
    main(setup_parser().parse_args())
