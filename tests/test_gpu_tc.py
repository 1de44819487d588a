"""tcgen05 3x3 convolution (split-bf16 x3) against the fp64 oracle conv.
Tolerance: 2e-5 relative to the output's max magnitude - bf16x3 drops only the lo*lo term
(2^-16 relative per product, averaged over K >= 576)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restate as R  # noqa: E402


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _relerr(got, want):
    want = want.double()
    return float((got.double().cpu() - want).abs().max() / max(want.abs().max().item(), 1e-30))


CASES = [
    # N, H, W, C, OC, same, relu
    (1, 16, 16, 64, 64, 1, 0),       # single full tile
    (2, 20, 18, 64, 64, 0, 0),       # residual-block VALID, ragged tiles
    (1, 19, 23, 64, 128, 1, 1),      # VGG SAME + bias + relu, odd sizes
    (2, 33, 40, 128, 256, 1, 1),     # two N tiles, two channel blocks, TH=16 path
    (1, 8, 8, 256, 512, 1, 1),       # TH=8 path, deep K
    (3, 64, 64, 64, 64, 1, 1),       # > 1 tile per CTA in flight (double-buffered TMEM)
    (1, 24, 40, 64, 128, 1, 1),      # 3 x 3 = 9 spatial tiles (8-row): the last CTA pair has a dummy peer tile
    (5, 40, 48, 128, 128, 0, 0),     # 45 16-row tiles: odd count, several waves of pairs
    (8, 64, 64, 256, 256, 1, 1),     # BASELINE config 3/4 shape (conv3_2 at per-GPU batch 8): multi-wave, 4 K blocks
]


@pytest.fixture(params=[1, 0], ids=["cta_pair", "single_cta"])
def tc_pair(request):
    """Both launch forms of the tcgen05 convolution: CTA pairs (cta_group::2, the default) and single CTAs."""
    from faststyle_b200 import _lib
    _lib.call("fs_set_tc_pair", request.param)
    yield request.param
    _lib.call("fs_set_tc_pair", 1)


@pytest.mark.parametrize("case", CASES)
def test_conv3x3_tc_forward_and_dgrad(built_lib, case, tc_pair):
    from faststyle_b200 import _lib
    lib = _lib.load()
    N, H, W, Ci, Co, same, relu = case
    rng = np.random.RandomState(abs(hash(case)) % 2**31)
    x = rng.standard_normal((N, H, W, Ci)).astype(np.float32)
    w = (rng.standard_normal((3, 3, Ci, Co)) * 0.05).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    xt = R.nhwc_to_nchw(torch.from_numpy(x).double()).requires_grad_(True)
    wt = torch.from_numpy(w).double()
    yo = R.conv2d_tf(xt, wt, 1, "SAME" if same else "VALID")
    ypre = yo
    if relu:
        yo = torch.relu(yo + torch.from_numpy(b).double().view(1, -1, 1, 1))
    OH, OW = yo.shape[2], yo.shape[3]
    xd, wd, bd = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), torch.from_numpy(b).cuda()
    nb = lib.fs_conv3x3_tc_scratch_bytes(N, H, W, Ci, Co)
    scratch = torch.empty(nb + 1024, dtype=torch.uint8, device="cuda")
    off = (-scratch.data_ptr()) % 1024
    sp = C.c_void_p(scratch.data_ptr() + off)
    y = torch.full((N, OH, OW, Co), float("nan"), device="cuda")
    _lib.call("fs_conv3x3_tc_forward", _ptr(xd), _ptr(wd), _ptr(bd) if relu else None, _ptr(y), sp,
              C.c_size_t(nb), N, H, W, Ci, Co, same, relu, _st())
    torch.cuda.synchronize()
    assert _relerr(y, R.nchw_to_nhwc(yo.detach())) < 2e-5
    # data gradient of the plain conv
    dy = rng.standard_normal((N, OH, OW, Co)).astype(np.float32)
    ypre.backward(R.nhwc_to_nchw(torch.from_numpy(dy).double()))
    dyd = torch.from_numpy(dy).cuda()
    dx = torch.full((N, H, W, Ci), float("nan"), device="cuda")
    _lib.call("fs_conv3x3_tc_dgrad", _ptr(dyd), _ptr(wd), _ptr(dx), sp, C.c_size_t(nb), N, H, W, Ci, Co,
              same, _st())
    torch.cuda.synchronize()
    assert _relerr(dx, R.nchw_to_nhwc(xt.grad)) < 2e-5


@pytest.mark.parametrize("case", [(1, 16, 16, 1), (2, 20, 18, 0), (3, 41, 37, 0), (2, 33, 40, 1)])
def test_wgrad3x3_tc(built_lib, case):
    """tcgen05 weight gradient (MN-major operands) of a 64->64 3x3 conv vs oracle autograd (fp64)."""
    from faststyle_b200 import _lib
    lib = _lib.load()
    N, H, W, same = case
    rng = np.random.RandomState(17 + H)
    x = rng.standard_normal((N, H, W, 64)).astype(np.float32)
    wt = torch.from_numpy((rng.standard_normal((3, 3, 64, 64)) * 0.05)).double().requires_grad_(True)
    yo = R.conv2d_tf(R.nhwc_to_nchw(torch.from_numpy(x).double()), wt, 1, "SAME" if same else "VALID")
    OH, OW = yo.shape[2], yo.shape[3]
    dy = rng.standard_normal((N, OH, OW, 64)).astype(np.float32)
    yo.backward(R.nhwc_to_nchw(torch.from_numpy(dy).double()))
    xd, dyd = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()
    nb = lib.fs_wgrad3x3_tc_scratch_bytes(N, H, W)
    scratch = torch.empty(nb + 1024, dtype=torch.uint8, device="cuda")
    sp = C.c_void_p(scratch.data_ptr() + (-scratch.data_ptr()) % 1024)
    dw = torch.full((3, 3, 64, 64), float("nan"), device="cuda")
    _lib.call("fs_wgrad3x3_tc", _ptr(xd), _ptr(dyd), _ptr(dw), sp, C.c_size_t(nb), N, H, W, same, _st())
    torch.cuda.synchronize()
    err = _relerr(dw, wt.grad)
    print("wgrad3x3_tc", case, "rel err", err)
    assert err < 2e-5


GRAM_CASES = [
    # N, H, W, C   (utils.py:66-83: G = F^T F / (h*w*c) per sample)
    (2, 5, 7, 64),          # fewer pixels than one chunk (TMA zero fill)
    (1, 33, 29, 128),       # ragged last chunk
    (2, 17, 16, 256),       # two row tiles
    (2, 9, 13, 512),        # four row tiles, two N halves (all 512 TMEM columns)
    (3, 64, 64, 64),        # many chunks per split
    (8, 256, 256, 64),      # BASELINE config 3/4 shapes, per-GPU batch 8: conv1_2
    (8, 128, 128, 128),     # conv2_2
    (8, 64, 64, 256),       # conv3_3
    (8, 32, 32, 512),       # conv4_3
    (1, 128, 128, 512),     # slow_style 1024x1024 (config 5): conv4_3
]


@pytest.mark.parametrize("case", GRAM_CASES)
def test_gram_tc(built_lib, case):
    """tcgen05 Gram kernel (split-bf16 x3, MN-major operands) vs an fp64 matmul of the same features, plus
    size-independent properties: symmetry and trace(G) = sum(F^2)/(hwc).  Tolerance 2e-5 of max|G|."""
    from faststyle_b200 import _lib
    lib = _lib.load()
    N, H, W, Cc = case
    g = torch.Generator().manual_seed(1000 + H + Cc)
    f = torch.rand((N, H, W, Cc), generator=g) * 3.0          # post-ReLU-like, non-negative
    f[:, ::3] = 0.0
    fd = f.cuda()
    nb = lib.fs_gram_tc_scratch_bytes(N, H, W, Cc)
    scratch = torch.empty(nb + 1024, dtype=torch.uint8, device="cuda")
    sp = C.c_void_p(scratch.data_ptr() + (-scratch.data_ptr()) % 1024)
    G = torch.full((N, Cc, Cc), float("nan"), device="cuda")
    _lib.call("fs_gram_tc_forward", _ptr(fd), _ptr(G), sp, C.c_size_t(nb), N, H, W, Cc, _st())
    torch.cuda.synchronize()
    F2 = f.double().reshape(N, H * W, Cc)
    want = torch.matmul(F2.transpose(1, 2), F2) / (H * W * Cc)
    err = _relerr(G, want)
    Gc = G.double().cpu()
    sym = float((Gc - Gc.transpose(1, 2)).abs().max() / want.abs().max())
    trw = (F2 ** 2).sum((1, 2)) / (H * W * Cc)
    tr = float(((torch.diagonal(Gc, dim1=1, dim2=2).sum(1) - trw).abs() / trw).max())
    print("gram_tc", case, "rel err %.3g asym %.3g trace %.3g" % (err, sym, tr))
    assert err < 2e-5 and sym < 2e-5 and tr < 2e-5
