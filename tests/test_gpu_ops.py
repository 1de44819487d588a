"""Op-level parity: each C-ABI kernel against the oracle restatement (fp64) on the
same seeded inputs.  Tolerances are relative to the output's max magnitude and are
fp32-reorder-noise sized (the CUDA path computes in fp32 FFMA; the oracle in fp64)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import restate as R  # noqa: E402


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _relerr(got, want):
    want = want.double()
    return float((got.double().cpu() - want).abs().max() / max(want.abs().max().item(), 1e-30))


CONV_CASES = [
    # N, H, W, C, K, OC, stride, same, relu/bias
    (2, 37, 41, 4, 9, 16, 1, 1, 0),      # initconv_0 shape class (Cin padded 3->4)
    (2, 33, 30, 16, 3, 32, 2, 1, 0),     # stride 2 SAME, odd/even sizes (pad 1,1 / 0,1)
    (1, 21, 21, 32, 3, 64, 2, 1, 0),
    (2, 20, 18, 64, 3, 64, 1, 0, 0),     # residual VALID
    (1, 19, 23, 64, 3, 128, 1, 1, 1),    # VGG SAME + bias + relu, OC=128 tile
    (1, 16, 16, 16, 9, 4, 1, 1, 0),      # upsample_2 shape class (Cout padded 3->4)
    (1, 12, 12, 256, 3, 512, 1, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward_dgrad_wgrad(built_lib, case):
    from faststyle_b200 import _lib
    N, H, W, Ci, K, Co, s, same, br = case
    rng = np.random.RandomState(hash(case) % 2**31)
    x = rng.standard_normal((N, H, W, Ci)).astype(np.float32)
    w = (rng.standard_normal((K, K, Ci, Co)) * 0.1).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32) if br else None
    pad = "SAME" if same else "VALID"
    xt = R.nhwc_to_nchw(torch.from_numpy(x).double()).requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    yo = R.conv2d_tf(xt, wt, s, pad)
    if br:
        yo = torch.relu(yo + torch.from_numpy(b).double().view(1, -1, 1, 1))
    OH, OW = yo.shape[2], yo.shape[3]
    dy = rng.standard_normal((N, OH, OW, Co)).astype(np.float32)
    yo.backward(R.nhwc_to_nchw(torch.from_numpy(dy).double()))

    xd, wd = _dev(x), _dev(w)
    y = torch.empty((N, OH, OW, Co), device="cuda")
    bd = _dev(b) if br else None
    _lib.call("fs_conv2d_forward", _ptr(xd), _ptr(wd), _ptr(bd) if br else None, _ptr(y),
              N, H, W, Ci, K, K, Co, s, same, br, _st())
    assert _relerr(y, R.nchw_to_nhwc(yo.detach())) < 2e-6
    if br:
        return      # gradients of the fused bias+relu form are covered by the VGG backward test
    dyd = _dev(dy)
    dx = torch.empty_like(xd)
    scratch = torch.empty(K * K * Ci * Co, device="cuda")
    _lib.call("fs_conv2d_dgrad", _ptr(dyd), _ptr(wd), _ptr(dx), _ptr(scratch), N, H, W, Ci, K, K, Co,
              s, same, _st())
    assert _relerr(dx, R.nchw_to_nhwc(xt.grad)) < 2e-6
    nsc = _lib.load().fs_conv2d_wgrad_scratch_floats(Ci, K, K, Co)
    sc = torch.empty(nsc, device="cuda")
    dw = torch.empty_like(wd)
    _lib.call("fs_conv2d_wgrad", _ptr(xd), _ptr(dyd), _ptr(dw), _ptr(sc), C.c_longlong(nsc), N, H, W, Ci,
              K, K, Co, s, same, _st())
    assert _relerr(dw, wt.grad) < 2e-6


@pytest.mark.parametrize("shape", [(2, 9, 11, 64, 32), (1, 16, 16, 32, 16)])
def test_upconv_forward(built_lib, shape):
    """Fused resize-conv == NN-resize x4 + 3x3 s2 SAME conv (im_transf_net.py:122-155)."""
    from faststyle_b200 import _lib
    N, H, W, Ci, Co = shape
    rng = np.random.RandomState(3)
    x = rng.standard_normal((N, H, W, Ci)).astype(np.float32)
    w = rng.standard_normal((3, 3, Ci, Co)).astype(np.float32)
    yo = R.upconv2d(R.nhwc_to_nchw(torch.from_numpy(x).double()), torch.from_numpy(w).double(), 2)
    y = torch.empty((N, 2 * H, 2 * W, Co), device="cuda")
    sc = torch.empty(16 * Ci * Co, device="cuda")
    xd, wd = _dev(x), _dev(w)          # keep the device tensors alive across the async launch
    _lib.call("fs_upconv2d_forward", _ptr(xd), _ptr(wd), _ptr(y), _ptr(sc), N, H, W, Ci, Co, _st())
    assert tuple(yo.shape[2:]) == (2 * H, 2 * W)
    assert _relerr(y, R.nchw_to_nhwc(yo)) < 2e-6


@pytest.mark.parametrize("C_,act", [(16, 1), (64, 0), (4, 2), (32, 1)])
def test_instnorm_forward(built_lib, C_, act):
    from faststyle_b200 import _lib
    N, H, W = 3, 29, 31
    rng = np.random.RandomState(5)
    x = (rng.standard_normal((N, H, W, C_)) * 3 + 50).astype(np.float32)   # large mean: cancellation stress
    g = rng.standard_normal(C_).astype(np.float32)
    b = rng.standard_normal(C_).astype(np.float32)
    yo = R.inst_norm(R.nhwc_to_nchw(torch.from_numpy(x).double()), torch.from_numpy(g).double(),
                     torch.from_numpy(b).double())
    yo = torch.relu(yo) if act == 1 else (R.scaled_tanh(yo) if act == 2 else yo)
    y = torch.empty((N, H, W, C_), device="cuda")
    stats = torch.empty(2 * N * C_, device="cuda")
    sc = torch.empty(N * 64 * 2 * C_, dtype=torch.float64, device="cuda")
    xd, gd, bd = _dev(x), _dev(g), _dev(b)
    _lib.call("fs_instnorm_forward", _ptr(xd), _ptr(gd), _ptr(bd), _ptr(y), _ptr(stats),
              _ptr(sc), N, H, W, C_, C.c_float(1e-3), act, _st())
    assert _relerr(y, R.nchw_to_nhwc(yo)) < 5e-6


@pytest.mark.parametrize("shape", [(2, 32, 32, 64), (1, 8, 8, 512), (2, 15, 17, 128)])
def test_gram_and_pool(built_lib, shape):
    from faststyle_b200 import _lib
    N, H, W, C_ = shape
    rng = np.random.RandomState(9)
    f = np.abs(rng.standard_normal((N, H, W, C_))).astype(np.float32)
    fo = R.nhwc_to_nchw(torch.from_numpy(f).double())
    g = torch.empty((N, C_, C_), device="cuda")
    nsc = _lib.load().fs_gram_scratch_floats(N, C_)
    sc = torch.empty(nsc, device="cuda")
    fd = _dev(f)
    _lib.call("fs_gram_forward", _ptr(fd), _ptr(g), _ptr(sc), C.c_longlong(nsc), N, H, W, C_, _st())
    assert _relerr(g, R.gram(fo)) < 2e-6
    p = torch.empty((N, (H + 1) // 2, (W + 1) // 2, C_), device="cuda")
    _lib.call("fs_maxpool2x2", _ptr(fd), _ptr(p), N, H, W, C_, _st())
    assert torch.equal(p.cpu().double(), R.nchw_to_nhwc(R.max_pool_same(fo)))


def test_adam_matches_tf_rule(built_lib):
    from faststyle_b200.engine import TFAdam
    rng = np.random.RandomState(11)
    p0 = rng.standard_normal(10007).astype(np.float32)
    p = torch.from_numpy(p0.copy()).cuda()
    opt = TFAdam(p, lr=1e-3)
    po = {"p": torch.from_numpy(p0.copy()).double()}
    oo = R.TFAdam(po, 1e-3)
    for _ in range(5):
        g = rng.standard_normal(10007).astype(np.float32)
        opt.step(torch.from_numpy(g).cuda())
        oo.step(po, {"p": torch.from_numpy(g).double()})
    assert _relerr(p, po["p"]) < 1e-6
    assert int(opt.step_counter.item()) == 5
