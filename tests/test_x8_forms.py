"""CPU check of the algebra behind the x8 forms of the 9x9 layers (reference im_transf_net.py:38,70: 9x9 stride-1 SAME
convolutions 3 -> 16 and 16 -> 3).  The CUDA side (Engine::tc9, prep.cu PJ_X16 with the x8 bit, wgrad9_tc.cu) runs

  forward / data gradient:  out[y, X, (dxo, co)] = sum_kh  Xw[y + kh - 4, X, :] @ Bt[kh]          (one horizontal tap)
  weight gradient:          Wt[kh][k][n] = sum_{y, X} Xw[y + kh - 4, X, k] * Dg[y, X, n],  folded back over dxo

with Xw = windowed planes [H, W/8, 16 px x 4 ch] (group X = pixels 8X-4 .. 8X+11) and Dg / out = grouped planes
[H, W/8, 8 px x 16 ch].  These tests restate both in numpy and compare with a direct SAME convolution / its weight
gradient, so the index conventions of the kernels are pinned independently of any GPU."""
import numpy as np

H, W = 7, 24
G = W // 8


def windowed(t):
    """[H, W, 4] -> [H, G, 64]: group X holds pixels 8X-4 .. 8X+11 (zero outside the image)."""
    out = np.zeros((H, G, 16, 4))
    for X in range(G):
        for dxi in range(16):
            px = 8 * X - 4 + dxi
            if 0 <= px < W:
                out[:, X, dxi, :] = t[:, px, :]
    return out.reshape(H, G, 64)


def conv_same(x, w):
    ci, co = w.shape[2], w.shape[3]
    xp = np.zeros((H + 8, W + 8, ci))
    xp[4:4 + H, 4:4 + W] = x
    y = np.zeros((H, W, co))
    for kh in range(9):
        for kw in range(9):
            y += xp[kh:kh + H, kw:kw + W] @ w[kh, kw]
    return y


def conv_wgrad(x, dy):
    ci, co = x.shape[2], dy.shape[2]
    xp = np.zeros((H + 8, W + 8, ci))
    xp[4:4 + H, 4:4 + W] = x
    dw = np.zeros((9, 9, ci, co))
    for kh in range(9):
        for kw in range(9):
            dw[kh, kw] = np.einsum("yxc,yxd->cd", xp[kh:kh + H, kw:kw + W], dy)
    return dw


def toeplitz_x8(w, mode):
    """prep.cu PJ_X16 with the x8 bit: Bt[kh][k = dxi*4 + c][n = dxo*16 + m], kw = dxi - dxo.
    mode 0: forward 4(3) -> 16; mode 1: data gradient of a 16 -> 4(3) conv (flipped taps, transposed channels)."""
    A, B = w.shape[2], w.shape[3]
    bt = np.zeros((9, 64, 128))
    for kh in range(9):
        for dxi in range(16):
            for dxo in range(8):
                kw = dxi - dxo
                if not 0 <= kw <= 8:
                    continue
                for c in range(4):
                    for m in range(16):
                        if mode == 0 and c < A and m < B:
                            bt[kh, dxi * 4 + c, dxo * 16 + m] = w[kh, kw, c, m]
                        if mode == 1 and c < B and m < A:
                            bt[kh, dxi * 4 + c, dxo * 16 + m] = w[8 - kh, 8 - kw, m, c]
    return bt


def conv_x8(xw, bt):
    out = np.zeros((H, G, 128))
    for kh in range(9):
        for y in range(H):
            yy = y + kh - 4
            if 0 <= yy < H:
                out[y] += xw[yy] @ bt[kh]
    return out.reshape(H, W, 16)


def test_x8_forward_equals_same_conv():
    rng = np.random.RandomState(0)
    x = rng.randn(H, W, 4); x[..., 3] = 0
    w = rng.randn(9, 9, 3, 16)
    got = conv_x8(windowed(x), toeplitz_x8(w, 0))
    assert np.abs(got - conv_same(x[..., :3], w)).max() < 1e-12


def test_x8_data_gradient_of_16_to_3_conv():
    rng = np.random.RandomState(1)
    w = rng.randn(9, 9, 16, 3)
    dy = rng.randn(H, W, 4); dy[..., 3] = 0
    got = conv_x8(windowed(dy), toeplitz_x8(w, 1))
    # data gradient of y = conv_same(x, w): dx = conv_same(dy, flip(w) with channels transposed)
    wf = np.transpose(w[::-1, ::-1], (0, 1, 3, 2))
    assert np.abs(got - conv_same(dy[..., :3], wf)).max() < 1e-12


def _wt(xw, dg):
    wt = np.zeros((9, 64, 128))
    for a in range(9):
        for y in range(H):
            yy = y + a - 4
            if 0 <= yy < H:
                wt[a] += np.einsum("gk,gn->kn", xw[yy], dg[y])
    return wt


def test_x8_weight_gradient_fold_mode0_and_mode1():
    rng = np.random.RandomState(2)
    # mode 0 (initconv_0): the 4-channel side is the conv input
    x = rng.randn(H, W, 4); x[..., 3] = 0
    dy = rng.randn(H, W, 16)
    wt = _wt(windowed(x), dy.reshape(H, G, 128))
    got = np.zeros((9, 9, 3, 16))
    for kh in range(9):
        for kw in range(9):
            for ci in range(3):
                for co in range(16):
                    got[kh, kw, ci, co] = sum(wt[kh, (d + kw) * 4 + ci, d * 16 + co] for d in range(8))
    assert np.abs(got - conv_wgrad(x[..., :3], dy)).max() < 1e-11
    # mode 1 (upsample_2): the 4-channel side is the output gradient
    x = rng.randn(H, W, 16)
    dy = rng.randn(H, W, 4); dy[..., 3] = 0
    wt = _wt(windowed(dy), x.reshape(H, G, 128))
    got = np.zeros((9, 9, 16, 3))
    for kh in range(9):
        for kw in range(9):
            for ci in range(16):
                for co in range(3):
                    got[kh, kw, ci, co] = sum(wt[8 - kh, (d + 8 - kw) * 4 + co, d * 16 + ci] for d in range(8))
    assert np.abs(got - conv_wgrad(x, dy[..., :3])).max() < 1e-11
