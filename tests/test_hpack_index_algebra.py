"""CPU: index algebra of the horizontal-tap-packed forward convolution (conv_tcgen05_halo_kernel in hpack mode,
RAMNET_HPACK=1, csrc/conv_tcgen05.cu) restated in numpy against torch's conv2d.

For layers with few output channels (dec2: 64 -> 32, 5x5) a 128 x N x 8 MMA with N = 32 is bound by the A-operand
shared-memory reads.  hpack makes the kw horizontal taps extra GEMM COLUMNS instead of extra MMAs:
  E[p, (s, co)] = sum_{r, c} X[p_y + r - pad, p_x, c] * W[r, s, co, c]        (kh row-shifted MMAs of N = kw * cs columns)
  D[p, co]      = sum_s E[(p_y, p_x + s - pad), (s, co)]                      (a shifted sum across the 32 lanes of a warp)
A tile is 32 pixels wide x 4 rows (TMEM lane = 32 * row + x, so one warp quarter = one image row and the shifted sum is
warp shuffles); tiles overlap by kw - 1 columns (valid outputs: x in [pad, 32 - pad)); weights are packed
[r][slice * kw * cs + s * cs + co_l][c]."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def pack_hpack(w, cs):
    """w [Cout, Cin, kh, kw] -> [kh][(Cout / cs) * kw * cs][Cin]   (ramnet_pack_weights_hpack)."""
    Cout, Cin, kh, kw = w.shape
    out = np.zeros((kh, (Cout // cs) * kw * cs, Cin))
    for r in range(kh):
        for co in range(Cout):
            sl, col = co // cs, co % cs
            for s in range(kw):
                out[r, sl * kw * cs + s * cs + col] = w[co, :, r, s]
    return out


def emulate(x, w, bias, cs=32, pty=1):
    N, Cin, H, W = x.shape
    Cout, _, kh, kw = w.shape
    pad = kh // 2
    wp = pack_hpack(w, cs)
    BN = kw * cs
    step_x = 32 - (kw - 1)
    y = np.zeros((N, Cout, H, W))
    written = np.zeros((N, Cout, H, W), dtype=int)
    for n in range(N):
        for pyi in range(-(-H // (4 * pty))):
            for pxi in range(-(-W // step_x)):
                x0, y0 = pxi * step_x - (kw // 2), pyi * 4 * pty          # tile origin (input / E columns x0 .. x0+31)
                # halo box: 32 px x (4*pty + kh - 1) rows x all channels, zero-filled out of bounds (TMA)
                HY = 4 * pty + kh - 1
                box = np.zeros((Cin, HY, 32))
                for yy in range(HY):
                    for xx in range(32):
                        Y, X = y0 - pad + yy, x0 + xx
                        if 0 <= Y < H and 0 <= X < W:
                            box[:, yy, xx] = x[n, :, Y, X]
                for sl in range(Cout // cs):
                    for tl in range(pty):
                        # accumulator of one tile: rows m = 32 * row + x, columns (s, co_l)
                        E = np.zeros((128, BN))
                        for r in range(kh):                                # one group of MMAs per filter row
                            a = box[:, tl * 4 + r: tl * 4 + r + 4, :]      # descriptor start shifted by r halo rows
                            a = a.reshape(Cin, 128).T                      # [m = 32*row + x][c]
                            E += a @ wp[r, sl * BN:(sl + 1) * BN].T
                        for q in range(4):                                 # warp quarter = image row of the tile
                            for lane in range(32):
                                ox, oy = x0 + lane, y0 + tl * 4 + q
                                if not (kw // 2 <= lane < 32 - kw // 2 and 0 <= ox < W and oy < H):
                                    continue
                                for col in range(cs):
                                    acc = 0.0
                                    for s in range(kw):                    # shuffle from lane + s - pad of the same quarter
                                        acc += E[32 * q + lane + s - kw // 2, s * cs + col]
                                    co = sl * cs + col
                                    y[n, co, oy, ox] = acc + bias[co]
                                    written[n, co, oy, ox] += 1
    assert (written == 1).all()                                            # every output pixel exactly once
    return y


@pytest.mark.parametrize('Cin,Cout,k,H,W,pty', [(32, 32, 5, 6, 40, 1), (64, 32, 3, 9, 33, 2), (32, 64, 5, 4, 57, 1)])
def test_hpack_geometry_reproduces_conv(Cin, Cout, k, H, W, pty):
    g = torch.Generator().manual_seed(Cin + Cout + k)
    x = torch.randn(1, Cin, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, k, k, generator=g, dtype=torch.float64) * 0.1
    b = torch.randn(Cout, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, b, padding=k // 2).numpy()
    np.testing.assert_allclose(emulate(x.numpy(), w.numpy(), b.numpy(), pty=pty), ref, rtol=1e-9, atol=1e-9)
