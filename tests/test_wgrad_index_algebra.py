"""CPU: the index algebra of the tap-packed weight-gradient kernel (csrc/conv_tcgen05.cu, conv_wgrad_packed_kernel +
wgrad_packed_scatter_kernel), restated in numpy and checked against torch's conv2d weight gradient.

The kernel's MMAs are emulated as plain matrix products over exactly the operands its descriptors address:
  * M operand: 4 boxes of 32 channels x (8 px x TR rows), no halo (optionally "folded": boxes 2,3 = boxes 0,1 read RG rows up)
  * N operand: one 32-channel box with a (kw-1) x (RG-1) halo at origin (ox, oy + u0); N block j = the box shifted by j pixels,
    row shift u = the box shifted by u rows
  * accumulator column (u_local * kw + j) * 32 + c, row = M channel; scatter: tap (r0 + dr*u, s0 + ds*j)
This pins the geometry (origins, tap maps, stride-2 parity classes, the 64-channel row folding) independently of the GPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def s2_axis(ks):
    pad = ks // 2
    k, rmin, dmin = [0, 0], [0, 0], [0, 0]
    for q in range(2):
        rmin[q] = (pad + q) & 1
        k[q] = (ks - 1 - rmin[q]) // 2 + 1
        e = rmin[q] - pad - q
        dmin[q] = e // 2 if e >= 0 else -((-e) // 2)
    return k, rmin, dmin


def problems(ks, stride, Ct, Cout, fold):
    """Restates plan_wgrad_packed's geometry: one dict per problem (1 for stride 1, 4 parity classes for stride 2)."""
    pad = ks // 2
    m_from_x = Ct > Cout
    out = []
    if stride == 1:
        g = dict(kh=ks, kw=ks, py=0, px=0)
        if m_from_x:
            g.update(oy=pad - (ks - 1), ox=pad - (ks - 1), r0=ks - 1, s0=ks - 1, dr=-1, ds=-1)
        else:
            g.update(oy=-pad, ox=-pad, r0=0, s0=0, dr=1, ds=1)
        out.append(g)
    else:
        k, rmin, dmin = s2_axis(ks)
        for cls in range(4):
            py, px = cls >> 1, cls & 1
            g = dict(kh=k[py], kw=k[px], py=py, px=px)
            if m_from_x:
                g.update(oy=-(dmin[py] + k[py] - 1), ox=-(dmin[px] + k[px] - 1), r0=rmin[py] + 2 * (k[py] - 1), dr=-2,
                         s0=rmin[px] + 2 * (k[px] - 1), ds=-2)
            else:
                g.update(oy=dmin[py], ox=dmin[px], r0=rmin[py], dr=2, s0=rmin[px], ds=2)
            out.append(g)
    for g in out:
        g['m_from_x'] = m_from_x
        g['RG'] = min(3, g['kh'])
        Mch = Ct if m_from_x else Cout
        g['mfold'] = bool(fold and stride == 1 and Mch == 64 and g['kh'] > g['RG'] and g['kh'] <= 2 * g['RG'])
    return out


def emulate(x, dz, ks, stride, fold=False, TR=4):
    """dW [Cout, Ct, ks, ks] from x [N, Ct, H, W], dz [N, Cout, Ho, Wo] through the kernel's tiling."""
    N, Ct, H, W = x.shape
    Cout = dz.shape[1]
    dw = np.zeros((Cout, Ct, ks, ks), dtype=np.float64)
    for g in problems(ks, stride, Ct, Cout, fold):
        plane = x[:, :, g['py']::2, g['px']::2] if stride == 2 else x       # the parity plane the 5-D TMA view addresses
        Hp, Wp = dz.shape[2], dz.shape[3]
        m_op, n_op = (plane, dz) if g['m_from_x'] else (dz, plane)
        Mch, Nch = m_op.shape[1], n_op.shape[1]

        def fetch(t, n, c0, y, xx, h, w):      # TMA box of 32 channels, zero-filled out of bounds (channels too)
            box = np.zeros((32, h, w))
            for ci in range(32):
                if c0 + ci >= t.shape[1]:
                    continue
                for yy in range(h):
                    for xc in range(w):
                        Y, X = y + yy, xx + xc
                        if 0 <= Y < t.shape[2] and 0 <= X < t.shape[3]:
                            box[ci, yy, xc] = t[n, c0 + ci, Y, X]
            return box
        RG, kh, kw = g['RG'], g['kh'], g['kw']
        row_groups = 1 if g['mfold'] else -(-kh // RG)
        tiles_y = -(-(Hp + (RG if g['mfold'] else 0)) // TR)
        for rg in range(row_groups):
            u0 = rg * RG
            rows = min(RG, kh - u0)
            for mb in range(-(-Mch // 128)):
                for nb in range(Nch // 32):
                    acc = np.zeros((128, RG * kw * 32))
                    for n in range(N):
                        for ty in range(tiles_y):
                            for tx in range(-(-Wp // 8)):
                                x0, y0 = tx * 8, ty * TR
                                a = np.zeros((128, TR, 8))
                                for q in range(4):
                                    ch, ym = (mb * 4 + q) * 32, y0
                                    if g['mfold'] and q >= 2:
                                        ch, ym = ch - 64, ym - RG
                                    a[q * 32:(q + 1) * 32] = fetch(m_op, n, ch, ym, x0, TR, 8)
                                b = fetch(n_op, n, nb * 32, y0 + g['oy'] + u0, x0 + g['ox'], TR + RG - 1, 8 + kw - 1)
                                for u in range(rows):
                                    for kk in range(TR):                      # one MMA: K = the 8 pixels of image row kk
                                        for j in range(kw):                   # N block j = box shifted by j pixels
                                            acc[:, (u * kw + j) * 32:(u * kw + j + 1) * 32] += a[:, kk, :] @ b[:, kk + u, j:j + 8].T
                    for row in range(128):                                    # wgrad_packed_scatter_kernel
                        mch, ufold = mb * 128 + row, 0
                        if g['mfold'] and mch >= 64:
                            mch, ufold = mch - 64, RG
                        if mch >= Mch:
                            continue
                        for t in range(rows * kw):
                            u, j = u0 + t // kw + ufold, t % kw
                            if u >= kh:
                                continue
                            r, s = g['r0'] + g['dr'] * u, g['s0'] + g['ds'] * j
                            for c in range(32):
                                nch = nb * 32 + c
                                co, ci = (nch, mch) if g['m_from_x'] else (mch, nch)
                                dw[co, ci, r, s] += acc[row, t * 32 + c]
    return dw


@pytest.mark.parametrize('Ct,Cout,ks,stride,H,W,fold', [
    (32, 64, 3, 1, 6, 10, False),      # M = dZ
    (64, 32, 3, 1, 5, 9, False),       # M = X
    (64, 32, 5, 1, 7, 9, False),       # two row groups (3 + 2 filter rows)
    (64, 32, 5, 1, 7, 9, True),        # 64-channel M operand folded: rows 64..127 = the same channels 3 rows up (RAMNET_WGRAD_FOLD)
    (32, 64, 5, 1, 6, 8, True),        # folded, M = dZ
    (32, 64, 5, 2, 8, 12, False),      # stride 2: parity classes 3x3 / 3x2 / 2x3 / 2x2, M = dZ
    (64, 32, 5, 2, 8, 8, False),       # stride 2, M = parity plane of X
    (32, 32, 3, 2, 6, 8, False),       # stride 2, 3x3: classes 1x1 / 1x2 / 2x1 / 2x2
])
def test_tap_packed_geometry_reproduces_conv_weight_gradient(Ct, Cout, ks, stride, H, W, fold):
    g = torch.Generator().manual_seed(Ct * 7 + Cout + ks + stride)
    x = torch.randn(1, Ct, H, W, generator=g, dtype=torch.float64)
    w = torch.zeros(Cout, Ct, ks, ks, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, None, stride=stride, padding=ks // 2)
    dz = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(dz)
    ours = emulate(x.numpy(), dz.numpy(), ks, stride, fold=fold)
    np.testing.assert_allclose(ours, w.grad.numpy(), rtol=1e-9, atol=1e-9)
