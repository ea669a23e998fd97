"""GPU tests of the hpack forward (RAMNET_FLAG_HPACK), the 64-channel row folding and the fused split sum of the
tap-packed weight gradient.  Written at the end of round 1, validated on hardware as the first call of round 2
(profiles/r02_first_call_experimental_paths.txt) and on by default since (RAMNET_HPACK=0 / RAMNET_WGRAD_FOLD=0 /
RAMNET_WGRAD_FUSED_SUM=0 switch them off)."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu]


def dev():
    return torch.device('cuda', 0)


def nhwc(t):
    return t.to(dev()).contiguous(memory_format=torch.channels_last)


def rna(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)


@pytest.mark.parametrize('N,Cin,H,W,k,epi', [(1, 32, 8, 40, 5, 'relu'), (2, 64, 11, 57, 5, 'relu'), (1, 64, 9, 33, 3, 'res'),
                                            (1, 64, 16, 96, 5, 'pred')])
def test_hpack_forward_vs_torch(N, Cin, H, W, k, epi):
    from rpg_ramnet_b200 import ops
    g = torch.Generator().manual_seed(N + Cin + H + W)
    x = rna(torch.randn(N, Cin, H, W, generator=g))
    w = rna(torch.randn(32, Cin, k, k, generator=g) * (1.0 / (Cin * k * k)) ** 0.5)
    b = torch.randn(32, generator=g) * 0.1
    res = torch.randn(N, 32, H, W, generator=g) if epi == 'res' else None
    y = F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)
    y = torch.relu(y + res.double()) if res is not None else torch.relu(y)
    wp = ops.pack_weights_hpack(w.to(dev()))
    assert getattr(wp, '_ramnet_hpack', False)
    if epi == 'pred':
        pw, pb = torch.randn(32, generator=g) * 0.3, torch.randn(1, generator=g)
        ref = torch.sigmoid((y * pw.double().view(1, -1, 1, 1)).sum(1, keepdim=True) + pb.double())
        out = ops.conv_fwd(nhwc(x), None, wp, b.to(dev()), 32, k, 1, ops.EPI_BIAS_RELU_PRED, ops.MMA_TF32,
                           aux0=pw.to(dev()), aux1=pb.to(dev()))
        out = out[0] if isinstance(out, tuple) else out
        assert (out.cpu().double() - ref).abs().max().item() <= 2e-5
        return
    epi_id = ops.EPI_BIAS_RES_RELU if res is not None else ops.EPI_BIAS_RELU
    out = ops.conv_fwd(nhwc(x), None, wp, b.to(dev()), 32, k, 1, epi_id, ops.MMA_TF32,
                       aux0=None if res is None else nhwc(res))
    assert (out.cpu().double() - y).abs().max().item() <= 2e-5 * max(1.0, y.abs().max().item())


@pytest.mark.parametrize('shape', [(1, 16, 24, 64, 32, 5), (1, 10, 20, 32, 64, 5), (2, 8, 16, 64, 64, 3)])
def test_wgrad_fold_and_fused_sum_vs_torch(shape):
    """Folding and the fused sum are the defaults (the library reads RAMNET_WGRAD_FOLD / _FUSED_SUM once)."""
    from rpg_ramnet_b200 import ops
    N, H, W, Ct, Cout, k = shape
    g = torch.Generator().manual_seed(sum(shape))
    x, dz = rna(torch.randn(N, Ct, H, W, generator=g)), rna(torch.randn(N, Cout, H, W, generator=g))
    w = torch.zeros(Cout, Ct, k, k, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, padding=k // 2).backward(dz.double())
    dw = torch.zeros(Cout, Ct, k, k, device=dev())
    ops.conv_wgrad(nhwc(dz), nhwc(x), None, Cout, k, 1, dw, None, ops.MMA_TF32)
    err = (dw.cpu().double() - w.grad).norm() / w.grad.norm()
    assert float(err) <= 1e-5


@pytest.mark.parametrize('N,Cin,Cout,H,W,relu', [(1, 32, 64, 32, 64, True), (2, 64, 128, 18, 46, True), (1, 32, 32, 13, 21, False),
                                                 (1, 128, 256, 64, 128, True), (3, 32, 64, 5, 7, True), (4, 32, 64, 256, 512, True)])
def test_s2seg_forward_vs_torch(N, Cin, Cout, H, W, relu):
    """5x5 stride-2 convolution as four parity-plane K segments (RAMNET_FLAG_S2SEG) == F.conv2d(stride=2, padding=2),
    even and odd sizes (the odd planes are one row / column shorter; TMA zero fill is the padding)."""
    from rpg_ramnet_b200 import ops
    g = torch.Generator().manual_seed(N + Cin + Cout + H + W)
    x = rna(torch.randn(N, Cin, H, W, generator=g))
    w = rna(torch.randn(Cout, Cin, 5, 5, generator=g) * (1.0 / (Cin * 25)) ** 0.5)
    b = torch.randn(Cout, generator=g) * 0.1
    y = F.conv2d(x.to(dev()).double(), w.to(dev()).double(), b.to(dev()).double(), stride=2, padding=2)
    y = torch.relu(y) if relu else y
    wp = ops.pack_weights_s2seg(w.to(dev()))
    assert getattr(wp, '_ramnet_s2seg', False)
    out = ops.conv_fwd(nhwc(x), None, wp, b.to(dev()), Cout, 5, 2, ops.EPI_BIAS_RELU if relu else ops.EPI_BIAS, ops.MMA_TF32)
    assert tuple(out.shape) == tuple(y.shape)
    assert (out.double() - y).abs().max().item() <= 2e-5 * max(1.0, y.abs().max().item())
    if H % 2 or W % 2:
        return            # the single-stage stride-2 path it replaces takes even sizes only
    old = ops.conv_fwd(nhwc(x), None, ops.pack_weights(w.to(dev()), ops.MMA_TF32), b.to(dev()), Cout, 5, 2,
                       ops.EPI_BIAS_RELU if relu else ops.EPI_BIAS, ops.MMA_TF32)
    assert (out - old).abs().max().item() <= 2e-5 * max(1.0, y.abs().max().item())
