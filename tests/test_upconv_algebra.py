"""CPU: the index algebra of the up-conv mode (RAMNET_FLAG_UPCONV, csrc/conv_tcgen05.cu conv_up_fwd_tf32): bilinear x2
(align_corners=False) + zero-padded 5x5 convolution == one 5x5 convolution of the LOW-resolution tensor with
4 * Cout phase columns and collapsed taps, PLUS eight border segments (first / last row, first / last column, four
corners) that read edge-only views of the same tensor.  The coefficient functions below restate up_cu / up_dT / up_dB and
the segment table of the CUDA source; the sum of all segments must equal F.conv2d(F.interpolate(x)) (the reference's
UpsampleConvLayer.forward, submodules.py:87-97) exactly in fp64, down to 2x2 inputs where every segment overlaps."""
import itertools

import pytest
import torch
import torch.nn.functional as F


def cu(p, t, r):            # coefficient of input row m + r in upsampled row 2m + p + t (uniform formula)
    s = p + t
    if s % 2 == 0:
        e = s // 2
        return {e - 1: .25, e: .75}.get(r, 0.)
    e = (s - 1) // 2
    return {e: .75, e + 1: .25}.get(r, 0.)


def dT(p, t, r):            # first-row correction
    return {(0, 0, 0): .25, (0, -1, 0): -.25, (1, -1, 0): .25, (1, -2, 0): -.25, (0, -2, -1): .25}.get((p, t, r), 0.)


def dB(p, t, r):            # last-row correction = mirror image
    return dT(1 - p, -t, -r)


# segment: (operator along y, operator along x, tap rows, tap columns, view rows, view columns)
SEGS = [(cu, cu, range(-2, 3), range(-2, 3), 'all', 'all'),
        (dT, cu, range(-1, 1), range(-2, 3), 'first', 'all'), (dB, cu, range(0, 2), range(-2, 3), 'last', 'all'),
        (cu, dT, range(-2, 3), range(-1, 1), 'all', 'first'), (cu, dB, range(-2, 3), range(0, 2), 'all', 'last'),
        (dT, dT, range(-1, 1), range(-1, 1), 'first', 'first'), (dT, dB, range(-1, 1), range(0, 2), 'first', 'last'),
        (dB, dT, range(0, 2), range(-1, 1), 'last', 'first'), (dB, dB, range(0, 2), range(0, 2), 'last', 'last')]


def _view(n, which):
    return slice(0, n) if which == 'all' else (slice(0, 1) if which == 'first' else slice(n - 1, n))


def upconv_by_segments(x, w):
    N, C, H, W = x.shape
    Co = w.shape[0]
    out = torch.zeros(N, Co, 2 * H, 2 * W, dtype=x.dtype)
    for fy, fx, rr, ss, vy, vx in SEGS:
        xv = torch.zeros_like(x)
        xv[:, :, _view(H, vy), _view(W, vx)] = x[:, :, _view(H, vy), _view(W, vx)]     # the edge-only view, zero elsewhere
        xp = F.pad(xv, (2, 2, 2, 2))                                                # TMA zero fill
        for py, px in itertools.product((0, 1), (0, 1)):
            for r in rr:
                for s in ss:
                    wc = sum(fy(py, t, r) * fx(px, u, s) * w[:, :, t + 2, u + 2]
                             for t in range(-2, 3) for u in range(-2, 3) if fy(py, t, r) and fx(px, u, s))
                    if isinstance(wc, int):
                        continue
                    out[:, :, py::2, px::2] += torch.einsum('oc,nchw->nohw', wc, xp[:, :, 2 + r:2 + r + H, 2 + s:2 + s + W])
    return out


@pytest.mark.parametrize('shape', [(1, 3, 2, 8, 8), (2, 2, 3, 4, 6), (1, 2, 2, 2, 2), (1, 2, 2, 3, 5), (1, 1, 1, 16, 24),
                                   (1, 1, 2, 2, 9)])
def test_upconv_segments_equal_interpolate_then_conv(shape):
    N, C, Co, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, C, H, W, dtype=torch.float64, generator=g)
    w = torch.randn(Co, C, 5, 5, dtype=torch.float64, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False), w, padding=2)
    assert (upconv_by_segments(x, w) - ref).abs().max().item() <= 1e-12


def test_main_segment_alone_is_exact_in_the_interior_only():
    """Why the border segments exist: the uniform collapse differs from the reference in a 3-pixel ring."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 2, 8, 8, dtype=torch.float64, generator=g)
    w = torch.randn(2, 2, 5, 5, dtype=torch.float64, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False), w, padding=2)
    global SEGS
    keep, SEGS = SEGS, SEGS[:1]
    try:
        main = upconv_by_segments(x, w)
    finally:
        SEGS = keep
    err = (main - ref).abs()
    assert err[:, :, 3:-3, 3:-3].max().item() <= 1e-12 and err.max().item() > 1e-3
