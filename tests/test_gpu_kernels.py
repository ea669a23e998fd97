"""GPU: every C-ABI kernel against the CPU oracle on seeded inputs (through the ctypes boundary)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ramnet_oracle as O
from helpers import GOLDEN

pytestmark = pytest.mark.gpu

KINDS = ['fp32', 'tf32']
# fp32 FFMA path: summation-order noise only.  tf32: operands carry 10 mantissa bits, fp32 accumulate
# -> error ~ 2^-11 * sqrt(K) relative to the magnitude of the accumulated terms.
TOL = {'fp32': 2e-5, 'tf32': 2e-3}


def dev():
    return torch.device('cuda', 0)


def nhwc(t):
    """CPU [N,C,H,W] -> CUDA NHWC-strided logical [N,C,H,W]."""
    return t.to(dev()).contiguous(memory_format=torch.channels_last)


def kind_id(k):
    from rpg_ramnet_b200 import ops
    return {'fp32': ops.MMA_FP32, 'tf32': ops.MMA_TF32}[k]


# ----------------------------------------------------------------------------- voxel grid
def _votes_check(ev, b, w, h):
    from rpg_ramnet_b200 import ops
    il, vl, ir, vr = ops.voxel_votes(torch.from_numpy(ev).to(dev()), b, w, h)
    oil, ovl, oir, ovr = O.voxel_grid_votes(ev, b, w, h)
    assert np.array_equal(il.cpu().numpy(), oil)            # integer index stream: bit exact
    assert np.array_equal(ir.cpu().numpy(), oir)
    assert np.array_equal(vl.cpu().numpy(), ovl.astype(np.float32))   # float64 math, one fp32 rounding: bit exact
    assert np.array_equal(vr.cpu().numpy(), ovr.astype(np.float32))


def test_voxel_grid_golden_cases():
    """The reference's own outputs (tests/golden/voxel.npz): indices bit-exact, sums to accumulation order."""
    import rpg_ramnet_b200 as R
    g = np.load(os.path.join(GOLDEN, 'voxel.npz'))
    for n in sorted({k.split('/')[0] for k in g.files}):
        b, w, h = (int(v) for v in g[n + '/shape'])
        ev = g[n + '/events']
        keep = ev.copy()
        out = R.events_to_voxel_grid(ev, b, w, h).cpu().numpy()
        assert np.array_equal(ev, keep)
        ref = g[n + '/grid']
        cnt = np.zeros(b * h * w)
        il, _, ir, _ = O.voxel_grid_votes(ev, b, w, h)
        np.add.at(cnt, il[il >= 0], 1)
        np.add.at(cnt, ir[ir >= 0], 1)
        tol = 1.2e-7 * np.maximum(cnt.reshape(b, h, w), 1) * 2       # <= 1 ulp(1.0) per accumulation step
        assert np.all(np.abs(out - ref) <= tol), n
        single = cnt.reshape(b, h, w) <= 1
        assert np.array_equal(out[single], ref[single]), n          # voxels with one vote: bit exact
        _votes_check(ev, b, w, h)


@pytest.mark.parametrize('n,hot', [(0, False), (1, False), (777, False), (100_000, False), (100_000, True),
                                   (1_000_000, False)])
def test_voxel_grid_seeded(n, hot):
    import rpg_ramnet_b200 as R
    W, H, B = 512, 256, 5
    ev = O.synth_events(n, W, H, seed=100 + n % 97, hot=hot) if n else np.zeros((0, 4))
    out = R.events_to_voxel_grid(ev, B, W, H).cpu().numpy()
    ref = O.voxel_grid(ev, B, W, H)
    cnt = np.zeros(B * H * W)
    if n:
        il, _, ir, _ = O.voxel_grid_votes(ev, B, W, H)
        np.add.at(cnt, il[il >= 0], 1)
        np.add.at(cnt, ir[ir >= 0], 1)
        _votes_check(ev, B, W, H)
    tol = 1.2e-7 * np.maximum(cnt.reshape(B, H, W), 1) ** 1.5 * 2
    assert np.all(np.abs(out - ref) <= tol)


@pytest.mark.parametrize('n,hot', [(2_000_000, False), (2_000_000, True)])
def test_voxel_grid_packed_accumulator_vs_oracle(n, hot):
    """>= 1.5e6 events take the packed path of ramnet_voxel_grid_ex (pixel-major accumulator, both votes of an event in
    one 128-bit vector reduction, then a transpose into [bins, H, W]): same indices, sums to accumulation order."""
    from rpg_ramnet_b200 import ops
    W, H, B = 512, 256, 5
    ev = O.synth_events(n, W, H, seed=11, hot=hot)
    evd = torch.from_numpy(ev).to(dev())
    out, stats = ops.voxel_grid_ex(evd, B, W, H, want_stats=True)
    out = out.cpu().numpy()
    ref = O.voxel_grid(ev, B, W, H)
    cnt = np.zeros(B * H * W)
    il, _, ir, _ = O.voxel_grid_votes(ev, B, W, H)
    np.add.at(cnt, il[il >= 0], 1)
    np.add.at(cnt, ir[ir >= 0], 1)
    cnt = cnt.reshape(B, H, W)
    tol = 1.2e-7 * np.maximum(cnt, 1) ** 1.5 * 2
    assert np.all(np.abs(out - ref) <= tol)
    assert np.array_equal(out[cnt == 0], np.zeros_like(out[cnt == 0]))
    # the direct kernel (small-n path) on the same events agrees to accumulation order as well
    direct = ops.voxel_grid(evd, B, W, H).cpu().numpy()
    assert np.all(np.abs(out - direct) <= 2 * tol)
    # fused statistics of the non-zero voxels
    nz = out[out != 0].astype(np.float64)
    st = stats.cpu().numpy()
    assert st[2] == nz.size
    np.testing.assert_allclose(st[:2], [nz.sum(), (nz ** 2).sum()], rtol=1e-9)


@pytest.mark.parametrize('n', [50_000, 2_000_000])
def test_voxel_grid_scatter_then_normalise_in_one_call(n):
    """events -> grid -> loaders' normalisation (event_dataset.py:144-151) in one C call, both paths,
    against the numpy oracle of the two reference steps."""
    import rpg_ramnet_b200 as R
    from oracle import dataio_oracle as D
    W, H, B = 512, 256, 5
    ev = O.synth_events(n, W, H, seed=13)
    out = R.events_to_voxel_grid(ev, B, W, H, normalize=True).cpu().numpy()
    ref = D.normalize_voxel_grid(O.voxel_grid(ev, B, W, H))
    assert np.array_equal(out == 0, ref == 0)
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-5)
    nz = out[out != 0].astype(np.float64)
    assert abs(nz.mean()) <= 1e-4 and abs(nz.std() - 1.0) <= 1e-4


def test_voxel_grid_full_size_properties():
    """BASELINE config 5 at 10M events: size-independent properties instead of a CPU replay."""
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200 import ops
    W, H, B, n = 512, 256, 5, 10_000_000
    ev = O.synth_events(n, W, H, seed=5)
    evd = torch.from_numpy(ev).to(dev())
    g1 = ops.voxel_grid(evd, B, W, H)
    # (1) mass conservation: every event with ti+1 < B deposits exactly p; sum(grid) == sum(p) up to fp32 adds
    pol = np.where(ev[:, 3] == 0, -1.0, 1.0)
    ts = (B - 1) * (ev[:, 0] - ev[0, 0]) / (ev[-1, 0] - ev[0, 0])
    expect = float(np.sum(pol * np.where(ts.astype(np.int64) + 1 < B, 1.0, 1.0 - (ts - ts.astype(np.int64)))))
    assert abs(float(g1.double().sum()) - expect) <= 1e-6 * n
    # (2) linearity: voxel(A ++ B) == voxel(A) + voxel(B) when both halves are normalised with the same t0/dT:
    #     flipping every polarity negates the grid exactly (fp32 adds are sign-symmetric)
    ev2 = ev.copy()
    ev2[:, 3] = 1.0 - ev2[:, 3]
    g2 = ops.voxel_grid(torch.from_numpy(ev2).to(dev()), B, W, H)
    assert float((g1 + g2).abs().max()) <= 2e-4
    # (3) permutation of events between the fixed first/last rows changes nothing but summation order
    perm = np.concatenate([[0], 1 + np.random.default_rng(0).permutation(n - 2), [n - 1]])
    g3 = ops.voxel_grid(torch.from_numpy(np.ascontiguousarray(ev[perm])).to(dev()), B, W, H)
    assert float((g1 - g3).abs().max()) <= 2e-4
    # (4) polarity-blind count: |p|=1 so sum |votes| per event is 1 -> grid of all-positive events sums to n'
    ev4 = ev.copy()
    ev4[:, 3] = 1.0
    g4 = ops.voxel_grid(torch.from_numpy(ev4).to(dev()), B, W, H)
    assert float(g4.min()) >= 0.0
    oob = torch.zeros(1, dtype=torch.int32, device=dev())
    ops.voxel_grid(evd, B, W, H, oob_count=oob)
    assert int(oob.item()) == 0


def test_voxel_grid_out_of_bounds_counted_not_written():
    from rpg_ramnet_b200 import ops
    ev = O.synth_events(1000, 32, 24, seed=3)
    ev[10, 1] = 32
    ev[20, 2] = -1
    ev[30, 2] = 24
    oob = torch.zeros(1, dtype=torch.int32, device=dev())
    g = ops.voxel_grid(torch.from_numpy(ev).to(dev()), 5, 32, 24, oob_count=oob)
    assert int(oob.item()) == 3
    keep = np.ones(1000, bool)
    keep[[10, 20, 30]] = False
    il, vl, ir, vr = O.voxel_grid_votes(ev, 5, 32, 24)
    ref = np.zeros(5 * 24 * 32)
    for idx, val in ((il, vl), (ir, vr)):
        ok = keep & (idx >= 0)
        np.add.at(ref, idx[ok], val[ok])
    assert np.allclose(g.cpu().numpy().ravel(), ref, atol=1e-5)


def test_voxel_grid_argument_errors():
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200 import ops
    ev = torch.zeros((4, 4), dtype=torch.float64, device=dev())
    with pytest.raises(R.RamnetError):
        ops.voxel_grid(ev, 0, 8, 8)
    with pytest.raises(R.RamnetError):
        ops.voxel_grid(ev.float(), 5, 8, 8)
    with pytest.raises(AssertionError):
        R.events_to_voxel_grid(np.zeros((3, 3)), 5, 8, 8)


# ----------------------------------------------------------------------------- convolutions
def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def _conv_case(kind, N, H, W, C0, C1, Cout, k, stride, seed):
    from rpg_ramnet_b200 import ops
    x0 = _rand((N, C0, H, W), seed)
    x1 = _rand((N, C1, H, W), seed + 1) if C1 else None
    w = _rand((Cout, C0 + C1, k, k), seed + 2, (1.0 / ((C0 + C1) * k * k)) ** 0.5)
    b = _rand((Cout,), seed + 3, 0.1)
    xin = x0 if x1 is None else torch.cat([x0, x1], 1)
    ref = F.conv2d(xin.double(), w.double(), b.double(), stride=stride, padding=k // 2).float()
    wp = ops.pack_weights(w.to(dev()), kind_id(kind))
    return x0, x1, w, b, ref, wp


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('shape', [
    # N, H, W, C0, C1, Cout, k, stride
    (1, 16, 16, 32, 0, 64, 5, 2),     # encoder
    (2, 12, 20, 64, 0, 128, 5, 2),    # ragged
    (1, 8, 8, 128, 0, 256, 5, 2),
    (2, 16, 24, 64, 0, 64, 3, 1),
    (1, 32, 43, 256, 0, 256, 3, 1),   # resblock width at MVSEC level-2 (W=43)
    (1, 16, 16, 256, 0, 128, 5, 1),   # decoders
    (1, 32, 32, 128, 0, 64, 5, 1),
    (1, 64, 64, 64, 0, 32, 5, 1),
    (3, 9, 7, 32, 32, 32, 3, 1),      # two sources
    (1, 6, 10, 64, 0, 32, 1, 1),      # 1x1
])
def test_conv_bias_relu(kind, shape):
    from rpg_ramnet_b200 import ops
    N, H, W, C0, C1, Cout, k, stride = shape
    x0, x1, w, b, ref, wp = _conv_case(kind, N, H, W, C0, C1, Cout, k, stride, seed=sum(shape))
    y = ops.conv_fwd(nhwc(x0), None if x1 is None else nhwc(x1), wp, b.to(dev()), Cout, k, stride,
                     ops.EPI_BIAS_RELU, kind_id(kind))
    assert y.shape == ref.shape
    err = (y.cpu() - torch.relu(ref)).abs().max().item()
    assert err <= TOL[kind] * max(1.0, ref.abs().max().item()), err
    y2 = ops.conv_fwd(nhwc(x0), None if x1 is None else nhwc(x1), wp, b.to(dev()), Cout, k, stride, ops.EPI_BIAS,
                      kind_id(kind))
    err = (y2.cpu() - ref).abs().max().item()
    assert err <= TOL[kind] * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize('kind', KINDS)
def test_conv_residual_epilogue(kind):
    from rpg_ramnet_b200 import ops
    N, H, W, C = 2, 8, 12, 64
    x0, _, w, b, ref, wp = _conv_case(kind, N, H, W, C, 0, C, 3, 1, seed=11)
    res = _rand((N, C, H, W), 12)
    y = ops.conv_fwd(nhwc(x0), None, wp, b.to(dev()), C, 3, 1, ops.EPI_BIAS_RES_RELU, kind_id(kind), aux0=nhwc(res))
    assert (y.cpu() - torch.relu(ref + res)).abs().max().item() <= TOL[kind] * 4


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('C,H,W,N', [(64, 16, 16, 1), (128, 8, 12, 2), (256, 4, 6, 1), (32, 10, 6, 2)])
def test_convgru_fused(kind, C, H, W, N):
    """Two fused convolutions == ConvGRU.forward (submodules.py:436-454)."""
    from rpg_ramnet_b200 import engine as E, ops
    from rpg_ramnet_b200.model.submodules import ConvGRU
    torch.manual_seed(C + H)
    gru = ConvGRU(C, C, 3)
    with torch.no_grad():
        for g_ in (gru.reset_gate, gru.update_gate, gru.out_gate):
            g_.bias.normal_(0, 0.2)
    sd = {'g.' + k: v.detach() for k, v in gru.state_dict().items()}
    x, h = _rand((N, C, H, W), 1), _rand((N, C, H, W), 2, 0.7)
    ref = O.conv_gru(sd, 'g', x, h)
    gru.to(dev())
    ru, out = E.pack_gru(E.WeightCache(), 'g', gru, kind_id(kind))
    y = E.run_gru(nhwc(x), nhwc(h), ru, out, kind_id(kind))
    assert (y.cpu() - ref).abs().max().item() <= TOL[kind] * 3
    ref0 = O.conv_gru(sd, 'g', x, None)
    y0 = E.run_gru(nhwc(x), None, ru, out, kind_id(kind))
    assert (y0.cpu() - ref0).abs().max().item() <= TOL[kind] * 3


@pytest.mark.parametrize('kind', KINDS)
@pytest.mark.parametrize('C,H,W,N', [(64, 16, 16, 1), (128, 8, 12, 2), (32, 10, 6, 2)])
def test_convlstm_fused(kind, C, H, W, N):
    """One fused convolution with gate-interleaved columns == ConvLSTM.forward (submodules.py:318-358)."""
    from rpg_ramnet_b200 import engine as E, ops
    from rpg_ramnet_b200.model.submodules import ConvLSTM
    torch.manual_seed(C + W)
    lstm = ConvLSTM(C, C, 3)
    sd = {'l.' + k: v.detach() for k, v in lstm.state_dict().items()}
    x, h, c = _rand((N, C, H, W), 1), _rand((N, C, H, W), 2, 0.7), _rand((N, C, H, W), 3)
    rh, rc = O.conv_lstm(sd, 'l', x, (h, c))
    lstm.to(dev())
    p = E.pack_lstm(E.WeightCache(), 'l', lstm, kind_id(kind))
    yh, yc = E.run_lstm(nhwc(x), (nhwc(h), nhwc(c)), p, kind_id(kind))
    assert (yh.cpu() - rh).abs().max().item() <= TOL[kind] * 3
    assert (yc.cpu() - rc).abs().max().item() <= TOL[kind] * 3
    rh0, rc0 = O.conv_lstm(sd, 'l', x, None)
    yh0, yc0 = E.run_lstm(nhwc(x), None, p, kind_id(kind))
    assert (yh0.cpu() - rh0).abs().max().item() <= TOL[kind] * 3


@pytest.mark.parametrize('Cin,H,W,N', [(1, 32, 48, 2), (5, 40, 33, 1), (6, 8, 8, 3), (5, 256, 512, 1)])
def test_head_conv(Cin, H, W, N):
    from rpg_ramnet_b200 import ops
    x = _rand((N, Cin, H, W), 5)
    w, b = _rand((32, Cin, 5, 5), 6, 0.2), _rand((32,), 7, 0.1)
    ref = torch.relu(F.conv2d(x, w, b, padding=2))
    y = ops.head_conv(x.to(dev()), w.to(dev()), b.to(dev()), round_tf32=False)
    assert ops._is_nhwc(y)
    assert (y.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('Cin,H,W,N', [(1, 32, 48, 2), (5, 40, 33, 1), (6, 8, 8, 3), (5, 256, 512, 1)])
def test_head_conv_tensor_core(Cin, H, W, N):
    """im2row (5 horizontal taps -> channels) + 5x1 tcgen05 conv vs torch on TF32-rounded operands."""
    from rpg_ramnet_b200 import ops
    x = _rand((N, Cin, H, W), 5)
    w, b = _rand((32, Cin, 5, 5), 6, 0.2), _rand((32,), 7, 0.1)
    def rna(t):      # cvt.rna.tf32.f32 on the host: round to nearest, ties away from zero, 10-bit mantissa
        b = t.contiguous().view(torch.int32)
        return ((b + 0x1000) & ~0x1fff).view(torch.float32)
    xr, wr = rna(x), rna(w)
    xe = ops.head_im2row(x.to(dev()))
    # the unrolled tensor: channel dx*Cin + ci = x[ci] shifted by dx - 2, zero outside / beyond 5*Cin
    pad = F.pad(xr, (2, 2))
    for dx in range(5):
        for ci in range(Cin):
            assert torch.equal(xe[:, dx * Cin + ci].cpu(), pad[:, ci, :, dx:dx + W])
    assert float(xe[:, 5 * Cin:].abs().max()) == 0.0
    ref = torch.relu(F.conv2d(xr.double(), wr.double(), b.double(), padding=2)).float()
    y = ops.head_conv_tc(xe, ops.pack_weights_head(w.to(dev())), b.to(dev()), Cin, 32, round_tf32=False)
    assert ops._is_nhwc(y)
    assert (y.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('with_skip', [False, True])
@pytest.mark.parametrize('N,C,H,W', [(1, 32, 8, 8), (2, 64, 5, 7), (1, 256, 32, 43), (1, 4, 1, 1)])
def test_upsample2x_add(with_skip, N, C, H, W):
    from rpg_ramnet_b200 import ops
    x, s = _rand((N, C, H, W), 1), _rand((N, C, H, W), 2)
    ref = F.interpolate(x + s if with_skip else x, scale_factor=2, mode='bilinear', align_corners=False)
    y = ops.upsample2x_add(nhwc(x), nhwc(s) if with_skip else None, round_tf32=False)
    assert y.shape == ref.shape
    assert (y.cpu() - ref).abs().max().item() <= 1e-6 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize('with_skip', [False, True])
def test_pred_sigmoid(with_skip):
    from rpg_ramnet_b200 import ops
    N, C, H, W = 2, 32, 17, 23
    x, s = _rand((N, C, H, W), 1), _rand((N, C, H, W), 2)
    w, b = _rand((1, C, 1, 1), 3, 0.3), _rand((1,), 4)
    logits_ref = F.conv2d(x + s if with_skip else x, w, b)
    d, l = ops.pred_sigmoid(nhwc(x), nhwc(s) if with_skip else None, w.to(dev()), b.to(dev()), want_logits=True)
    assert d.shape == (N, 1, H, W)
    assert (l.cpu() - logits_ref).abs().max().item() <= 1e-5
    assert (d.cpu() - torch.sigmoid(logits_ref)).abs().max().item() <= 1e-6


def test_pred_concat_forward_backward():
    """skip_type 'concat' (unet.py:11-12,129): 1x1 conv over cat([x, skip]) without the concatenated tensor, and its adjoint."""
    from rpg_ramnet_b200 import autograd as AG, ops
    N, C, H, W = 2, 32, 17, 23
    x, s = _rand((N, C, H, W), 1), _rand((N, C, H, W), 2)
    w, b, gd = _rand((1, 2 * C, 1, 1), 3, 0.3), _rand((1,), 4), _rand((N, 1, H, W), 5)
    tx, ts = x.double().requires_grad_(True), s.double().requires_grad_(True)
    tw, tb = w.double().requires_grad_(True), b.double().requires_grad_(True)
    ref = torch.sigmoid(F.conv2d(torch.cat([tx, ts], 1), tw, tb))
    ref.backward(gd.double())
    gx, gs = nhwc(x).requires_grad_(True), nhwc(s).requires_grad_(True)
    gw, gb = w.to(dev()).requires_grad_(True), b.to(dev()).requires_grad_(True)
    d = AG.PredFn.apply(gx, gs, gw, gb, True)
    assert (d.cpu().double() - ref.detach()).abs().max().item() <= 1e-6
    d.backward(gd.to(dev()))
    for got, want in ((gx.grad, tx.grad), (gs.grad, ts.grad), (gw.grad, tw.grad), (gb.grad, tb.grad)):
        assert float((got.cpu().double() - want).abs().max()) <= 1e-4 * max(1.0, float(want.abs().max()))
    l = ops.pred_logits(nhwc(x), nhwc(s), w.to(dev()), b.to(dev()), concat=True)
    assert (l.cpu().double() - F.conv2d(torch.cat([x, s], 1).double(), w.double(), b.double())).abs().max().item() <= 1e-5


def test_layout_roundtrip_and_tf32_rounding():
    from rpg_ramnet_b200 import _lib, ops
    x = _rand((2, 37, 9, 13), 8)
    a = ops.as_nhwc(x.to(dev()))
    assert torch.equal(a.cpu(), x)
    assert torch.equal(ops.to_nchw_contiguous(a).cpu(), x)
    import ctypes
    xd = x.to(dev()).contiguous()
    y = torch.empty_like(xd)
    _lib.check(_lib.load().ramnet_round_tf32(_lib.handle(0), ctypes.c_void_p(xd.data_ptr()),
                                             ctypes.c_void_p(y.data_ptr()), xd.numel(), None))
    bits = y.cpu().view(torch.int32)
    assert int((bits & 0x1fff).abs().max()) == 0
    assert (y.cpu() - x).abs().max().item() <= 2 ** -11 * x.abs().max().item()


# ----------------------------------------------------------------------------- loss / optimiser
def test_si_loss_golden_and_autograd():
    import rpg_ramnet_b200 as R
    g = np.load(os.path.join(GOLDEN, 'si_loss.npz'))
    for i in range(4):
        lam, w = (float(v) for v in g[f'{i}/params'])
        p = torch.from_numpy(g[f'{i}/pred']).to(dev()).requires_grad_(True)
        t = torch.from_numpy(g[f'{i}/target']).to(dev())
        loss = R.scale_invariant_loss(p, t, w, lam)
        assert abs(loss.item() - float(g[f'{i}/loss'])) <= 1e-6
        (3.0 * loss).backward()
        np.testing.assert_allclose(p.grad.cpu().numpy(), 3.0 * g[f'{i}/grad'], rtol=1e-4, atol=1e-9)


def test_si_loss_large_vs_oracle():
    import rpg_ramnet_b200 as R
    gen = torch.Generator().manual_seed(3)
    p, t = torch.rand(4, 1, 256, 512, generator=gen), torch.rand(4, 1, 256, 512, generator=gen)
    t[:, :, 3:40, 100:300] = float('nan')
    loss = R.scale_invariant_loss(p.to(dev()), t.to(dev()), 1.0, 1.0)
    assert abs(loss.item() - O.si_loss(p.double(), t.double()).item()) <= 1e-6


@pytest.mark.parametrize('shape', [(2, 1, 64, 96), (4, 1, 256, 512), (1, 1, 8, 8)])
@pytest.mark.parametrize('nan_patch', [False, True])
def test_multi_scale_grad_loss_vs_oracle(nan_patch, shape):
    """Value and gradient vs the oracle restatement + torch autograd (kornia semantics are restated, see oracle docstring).
    Shapes: a small rectangle, the bench shape, and the smallest legal map (the coarsest scale is ONE pooled pixel: every
    Sobel tap is a replicate-clamped copy of it)."""
    import rpg_ramnet_b200 as R
    gen = torch.Generator().manual_seed(9)
    p, t = torch.rand(shape, generator=gen), torch.rand(shape, generator=gen)
    if nan_patch and shape[2] >= 64:
        t[0, :, 5:23, 40:61] = float('nan')
        t[-1, :, 0:3, 0:9] = float('nan')
    elif nan_patch:
        t[0, :, 0, 0] = float('nan')
    pr = p.clone().requires_grad_(True)
    ref = O.multi_scale_grad_loss(pr, t)
    ref.backward()
    pg = p.to(dev()).requires_grad_(True)
    loss = R.multi_scale_grad_loss(pg, t.to(dev()))
    if torch.isnan(ref):        # 8x8 with a NaN: the coarsest scale has no valid gradient pixel, 0 / 0 in the reference too
        assert torch.isnan(loss).item()
        return
    assert abs(loss.item() - ref.item()) <= 2e-6 * max(1.0, abs(ref.item()))
    (2.5 * loss).backward()
    got, want = pg.grad.cpu(), 2.5 * pr.grad
    assert float((got - want).abs().max()) <= 1e-6 + 1e-4 * float(want.abs().max())
    assert float(got[torch.isnan(t)].abs().max() if nan_patch else 0.0) == 0.0


def test_adam_golden_trajectory():
    from rpg_ramnet_b200 import ops
    g = np.load(os.path.join(GOLDEN, 'adam.npz'))
    p = torch.from_numpy(g['p0'].copy()).to(dev())
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for s in range(3):
        ops.adam_step(p, torch.from_numpy(g[f'g{s}']).to(dev()), m, v, s + 1, lr=3e-4)
        np.testing.assert_allclose(p.cpu().numpy(), g[f'p{s + 1}'], rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(m.cpu().numpy(), g['m'], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(v.cpu().numpy(), g['v'], rtol=1e-5, atol=1e-12)


def test_conv_argument_errors():
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200 import ops
    x = nhwc(_rand((1, 24, 8, 8), 1))          # 24 channels: not a multiple of 16
    w = torch.zeros(32 * 24 * 9, device=dev())
    with pytest.raises(R.RamnetError):
        ops.conv_fwd(x, None, w, None, 32, 3, 1, ops.EPI_BIAS, ops.MMA_FP32)
    x = nhwc(_rand((1, 32, 8, 8), 1))
    with pytest.raises(R.RamnetError):
        ops.conv_fwd(x, None, w, None, 32, 7, 1, ops.EPI_BIAS, ops.MMA_FP32)      # ksize 7
    with pytest.raises(R.RamnetError):
        ops.conv_fwd(x, None, w, None, 32, 3, 1, ops.EPI_GRU_OUT, ops.MMA_FP32)   # missing aux
    with pytest.raises(R.RamnetError):
        ops.conv_fwd(x.cpu(), None, w, None, 32, 3, 1, ops.EPI_BIAS, ops.MMA_FP32)  # CPU tensor: no fallback


# ----------------------------------------------------------------------------- loader wire format / trainer metrics
def test_voxel_normalize_label_and_metrics_on_device():
    """device twins of the loader normalisation, the log-depth label transform and the trainer's
    metrics vs the reference's own outputs (tests/golden/dataio.npz) and the numpy oracle."""
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200.model import metric as M
    from rpg_ramnet_b200.utils.event_tensor_utils import depth_to_log_label, normalize_voxel_grid
    from oracle import dataio_oracle as D
    g = np.load(os.path.join(GOLDEN, 'dataio.npz'))
    cases = D.synth_cases(0)
    for k in ('vox_sparse', 'vox_zero', 'vox_const'):
        x = torch.from_numpy(cases[k].copy()).to(dev())
        y = normalize_voxel_grid(x)
        assert y.data_ptr() == x.data_ptr()                                   # in place, like the reference
        ref = g[k + '_numpy']
        np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-5, atol=2e-6)    # float64 vs numpy-float32 statistics
        np.testing.assert_array_equal(y.cpu().numpy() == 0, ref == 0)
    big = (torch.randn(5, 256, 512, generator=torch.Generator().manual_seed(3)) *
           (torch.rand(5, 256, 512, generator=torch.Generator().manual_seed(4)) < 0.1)).float()
    np.testing.assert_allclose(normalize_voxel_grid(big.clone().to(dev())).cpu().numpy(),
                               D.normalize_voxel_grid(big.numpy()), rtol=1e-4, atol=1e-5)
    batch = torch.stack([big, big * 2.0 + (big != 0).float(), torch.zeros_like(big)]).to(dev())
    out = normalize_voxel_grid(batch.clone()).cpu().numpy()             # per-sample statistics in one launch pair
    for i in range(3):
        np.testing.assert_allclose(out[i], D.normalize_voxel_grid(batch[i].cpu().numpy()), rtol=1e-4, atol=1e-5)
    for clip, reg in ((80.0, 3.70378), (1000.0, 6.2044)):
        lab = depth_to_log_label(torch.from_numpy(cases['depth']).to(dev()), clip, reg).cpu().numpy()
        ref = g[f'label_{int(clip)}']
        assert np.array_equal(np.isnan(lab), np.isnan(ref))
        np.testing.assert_allclose(np.nan_to_num(lab), np.nan_to_num(ref), rtol=0, atol=2e-7 * 8)   # logf vs np.log: a few ulp
    p, t = torch.from_numpy(cases['metric_pred']).to(dev()), torch.from_numpy(cases['metric_target']).to(dev())
    names = ['mse', 'abs_rel_diff', 'scale_invariant_error', 'median_error', 'squ_rel_diff', 'rms_linear', 'mean_error']
    vals = M.eval_metrics(p, t, names)
    for n, v in zip(names, vals):
        np.testing.assert_allclose(v, float(g['metric_' + n]), rtol=2e-5, err_msg=n)
        np.testing.assert_allclose(getattr(M, n)(p, t), v, rtol=1e-12)


# ----------------------------------------------------------------------------- live normalisation layers
@pytest.mark.parametrize('kind,act,res', [('BN', 'relu', False), ('BN', 'relu', True), ('IN', 'relu', False),
                                          ('IN', None, True), ('BN', 'sigmoid', False), ('BN', None, False)])
@pytest.mark.parametrize('shape', [(2, 32, 9, 11), (3, 256, 8, 8), (4, 64, 128, 256), (2, 1, 16, 24), (1, 128, 5, 3)])
def test_norm_fwd_bwd_vs_torch(kind, act, res, shape):
    """ramnet_norm_fwd / ramnet_norm_bwd (train-mode BatchNorm2d / InstanceNorm2d + activation + residual) against
    torch CPU float64 F.batch_norm / F.instance_norm: output, running statistics (momentum, unbiased variance), and
    all gradients (dz, dres, dgamma, dbeta).  Shapes cover the float4 path, C = 1 (the pred layer) and a 8 M-element
    tensor that makes every block walk several grid strides."""
    from rpg_ramnet_b200 import ops
    N, C, H, W = shape
    g = torch.Generator().manual_seed(N * 1000 + C + H)
    z = torch.randn(shape, generator=g) * 1.7 + 0.4 * torch.randn(1, C, 1, 1, generator=g)
    r = torch.randn(shape, generator=g) if res else None
    affine = kind == 'BN'
    gamma = (0.75 + 0.5 * torch.rand(C, generator=g)) if affine else None
    beta = (0.1 * torch.randn(C, generator=g)) if affine else None
    rm0, rv0 = 0.1 * torch.randn(C, generator=g), 0.5 + torch.rand(C, generator=g)
    dy = torch.randn(shape, generator=g)
    # reference in float64
    zd = z.double().requires_grad_(True)
    rd = None if r is None else r.double().requires_grad_(True)
    gd = None if gamma is None else gamma.double().requires_grad_(True)
    bd = None if beta is None else beta.double().requires_grad_(True)
    rm_ref, rv_ref = rm0.double().clone(), rv0.double().clone()
    if kind == 'BN':
        t = F.batch_norm(zd, rm_ref, rv_ref, gd, bd, True, 0.1, 1e-5)
    else:
        t = F.instance_norm(zd, rm_ref, rv_ref, None, None, True, 0.1, 1e-5)
    if rd is not None:
        t = t + rd
    yr = torch.relu(t) if act == 'relu' else (torch.sigmoid(t) if act == 'sigmoid' else t)
    yr.backward(dy.double())
    # ours
    rm, rv = rm0.to(dev()), rv0.to(dev())
    zc = nhwc(z)
    y, stats = ops.norm_fwd(zc, kind, act, None if gamma is None else gamma.to(dev()), None if beta is None else beta.to(dev()),
                            None if r is None else nhwc(r), rm, rv, 0.1, 1e-5, True, False)
    assert float((y.cpu().double() - yr.detach()).abs().max()) <= 2e-5
    assert float((rm.cpu().double() - rm_ref).abs().max()) <= 1e-6
    assert float((rv.cpu().double() - rv_ref).abs().max()) <= 1e-5 * float(rv_ref.abs().max())
    dgam = torch.zeros(C, device=dev()) if affine else None
    dbet = torch.zeros(C, device=dev()) if affine else None
    dz, dres = ops.norm_bwd(nhwc(dy), y, zc, stats, kind, act, None if gamma is None else gamma.to(dev()), True, False,
                            res, dgam, dbet)
    scale = float(zd.grad.abs().max())
    # relu masks are taken from the fp32 forward: entries within float rounding of zero may flip (none at this tolerance)
    assert float((dz.cpu().double() - zd.grad).abs().max()) <= 5e-5 * max(scale, 1.0)
    if res:
        assert float((dres.cpu().double() - rd.grad).abs().max()) <= 1e-6 * max(1.0, float(rd.grad.abs().max()))
    if affine:
        assert float((dgam.cpu().double() - gd.grad).abs().max()) <= 2e-5 * max(1.0, float(gd.grad.abs().max()))
        assert float((dbet.cpu().double() - bd.grad).abs().max()) <= 2e-5 * max(1.0, float(bd.grad.abs().max()))


def test_norm_running_mode_matches_eval_batchnorm():
    """RAMNET_NORM_RUNNING: an eval-mode BatchNorm2d that gradients flow through (running statistics, nothing updated)."""
    from rpg_ramnet_b200 import ops
    N, C, H, W = 2, 64, 12, 20
    g = torch.Generator().manual_seed(5)
    z, dy = torch.randn(N, C, H, W, generator=g), torch.randn(N, C, H, W, generator=g)
    gamma, beta = 0.75 + 0.5 * torch.rand(C, generator=g), 0.1 * torch.randn(C, generator=g)
    rm0, rv0 = 0.1 * torch.randn(C, generator=g), 0.5 + torch.rand(C, generator=g)
    zd, gd, bd = z.double().requires_grad_(True), gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    yr = torch.relu(F.batch_norm(zd, rm0.double(), rv0.double(), gd, bd, False, 0.1, 1e-5))
    yr.backward(dy.double())
    rm, rv = rm0.to(dev()), rv0.to(dev())
    y, stats = ops.norm_fwd(nhwc(z), 'BN', 'relu', gamma.to(dev()), beta.to(dev()), None, rm, rv, 0.1, 1e-5, False, False)
    assert torch.equal(rm.cpu(), rm0) and torch.equal(rv.cpu(), rv0)
    assert float((y.cpu().double() - yr.detach()).abs().max()) <= 1e-5
    dgam, dbet = torch.zeros(C, device=dev()), torch.zeros(C, device=dev())
    dz, _ = ops.norm_bwd(nhwc(dy), y, nhwc(z), stats, 'BN', 'relu', gamma.to(dev()), False, False, False, dgam, dbet)
    assert float((dz.cpu().double() - zd.grad).abs().max()) <= 1e-5
    assert float((dgam.cpu().double() - gd.grad).abs().max()) <= 2e-5 * float(gd.grad.abs().max())
    assert float((dbet.cpu().double() - bd.grad).abs().max()) <= 2e-5 * float(bd.grad.abs().max())
