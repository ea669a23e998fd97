"""GPU: up-conv mode of ramnet_conv_fwd (RAMNET_FLAG_UPCONV) against the reference's UpsampleConvLayer arithmetic
(F.interpolate bilinear x2, align_corners=False, then F.conv2d 5x5 pad 2; submodules.py:87-97) in fp64 on TF32-rounded
operands' nearest fp32 values.  The collapsed weights are rounded to TF32 AFTER the collapse, so the comparison carries
one extra operand rounding (2^-11 relative per weight) on top of the accumulation order: tolerance 2e-3 of the output
scale, and exact border handling is what the shapes probe (2x2 inputs, odd sizes, patches that straddle the border)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def nhwc(t):
    return t.to(dev()).contiguous(memory_format=torch.channels_last)


def rna(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)


def _ref(x, w, b):
    up = F.interpolate(x.double(), scale_factor=2, mode='bilinear', align_corners=False)
    return F.conv2d(up, w.double(), b.double(), padding=2)


SHAPES = [(1, 32, 32, 2, 2), (1, 32, 32, 3, 5), (2, 64, 32, 16, 24), (1, 64, 64, 17, 33), (1, 128, 64, 40, 24),
          (3, 32, 16, 9, 70), (1, 64, 32, 64, 128)]


@pytest.mark.parametrize('N,Cin,Cout,H,W', SHAPES)
@pytest.mark.parametrize('pair', ['1', '0'])
def test_upconv_relu_vs_interpolate_then_conv(N, Cin, Cout, H, W, pair, monkeypatch):
    from rpg_ramnet_b200 import ops
    g = torch.Generator().manual_seed(N + Cin + H + W)
    x = rna(torch.randn(N, Cin, H, W, generator=g))
    w = torch.randn(Cout, Cin, 5, 5, generator=g) * (1.0 / (Cin * 25)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    ref = torch.relu(_ref(x, w, b))
    wp = ops.pack_weights_upconv(w.to(dev()))
    out = ops.conv_up_fwd(nhwc(x), wp, b.to(dev()), Cout, ops.EPI_BIAS_RELU)
    assert tuple(out.shape) == (N, Cout, 2 * H, 2 * W)
    err = (out.cpu().double() - ref).abs()
    scale = max(1.0, float(ref.abs().max()))
    assert float(err.max()) <= 2e-3 * scale, (float(err.max()), scale)
    # the border ring is as accurate as the interior (a missing / wrong border segment shows up as O(0.1) errors there)
    ring = torch.ones_like(err, dtype=torch.bool)
    if H > 3 and W > 3:
        ring[:, :, 3:-3, 3:-3] = False
    assert float(err[ring].max()) <= 2e-3 * scale


def test_upconv_relu_add_forms_the_next_skip_sum():
    from rpg_ramnet_b200 import ops
    g = torch.Generator().manual_seed(7)
    N, Cin, Cout, H, W = 2, 64, 64, 12, 20
    x = rna(torch.randn(N, Cin, H, W, generator=g))
    w = torch.randn(Cout, Cin, 5, 5, generator=g) * (1.0 / (Cin * 25)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    skip = torch.randn(N, Cout, 2 * H, 2 * W, generator=g)
    ref = torch.relu(_ref(x, w, b)) + skip.double()
    out = ops.conv_up_fwd(nhwc(x), ops.pack_weights_upconv(w.to(dev())), b.to(dev()), Cout, ops.EPI_BIAS_RELU_ADD,
                          aux0=nhwc(skip), round_tf32=True)
    assert float((out.cpu().double() - ref).abs().max()) <= 3e-3 * max(1.0, float(ref.abs().max()))
    assert torch.equal(out.cpu(), rna(out.cpu()))                 # stored TF32-rounded for the next tensor-core layer


@pytest.mark.parametrize('N,Cin,H,W', [(1, 64, 8, 8), (2, 64, 24, 40), (1, 64, 128, 256)])
def test_upconv_fused_prediction_head(N, Cin, H, W):
    """Last decoder + 1x1 pred + sigmoid (statenet.py:116-117,313) from the low-resolution tensor: depth [N,1,2H,2W]."""
    from rpg_ramnet_b200 import ops
    g = torch.Generator().manual_seed(H + W)
    Cout = 32
    x = rna(torch.randn(N, Cin, H, W, generator=g))
    w = torch.randn(Cout, Cin, 5, 5, generator=g) * (1.0 / (Cin * 25)) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    pw, pb = torch.randn(Cout, generator=g) * 0.3, torch.randn(1, generator=g)
    t = torch.relu(_ref(x, w, b))
    logit = (t * pw.double().view(1, -1, 1, 1)).sum(1, keepdim=True) + pb.double()
    logits = torch.empty((N, 1, 2 * H, 2 * W), dtype=torch.float32, device=dev())
    depth = ops.conv_up_fwd(nhwc(x), ops.pack_weights_upconv(w.to(dev())), b.to(dev()), Cout, ops.EPI_BIAS_RELU_PRED,
                            aux0=pw.to(dev()), aux1=pb.to(dev()), out1=logits)
    assert tuple(depth.shape) == (N, 1, 2 * H, 2 * W)
    assert float((logits.cpu().double() - logit).abs().max()) <= 3e-3
    assert float((depth.cpu().double() - torch.sigmoid(logit)).abs().max()) <= 1e-3


def test_model_decoder_uses_upconv_and_matches_the_materialised_path(monkeypatch):
    """Whole model, same weights: up-conv decoders (default) vs RAMNET_UPCONV=0 (upsample2x_add + conv) and the oracle."""
    import contextlib
    import io
    import rpg_ramnet_b200 as R
    from oracle import ramnet_oracle as O
    cfg = dict(num_bins_rgb=1, num_bins_events=5, skip_type='sum', recurrent_block_type='conv', state_combination='convgru',
               num_encoders=3, base_num_channels=32, num_residual_blocks=2, use_upsample_conv=True, norm='none',
               every_x_rgb_frame=1, gpu=0)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = R.ERGB2DepthRecurrent(cfg)
    model.eval().to(dev())
    item = O.synth_sequence(2, 64, 96, 1, 1, seed=5, with_targets=False)[0]
    lstm = {'events0': None, 'image': None}
    with torch.no_grad():
        l0 = R.launch_count(0)
        a = model(item, None, lstm)[0]
        n_up = R.launch_count(0) - l0
        monkeypatch.setenv('RAMNET_UPCONV', '0')
        model.statenetphasedrecurrent._wcache.clear()
        l0 = R.launch_count(0)
        b = model(item, None, lstm)[0]
        n_old = R.launch_count(0) - l0
    assert n_up < n_old                              # the upsample launches of the up-conv decoders are gone
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = O.ergb2depth_recurrent(sd, cfg, item, None, lstm)[0]
    for k in ref:
        assert float(((a[k].cpu() - ref[k]).abs() / ref[k].abs()).max()) <= 1e-3
        assert float((a[k] - b[k]).abs().max()) <= 2e-4
