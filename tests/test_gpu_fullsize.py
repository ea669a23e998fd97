"""GPU: parity at the BENCHED sizes.  The 128x128 goldens leave most layers with fewer work items
than SMs, so they never exercise what produces the headline number: persistent CTAs walking many items, the epilogue
of item i under the MMAs of item i+1 (TMEM double buffering), CTA-pair mode on a 148-CTA grid with an odd patch count,
CUDA-graph replay with the H2D staging ring.  Here the shipped block runs at BASELINE configs[1] (256x512, batch 4),
configs[3] (256x344, batch 1, irregular schedule) and configs[2] (one L=2 training step at 256x512, batch 4) against
the CPU oracle on the same seeded inputs, and every fused epilogue is checked on >= 64k pixels against torch fp64.
Tolerance: north_star's 1e-3 relative on depth maps (TF32 operands, fp32 accumulate)."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ramnet_oracle as O
from helpers import max_rel_err

pytestmark = pytest.mark.gpu

CFG = dict(num_bins_rgb=1, num_bins_events=5, skip_type='sum', recurrent_block_type='conv',
           state_combination='convgru', num_encoders=3, base_num_channels=32, num_residual_blocks=2,
           use_upsample_conv=True, norm='none', every_x_rgb_frame=1, gpu=0)


def dev():
    return torch.device('cuda', 0)


def build(cfg, train=False):
    import rpg_ramnet_b200 as R
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.ERGB2DepthRecurrent(cfg)
    m = m.train() if train else m.eval()
    return m.to(dev())


def nhwc(t):
    return t.to(dev()).contiguous(memory_format=torch.channels_last)


def rna(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)


@pytest.mark.parametrize('graphs', [False, True])
def test_config2_shape_two_timesteps_vs_oracle(graphs):
    """BASELINE configs[1]: 256x512, batch 4, K=1 — two timesteps (16 depth maps) with state carry, eager and as CUDA-graph
    replays fed from pinned host memory (the bench's e2e path), against the oracle."""
    model = build(dict(CFG, cuda_graphs=graphs))
    seq = O.synth_sequence(4, 256, 512, 2, 1, seed=2, with_targets=False)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    if graphs:
        seq_in = [{k: v.pin_memory() for k, v in it.items()} for it in seq]
    else:
        seq_in = seq
    prev_s, prev_l = None, {'events0': None, 'image': None}
    o_s, o_l = None, {'events0': None, 'image': None}
    worst = 0.0
    with torch.no_grad():
        for item_in, item in zip(seq_in, seq):
            preds, supers, lstm = model(item_in, prev_s, prev_l)
            o_preds, o_supers, o_lstm = O.ergb2depth_recurrent(sd, CFG, item, o_s, o_l)
            for k in o_preds:
                assert preds[k].shape == (4, 1, 256, 512)
                worst = max(worst, max_rel_err(preds[k].cpu().numpy(), o_preds[k].numpy()))
            for a, b in zip(supers['image'], o_supers['image']):
                scale = max(1e-3, float(b.abs().max()))
                assert float((a.cpu() - b).abs().max()) <= 3e-3 * scale
            prev_s, prev_l = supers['image'], lstm
            o_s, o_l = o_supers['image'], o_lstm
    assert worst <= 1e-3, f'256x512 B=4 graphs={graphs}: max rel err {worst:.3e}'


def test_config4_shape_irregular_schedule_vs_oracle():
    """BASELINE configs[3]: MVSEC crop 256x344, batch 1, a variable number of event passes between frames, through the
    inner API (forward_events / forward_images / forward_decoder) the irregular branch would drive."""
    model = build(CFG)
    net = model.statenetphasedrecurrent
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ocfg = O.NetCfg(dict(CFG))
    gen = torch.Generator().manual_seed(3)
    Hh, Ww = 256, 344
    sched = [3, 1, 2]                       # event passes before each frame
    s, s_ref = model._zero_states(1, Hh, Ww), O.zero_super_states(ocfg, 1, Hh, Ww)
    worst, n = 0.0, 0
    with torch.no_grad():
        for n_ev in sched:
            for which in ['e'] * n_ev + ['i']:
                if which == 'e':
                    x = torch.randn(1, 5, Hh, Ww, generator=gen) * (torch.rand(1, 5, Hh, Ww, generator=gen) < 0.1)
                    s, _ = net.forward_events(x.to(dev()), s, None, None)
                    s_ref, _ = O.forward_events(sd, ocfg, x, s_ref, None)
                else:
                    x = torch.rand(1, 1, Hh, Ww, generator=gen)
                    s, _ = net.forward_images(x.to(dev()), s, None, None)
                    s_ref, _ = O.forward_images(sd, ocfg, x, s_ref, None)
                d, d_ref = net.forward_decoder(s), O.forward_decoder(sd, ocfg, s_ref)
                worst = max(worst, max_rel_err(d.cpu().numpy(), d_ref.numpy()))
                n += 1
    assert n == sum(sched) + len(sched)
    assert worst <= 1e-3, f'256x344 irregular schedule: max rel err {worst:.3e}'


def _conv_ref(x0, x1, w, b, k):
    x = x0 if x1 is None else torch.cat([x0, x1], 1)
    return F.conv2d(x.double(), w.double(), b.double(), padding=k // 2)


@pytest.mark.parametrize('C,H,W,N', [(64, 128, 256, 2), (128, 64, 128, 4), (256, 32, 64, 4)])
def test_gru_epilogues_at_bench_size(C, H, W, N):
    """EPI_GRU_RU + EPI_GRU_OUT (submodules.py:436-454) on >= 64k (level 0/1) and the level-2 8192-pixel shape:
    multi-item persistent CTAs, pair mode, TMEM double buffering."""
    from rpg_ramnet_b200 import ops
    g = torch.Generator().manual_seed(C + H)
    x, h = rna(torch.randn(N, C, H, W, generator=g)), rna(torch.randn(N, C, H, W, generator=g) * 0.5)
    sc = (1.0 / (2 * C * 9)) ** 0.5
    w_ru, b_ru = rna(torch.randn(2 * C, 2 * C, 3, 3, generator=g) * sc), torch.randn(2 * C, generator=g) * 0.1
    w_o, b_o = rna(torch.randn(C, 2 * C, 3, 3, generator=g) * sc), torch.randn(C, generator=g) * 0.1
    z = _conv_ref(x, h, w_ru, b_ru, 3)
    r, u = torch.sigmoid(z[:, :C]), torch.sigmoid(z[:, C:])
    xd, hd = nhwc(x), nhwc(h)
    u_g, rh_g = ops.conv_fwd(xd, hd, ops.pack_weights(w_ru.to(dev()), ops.MMA_TF32), b_ru.to(dev()), 2 * C, 3, 1,
                             ops.EPI_GRU_RU, ops.MMA_TF32, aux0=hd, round_tf32=True)
    assert float((u_g.cpu().double() - u).abs().max()) <= 1e-3
    assert float((rh_g.cpu().double() - h.double() * r).abs().max()) <= 2e-3
    # OUT on the device's own (TF32-rounded) r*h so that the check isolates the second kernel
    rh_in = rh_g.cpu()
    o = torch.tanh(_conv_ref(x, rh_in, w_o, b_o, 3))
    u_in = u_g.cpu().double()
    hn = h.double() * (1 - u_in) + o * u_in
    hn_g = ops.conv_fwd(xd, rh_g, ops.pack_weights(w_o.to(dev()), ops.MMA_TF32), b_o.to(dev()), C, 3, 1,
                        ops.EPI_GRU_OUT, ops.MMA_TF32, aux0=hd, aux1=u_g, round_tf32=False)
    assert float((hn_g.cpu().double() - hn).abs().max()) <= 1e-3


def test_lstm_epilogue_at_bench_size():
    """EPI_LSTM (submodules.py:318-358), C=64 at 128x256, batch 2 = 65536 pixels, N = 256 gate columns."""
    from rpg_ramnet_b200 import ops
    C, H, W, N = 64, 128, 256, 2
    g = torch.Generator().manual_seed(11)
    x, h = rna(torch.randn(N, C, H, W, generator=g)), rna(torch.randn(N, C, H, W, generator=g) * 0.5)
    c = torch.randn(N, C, H, W, generator=g) * 0.5
    w = rna(torch.randn(4 * C, 2 * C, 3, 3, generator=g) * (1.0 / (2 * C * 9)) ** 0.5)
    b = torch.randn(4 * C, generator=g) * 0.1
    z = _conv_ref(x, h, w, b, 3)
    i_, f_, o_, g_ = z.chunk(4, 1)                         # in, remember, out, cell (submodules.py:344)
    cn = torch.sigmoid(f_) * c.double() + torch.sigmoid(i_) * torch.tanh(g_)
    hn = torch.sigmoid(o_) * torch.tanh(cn)
    bp = b.view(4, C).t().contiguous().view(-1)
    hn_g, cn_g = ops.conv_fwd(nhwc(x), nhwc(h), ops.pack_weights(w.to(dev()), ops.MMA_TF32, lstm_interleave=True),
                              bp.to(dev()), 4 * C, 3, 1, ops.EPI_LSTM, ops.MMA_TF32, aux0=nhwc(c), round_tf32=False)
    assert float((cn_g.cpu().double() - cn).abs().max()) <= 1.5e-3
    assert float((hn_g.cpu().double() - hn).abs().max()) <= 1.5e-3


@pytest.mark.parametrize('epi', ['res', 'pred'])
def test_res_and_pred_epilogues_at_bench_size(epi):
    """EPI_BIAS_RES_RELU on the level-2 resblock shape at batch 32 (65536 pixels) and the fused last decoder + pred +
    sigmoid (EPI_BIAS_RELU_PRED) at 256x512 (131072 pixels, hpack path)."""
    from rpg_ramnet_b200 import ops
    g = torch.Generator().manual_seed(5)
    if epi == 'res':
        N, C, H, W, Co, k = 32, 256, 32, 64, 256, 3
    else:
        N, C, H, W, Co, k = 1, 64, 256, 512, 32, 5
    x = rna(torch.randn(N, C, H, W, generator=g))
    w = rna(torch.randn(Co, C, k, k, generator=g) * (1.0 / (C * k * k)) ** 0.5)
    b = torch.randn(Co, generator=g) * 0.1
    y = _conv_ref(x, None, w, b, k)
    if epi == 'res':
        res = torch.randn(N, Co, H, W, generator=g)
        ref = torch.relu(y + res.double())
        out = ops.conv_fwd(nhwc(x), None, ops.pack_weights(w.to(dev()), ops.MMA_TF32), b.to(dev()), Co, k, 1,
                           ops.EPI_BIAS_RES_RELU, ops.MMA_TF32, aux0=nhwc(res))
        assert float((out.cpu().double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    else:
        pw, pb = torch.randn(Co, generator=g) * 0.3, torch.randn(1, generator=g)
        ref = torch.sigmoid((torch.relu(y) * pw.double().view(1, -1, 1, 1)).sum(1, keepdim=True) + pb.double())
        wd = w.to(dev())
        wp = ops.pack_weights_hpack(wd) if ops.hpack_eligible(Co, k, 1, ops.MMA_TF32) else ops.pack_weights(wd, ops.MMA_TF32)
        out = ops.conv_fwd(nhwc(x), None, wp, b.to(dev()), Co, k, 1, ops.EPI_BIAS_RELU_PRED, ops.MMA_TF32,
                           aux0=pw.to(dev()), aux1=pb.to(dev()))
        assert float((out.cpu().double() - ref).abs().max()) <= 2e-5


def test_config3_shape_training_step_loss_and_grads_vs_oracle():
    """BASELINE configs[2] per GPU at L=2: 256x512, batch 4, SI loss on events0+image with the trainer's K_keys aliasing,
    full BPTT.  Loss within 5e-4 of the fp32 CPU oracle; gradient norms of a few tensors against torch autograd on
    the oracle graph."""
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200.model.loss import SILossBatch
    model = build(CFG, train=True)
    seq = O.synth_sequence(4, 256, 512, 2, 1, seed=2, with_targets=True)
    keys = ['events0', 'image']
    L = len(seq)
    prev_s, prev_l = None, {'events0': None, 'image': None}
    batch = SILossBatch(L * len(keys), dev())
    for item in seq:
        preds, supers, lstm = model(item, prev_s, prev_l)
        for k in keys:
            batch.add(preds[k], item['depth_' + k].to(dev()), 1.0, 1.0)
        prev_s, prev_l = supers['image'], lstm
    loss = len(keys) * batch.finish().sum() / float(L)
    loss.backward()
    # oracle: same graph with torch autograd on the CPU
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    o_s, o_l = None, {'events0': None, 'image': None}
    terms = []
    for item in seq:
        o_preds, o_supers, o_lstm = O.ergb2depth_recurrent(sd, CFG, item, o_s, o_l)
        for k in keys:
            terms.append(O.si_loss(o_preds[k], item['depth_' + k], 1.0, 1.0))
        o_s, o_l = o_supers['image'], o_lstm
    ref = len(keys) * sum(terms) / float(L)
    assert abs(loss.item() - ref.item()) <= 5e-4, (loss.item(), ref.item())
    ref.backward()
    named = dict(model.named_parameters())
    for name in ['statenetphasedrecurrent.pred.conv2d.weight', 'statenetphasedrecurrent.decoders.2.conv2d.weight',
                 'statenetphasedrecurrent.resblocks.0.conv1.weight',
                 'statenetphasedrecurrent.state_combination_events.0.recurrent_block.out_gate.weight',
                 'statenetphasedrecurrent.encoders_rgb.1.conv2d.weight', 'statenetphasedrecurrent.head_events.conv2d.weight']:
        g_ours, g_ref = named[name].grad.detach().cpu().double(), sd[name].grad.double()
        rel = float((g_ours - g_ref).norm() / g_ref.norm().clamp_min(1e-30))
        assert rel <= 6e-2, f'{name}: relative Frobenius error {rel:.3e}'      # TF32 vs fp32 oracle; measured worst 4.4e-2


@pytest.mark.parametrize('layer', ['gru0_ru', 'gru0_out', 'enc0_s2seg', 'dec2_upconv_pred', 'dec1_upconv', 'lstm0', 'res_single_wave'])
def test_dynamic_work_distribution_is_bit_identical(layer):
    """RAMNET_FLAG_DYNAMIC (items drawn from a global counter through the shared-memory ring, engine.GraphRunner's overlapped
    launches) against the static round robin on multi-wave layers of the bench shape: bit-identical outputs, and the
    counter pair resets itself (the same 8-byte slot serves five launches in a row and reads zero afterwards)."""
    from rpg_ramnet_b200 import ops
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(5)
    B = 4

    def rnd(n, c, h, w, scale=1.0):
        return (torch.randn(n, c, h, w, generator=g) * scale).to(dev).contiguous(memory_format=torch.channels_last)

    kind = ops.MMA_TF32
    if layer in ('gru0_ru', 'gru0_out', 'lstm0'):
        C, H, W = 64, 128, 256
        x, h = rnd(B, C, H, W), rnd(B, C, H, W)
        cout = {'gru0_ru': 2 * C, 'gru0_out': C, 'lstm0': 4 * C}[layer]
        epi = {'gru0_ru': ops.EPI_GRU_RU, 'gru0_out': ops.EPI_GRU_OUT, 'lstm0': ops.EPI_LSTM}[layer]
        w = torch.randn(cout, 2 * C, 3, 3, generator=g).to(dev) * 0.05
        wp = ops.pack_weights(w, kind, lstm_interleave=(layer == 'lstm0'))
        b = torch.randn(cout, generator=g).to(dev) * 0.1
        aux1 = rnd(B, C, H, W).sigmoid() if layer == 'gru0_out' else None
        run = lambda: ops.conv_fwd(x, h, wp, b, cout, 3, 1, epi, kind, aux0=h, aux1=aux1, round_tf32=True)
    elif layer == 'enc0_s2seg':
        x = rnd(B, 32, 256, 512)
        w = torch.randn(64, 32, 5, 5, generator=g).to(dev) * 0.05
        wp, b = ops.pack_weights_s2seg(w), torch.randn(64, generator=g).to(dev) * 0.1
        run = lambda: ops.conv_fwd(x, None, wp, b, 64, 5, 2, ops.EPI_BIAS_RELU, kind, round_tf32=True)
    elif layer in ('dec2_upconv_pred', 'dec1_upconv'):
        cin, cout, H, W = (64, 32, 128, 256) if layer == 'dec2_upconv_pred' else (128, 64, 64, 128)
        x = rnd(B, cin, H, W)
        w = torch.randn(cout, cin, 5, 5, generator=g).to(dev) * 0.05
        wp, b = ops.pack_weights_upconv(w), torch.randn(cout, generator=g).to(dev) * 0.1
        if layer == 'dec2_upconv_pred':
            pw, pb = torch.randn(cout, generator=g).to(dev) * 0.3, torch.zeros(1, device=dev)
            run = lambda: ops.conv_up_fwd(x, wp, b, cout, ops.EPI_BIAS_RELU_PRED, aux0=pw, aux1=pb)
        else:
            run = lambda: ops.conv_up_fwd(x, wp, b, cout, ops.EPI_BIAS_RELU, round_tf32=True)
    else:   # one work item per worker: the launcher keeps the static assignment (nothing to redistribute)
        x = rnd(B, 256, 32, 64)
        w = torch.randn(256, 256, 3, 3, generator=g).to(dev) * 0.02
        wp, b = ops.pack_weights(w, kind), torch.zeros(256, device=dev)
        run = lambda: ops.conv_fwd(x, None, wp, b, 256, 3, 1, ops.EPI_BIAS_RELU, kind, round_tf32=True)

    def outs(r):
        return list(r) if isinstance(r, (tuple, list)) else [r]

    ref = [t.clone() for t in outs(run())]
    slot = torch.zeros(2, dtype=torch.int32, device=dev)
    ops.SCHED_SLOT_OVERRIDE = slot
    try:
        with ops.plan_flags(ops.FLAG_DYNAMIC):
            for rep in range(5):
                got = outs(run())
                for a, bb in zip(ref, got):
                    assert torch.equal(a, bb), (layer, rep)
                torch.cuda.synchronize()
                assert slot.tolist() == [0, 0], (layer, rep, slot.tolist())
    finally:
        ops.SCHED_SLOT_OVERRIDE = None
