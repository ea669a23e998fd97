"""GPU: the boundary as the reference trainer uses it — the loss surface of model/loss.py, the loss mix of
LSTMTrainer.forward_pass_sequence with the multi-scale gradient term enabled, and FusedAdam as the
`torch.optim.Optimizer` that base/base_trainer.py builds, schedules, checkpoints and resumes."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ramnet_oracle as O

pytestmark = pytest.mark.gpu

CFG = dict(num_bins_rgb=1, num_bins_events=5, skip_type='sum', recurrent_block_type='conv',
           state_combination='convgru', num_encoders=3, base_num_channels=32, num_residual_blocks=2,
           use_upsample_conv=True, norm='none', every_x_rgb_frame=1, gpu=0)


def dev():
    return torch.device('cuda', 0)


def _pair(seed, shape=(2, 1, 32, 48), nan=True, positive=False):
    g = torch.Generator().manual_seed(seed)
    p, t = torch.rand(shape, generator=g), torch.rand(shape, generator=g)
    if positive:
        p, t = p + 0.05, t + 0.05
    if nan:
        t[0, :, 3:11, 5:17] = float('nan')
    return p, t


def test_mse_loss_and_log_loss_value_and_gradient():
    """model/loss.py:12-19 restated with torch on the CPU (boolean-mask gathers, as the reference) vs the device path."""
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200.model.loss import mse_loss, scale_invariant_log_loss
    p, t = _pair(1)
    pr = p.clone().requires_grad_(True)
    ref = F.mse_loss(pr[~torch.isnan(t)], t[~torch.isnan(t)])
    ref.backward()
    pg = p.to(dev()).requires_grad_(True)
    out = mse_loss(pg, t.to(dev()))
    (out * 3.0).backward()
    assert abs(out.item() - ref.item()) <= 1e-6
    np.testing.assert_allclose(pg.grad.cpu().numpy(), 3.0 * pr.grad.numpy(), rtol=1e-5, atol=1e-9)
    # log-space statistic (metric-depth inputs)
    p, t = _pair(2, positive=True)
    pr = p.clone().requires_grad_(True)
    d = torch.log(pr) - torch.log(t)
    ok = ~torch.isnan(d)
    ref = (d[ok] ** 2).mean() - 0.85 * d[ok].mean() ** 2
    ref.backward()
    pg = p.to(dev()).requires_grad_(True)
    out = scale_invariant_log_loss(pg, t.to(dev()), n_lambda=0.85)
    out.backward()
    assert abs(out.item() - ref.item()) <= 2e-6
    np.testing.assert_allclose(pg.grad.cpu().numpy(), pr.grad.numpy(), rtol=2e-4, atol=1e-8)
    assert R.scale_invariant_loss is not None


def test_multi_scale_gradient_preview_branch():
    """MultiScaleGradient(preview=True) (loss.py:46-47,59-60, called at lstm_trainer.py:162-165): per scale the Sobel
    magnitude of the pooled difference, bicubic-resized to (2H, 2W)."""
    from rpg_ramnet_b200.model.loss import multi_scale_grad_loss
    p, t = _pair(3, shape=(2, 1, 32, 64), nan=False)
    rec = multi_scale_grad_loss(p.to(dev()), t.to(dev()), preview=True)
    assert len(rec) == 4
    diff = p - t
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]]) / 8.0
    k = torch.stack([kx, kx.t()]).unsqueeze(1)
    for s, r in enumerate(rec):
        assert r.shape == (2, 1, 64, 128)
        q = F.avg_pool2d(diff, 2 ** s, 2 ** s)
        g = F.conv2d(F.pad(q, [1, 1, 1, 1], mode='replicate'), k)
        mag = torch.sqrt((g ** 2).sum(1, keepdim=True) + 1e-6)
        ref = F.interpolate(mag, size=(64, 128), mode='bicubic', align_corners=True)
        np.testing.assert_allclose(r.cpu().numpy(), ref.numpy(), rtol=1e-4, atol=1e-6)


def test_si_loss_batch_equals_individual_terms():
    """SILossBatch (one statistics buffer, one exchange per sequence) == T calls of scale_invariant_loss, values and
    gradients, including the upstream gradient read on the device."""
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200.model.loss import SILossBatch
    pairs = [_pair(10 + i, nan=(i % 2 == 0)) for i in range(5)]
    ws = [1.0, 0.5, 2.0, 1.0, 0.25]
    a = [p.to(dev()).requires_grad_(True) for p, _ in pairs]
    b = [p.to(dev()).requires_grad_(True) for p, _ in pairs]
    batch = SILossBatch(8, dev())
    for x, (_, t), w in zip(a, pairs, ws):
        batch.add(x, t.to(dev()), w, 0.85)
    terms = batch.finish()
    (2.0 * terms.sum() / 3.0).backward()
    singles = [R.scale_invariant_loss(x, t.to(dev()), w, 0.85) for x, (_, t), w in zip(b, pairs, ws)]
    (2.0 * sum(singles) / 3.0).backward()
    for i in range(5):
        assert abs(terms[i].item() - singles[i].item()) <= 1e-7
        assert torch.equal(a[i].grad, b[i].grad)


def test_trainer_loss_mix_with_grad_loss_matches_oracle():
    """The loss LSTMTrainer.forward_pass_sequence builds for the shipped config (lstm_trainer.py:152-226,275-288,381-382:
    per key and timestep  w_key * [SI + grad_loss.weight * MSG], summed / L, times K_keys through the shared-dict
    aliasing) on our model, losses and backward, vs the CPU oracle with torch autograd."""
    import rpg_ramnet_b200 as R
    from rpg_ramnet_b200.model.loss import multi_scale_grad_loss, scale_invariant_loss
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = R.ERGB2DepthRecurrent(dict(CFG, mma_kind='fp32'))
    model.train().to(dev())
    seq = O.synth_sequence(2, 64, 96, 2, 1, seed=4, with_targets=True)
    comp, wts, w_grad, L = ['image', 'events0'], [1.0, 1.0], 0.25, len(seq)

    def mix(preds_per_step, tgt, si, msg):
        loss_dict = {'losses': [], 'grad_losses': []}
        keys = []
        for preds, item in zip(preds_per_step, seq):
            for key, p in preds.items():
                if key in comp:
                    if key not in keys:
                        keys.append(key)
                    w = wts[comp.index(key)]
                    loss_dict['losses'].append(w * si(p, tgt(item['depth_' + key]), weight=1.0, n_lambda=1.0))
                    loss_dict['grad_losses'].append(w * msg(p, tgt(item['depth_' + key])))
        per_key = sum(loss_dict['losses']) / float(L) + w_grad * sum(loss_dict['grad_losses']) / float(L)
        return len(keys) * per_key

    prev_s, prev_l, outs = None, {'events0': None, 'image': None}, []
    for item in seq:
        preds, supers, lstm = model(item, prev_s, prev_l)
        outs.append(preds)
        prev_s, prev_l = supers['image'], lstm
    loss = mix(outs, lambda t: t.to(dev()), scale_invariant_loss, multi_scale_grad_loss)
    loss.backward()
    sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    o_s, o_l, o_outs = None, {'events0': None, 'image': None}, []
    for item in seq:
        p, s, l_ = O.ergb2depth_recurrent(sd, CFG, item, o_s, o_l)
        o_outs.append(p)
        o_s, o_l = s['image'], l_
    ref = mix(o_outs, lambda t: t, O.si_loss, O.multi_scale_grad_loss)
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 1e-4 * max(1.0, abs(ref.item())), (loss.item(), ref.item())
    for n, p in model.named_parameters():
        a, b = p.grad.detach().cpu().double(), sd[n].grad.double()
        assert float((a - b).norm()) <= 5e-3 * float(b.norm()) + 1e-9, n


def test_fused_adam_is_a_torch_optimizer_with_interchangeable_checkpoints():
    """base/base_trainer.py:36-43 builds the optimiser by name and wraps it in a torch lr scheduler; :133-179 saves
    optimizer.state_dict() and on resume loads it and walks optimizer.state.values().  FusedAdam must support all of
    that, and its checkpoints must load into torch.optim.Adam (and back) with the same continuation."""
    import rpg_ramnet_b200 as R
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 5, 1)).to(dev())
    twin = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 5, 1)).to(dev())
    twin.load_state_dict(net.state_dict())
    opt = R.FusedAdam(net.parameters(), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False)
    ref = torch.optim.Adam(twin.parameters(), lr=1e-2)
    assert isinstance(opt, torch.optim.Optimizer) and len(opt.param_groups) == 1
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.5)
    sched_ref = torch.optim.lr_scheduler.ExponentialLR(ref, gamma=0.5)
    g = torch.Generator().manual_seed(2)

    for step in range(3):
        opt.zero_grad()
        ref.zero_grad()
        for (p, q) in zip(net.parameters(), twin.parameters()):
            gr = (torch.randn(p.shape, generator=g) * 0.1).to(dev())
            p.grad.copy_(gr)
            q.grad = gr.clone()
        opt.step()
        ref.step()
        if step == 1:
            sched.step()
            sched_ref.step()
    assert opt.param_groups[0]['lr'] == ref.param_groups[0]['lr'] == 5e-3
    for p, q in zip(net.parameters(), twin.parameters()):
        assert float((p - q).abs().max()) <= 2e-6
    # checkpoint: ours -> torch.optim.Adam and -> a fresh FusedAdam (the resume path), then one more identical step
    import copy
    ckpt = copy.deepcopy(opt.state_dict())      # as torch.save / torch.load would: state_dict() hands out live tensors
    assert int(float(ckpt['state'][0]['step'])) == 3
    for state in opt.state.values():                       # base_trainer.py:171-175
        for k, v in state.items():
            if isinstance(v, torch.Tensor):
                state[k] = v.to(dev())
    twin2 = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 5, 1)).to(dev())
    twin2.load_state_dict(net.state_dict())
    ref2 = torch.optim.Adam(twin2.parameters(), lr=1e-2)
    ref2.load_state_dict(ckpt)
    net3 = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.Conv2d(8, 5, 1)).to(dev())
    net3.load_state_dict(net.state_dict())
    opt3 = R.FusedAdam(net3.parameters(), lr=1e-2, capturable=True)
    opt3.load_state_dict(ckpt)
    assert int(opt3.step_dev[0].item()) == 3               # bias correction must continue, not restart
    gr = [(torch.randn(p.shape, generator=g) * 0.1).to(dev()) for p in net.parameters()]
    for mod, o in ((net, opt), (twin2, ref2), (net3, opt3)):
        o.zero_grad()
        for p, x in zip(mod.parameters(), gr):
            if p.grad is None:
                p.grad = x.clone()
            else:
                p.grad.copy_(x)
        o.step()
    for p, q, r in zip(net.parameters(), twin2.parameters(), net3.parameters()):
        assert float((p - q).abs().max()) <= 2e-6 and float((p - r).abs().max()) <= 1e-7
    assert int(float(opt3.state_dict()['state'][0]['step'])) == 4     # capturable: the step is read back from the device


def test_graph_runner_rejects_stale_states():
    """cuda_graphs=True: states returned by a pass alias the runner's ping-pong buffers and are valid for one further
    pass; handing back an older set raises instead of silently reading overwritten memory."""
    import rpg_ramnet_b200 as R
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = R.ERGB2DepthRecurrent(dict(CFG, cuda_graphs=True))
    model.eval().to(dev())
    seq = O.synth_sequence(1, 32, 32, 3, 1, seed=1, with_targets=False)
    lstm = {'events0': None, 'image': None}
    with torch.no_grad():
        _, s0, _ = model(seq[0], None, lstm)
        old = s0['events0']                                 # written by the first pass, overwritten by the third
        _, s1, _ = model(seq[1], s0['image'], lstm)
        with pytest.raises(R.RamnetError):
            model(seq[2], old, lstm)
        model(seq[2], s1['image'], lstm)                    # the latest states are fine


def test_inference_output_stage_matches_test_py_restatement():
    """grey / colour-map PNG payloads and the metric-space scale of test.py:259-290,365-379 produced on
    the device (ramnet_depth_output) vs the numpy restatement of those lines (oracle/dataio_oracle.py)."""
    from oracle import dataio_oracle as D
    from rpg_ramnet_b200.utils.inference_output import colormap_lut, depth_outputs
    g = torch.Generator().manual_seed(21)
    pred = torch.rand(3, 1, 64, 96, generator=g)
    pred[1] = 0.37                                    # constant map: max - min == 0
    gt = torch.rand(3, 1, 64, 96, generator=g)
    gt[2, :, 5:20, 7:30] = float('nan')               # GT with NaN: make_colormap collapses to a constant image
    lut = colormap_lut()
    reg, clip = 3.70378, 80.0
    grey, bgr, scale = depth_outputs(pred.to(dev()), lut=lut, target=gt.to(dev()), reg_factor=reg, clip_distance=clip)
    grey_gt, bgr_gt, _ = depth_outputs(gt.to(dev()), lut=lut)
    for n in range(3):
        want_grey = D.grey_png_payload(pred[n].numpy())
        got = grey[n].cpu().numpy()
        assert np.abs(got.astype(int) - want_grey.astype(int)).max() <= 1 and (got != want_grey).mean() <= 1e-3
        for img, out in ((pred[n].numpy(), bgr[n]), (gt[n].numpy(), bgr_gt[n])):
            want = D.make_colormap_payload(img, lut)
            o = out.cpu().numpy()
            assert o.shape == want.shape
            # a LUT index may differ by one where x * 256 sits on an integer boundary (fp32 division order)
            assert (np.abs(o.astype(int) - want.astype(int)).max(axis=-1) > 3).mean() <= 2e-3
        if n < 2:
            want_scale = D.metric_scale(pred[n, 0].numpy(), gt[n, 0].numpy(), reg, clip)
            assert abs(float(scale[n]) - float(want_scale)) <= 1e-5 * abs(float(want_scale))
    assert torch.isnan(scale[2])                       # NaN in the target propagates, as in the reference
    assert float(bgr_gt[2].reshape(-1, 3).float().std(0).max()) == 0.0      # one colour for the NaN-carrying ground truth


@pytest.mark.parametrize('mode', ['host_pinned', 'device_static', 'device', 'no_overlap'])
def test_graph_runner_pass_overlap_is_bit_identical_to_eager(mode, monkeypatch):
    """Round 2: the front of the next pass (head, encoders, state update) runs on a second stream under the decoder of
    the current one.  Same kernels on the same data => bit-identical depth maps and states over a long sequence, for
    host inputs (copy stream), device inputs ordered after the caller's stream, and device inputs declared static."""
    import rpg_ramnet_b200 as R
    if mode == 'no_overlap':
        monkeypatch.setenv('RAMNET_PASS_OVERLAP', '0')
    cfg = dict(CFG, every_x_rgb_frame=2)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        eager = R.ERGB2DepthRecurrent(dict(cfg))
        graphed = R.ERGB2DepthRecurrent(dict(cfg, cuda_graphs=True, inputs_static=(mode == 'device_static')))
    graphed.load_state_dict(eager.state_dict())
    eager.eval().to(dev())
    graphed.eval().to(dev())
    seq = O.synth_sequence(2, 64, 96, 12, 2, seed=5, with_targets=False)
    if mode == 'host_pinned':
        seq_g = [{k: v.pin_memory() for k, v in it.items()} for it in seq]
    else:
        seq_g = [{k: v.to(dev()) for k, v in it.items()} for it in seq]
    torch.cuda.synchronize()
    sa = sb = None
    la = lb = {'events0': None, 'events1': None, 'image': None}
    with torch.no_grad():
        for t, (it_a, it_b) in enumerate(zip(seq, seq_g)):
            pa, sa_d, la = eager(it_a, sa, la)
            pb, sb_d, lb = graphed(it_b, sb, lb)
            sa, sb = sa_d['image'], sb_d['image']
            for k in pa:
                assert torch.equal(pa[k], pb[k]), (mode, t, k)
            for x, y in zip(sa, sb):
                assert torch.equal(x, y), (mode, t)
    runner = next(iter(graphed._runners.values()))
    assert runner.overlap == (mode != 'no_overlap')
    assert {k[0] for k in runner.graphs} == {'front', 'back'}
