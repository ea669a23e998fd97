"""GPU: whole-model parity — our CUDA graph vs (a) the reference's own outputs (tests/golden) and
(b) the CPU oracle on the same seeded inputs, through the reference-shaped module API."""
import numpy as np
import pytest
import torch

from oracle import ramnet_oracle as O
from helpers import (MODEL_CASES, build_product_model, case_inputs, flat_supers, load_case, max_rel_err,
                     run_oracle_sequence, run_product_sequence)

pytestmark = pytest.mark.gpu

SUPPORTED = list(MODEL_CASES)
# north_star tolerance: fp32 depth maps within 1e-3 relative of the reference forward.
REL_TOL = {'fp32': 1e-4, 'tf32': 1e-3}
# Train-mode norm configurations (not shipped): the pred layer's BatchNorm / InstanceNorm rescales the logits to unit
# variance, so the depth map is ~30x more sensitive to the 2^-11 operand rounding of kind::tf32 than the norm-free
# models (whose seed-0 logits sit near 0).  Against the reference (fp32 operands) TF32 mode is held to 2e-2 there, and to
# no worse than the CPU oracle evaluated on TF32-rounded operands deviates from it (test_live_norm_tf32_error_is_the_operand_
# rounding_floor: 0.5 - 1.6e-2, i.e. any TF32 convolution shows it); mma_kind='fp32' meets 1e-4 against the reference itself.
LIVE_NORM_TF32_TOL = 2e-2


def _tol(name, kind, meta):
    return LIVE_NORM_TF32_TOL if (kind == 'tf32' and meta.get('train', False)) else REL_TOL[kind]


@pytest.mark.parametrize('kind', ['fp32', 'tf32'])
@pytest.mark.parametrize('name', SUPPORTED)
def test_model_matches_reference_golden(name, kind):
    g, meta = load_case(name)
    model, _ = build_product_model(meta, mma_kind=kind)
    model.to('cuda:0')
    outs = run_product_sequence(model, meta, case_inputs(meta))
    n = 0
    for l, (preds, supers) in enumerate(outs):
        assert [k for k in preds] == [k.split('/')[-1] for k in g.files if k.startswith(f'pred/{l}/')]
        for key, p in preds.items():
            ref = g[f'pred/{l}/{key}']
            assert p.shape == ref.shape and p.dtype == torch.float32
            err = max_rel_err(p.cpu().numpy(), ref)
            assert err <= _tol(name, kind, meta), f'{name} {kind} pred/{l}/{key}: max rel err {err:.3e}'
            n += 1
        if supers.get('image') is not None:
            for key, s in supers.items():
                for j, t in enumerate(flat_supers(s)):
                    ref = g[f'super/{l}/{key}/{j}']
                    got = t[:, ::8, ::4, ::4].cpu().numpy()
                    scale = max(1e-3, float(np.abs(ref).max()))
                    assert np.abs(got - ref).max() <= (5e-5 if kind == 'fp32' else 3e-3) * scale, \
                        f'{name} {kind} super/{l}/{key}/{j}'
    assert n == len([k for k in g.files if k.startswith('pred/')])
    # train-mode norm cases: the running statistics the sequence leaves behind (momentum update, unbiased variance)
    sd = model.state_dict()
    for k in [k for k in g.files if k.startswith('buf/')]:
        ref = g[k]
        got = sd[k[4:]].cpu().numpy()
        if ref.dtype.kind in 'iu':
            assert np.array_equal(got, ref), k
        else:
            assert np.abs(got - ref).max() <= (1e-5 if kind == 'fp32' else 2e-3) * max(1.0, float(np.abs(ref).max())), \
                f'{name} {kind} {k}'


@pytest.mark.parametrize('name', ['bn_train', 'in_train', 'bn_train_tconv_lstm', 'unet_bn_train'])
def test_live_norm_tf32_error_is_the_operand_rounding_floor(name):
    """Why LIVE_NORM_TF32_TOL is not 1e-3: the CPU oracle itself, evaluated on TF32-rounded conv operands (fp32
    accumulate, everything else exact), deviates from the reference's fp32 output by 0.5 - 1.6e-2 on these
    configurations (1e-5 on the shipped block, 2e-4 with eval-mode BatchNorm) -- the sensitivity belongs to TF32 operand
    rounding under unit-variance logits, not to this implementation.  The CUDA path must not be worse than that floor
    (x1.5 for the different rounding points: the product also keeps recurrent states TF32-rounded)."""
    g, meta = load_case(name)
    model, _ = build_product_model(meta, mma_kind='tf32')
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model.to('cuda:0')
    seq = case_inputs(meta)
    outs = run_product_sequence(model, meta, seq)
    with torch.no_grad(), O.tf32_operands():
        ref = run_oracle_sequence(sd, meta, seq)
    ours = floor = 0.0
    for l, ((preds, _), (rpreds, _)) in enumerate(zip(outs, ref)):
        for key, p in preds.items():
            gold = g[f'pred/{l}/{key}']
            ours = max(ours, max_rel_err(p.cpu().numpy(), gold))
            floor = max(floor, max_rel_err(rpreds[key].numpy(), gold))
    print(f'{name}: vs reference: CUDA tf32 {ours:.3e}, oracle on tf32-rounded operands {floor:.3e}')
    assert floor >= 2e-3, 'the premise of LIVE_NORM_TF32_TOL no longer holds: tighten it'
    assert ours <= 1.5 * floor, f'{name}: {ours:.3e} vs floor {floor:.3e}'


@pytest.mark.parametrize('kind', ['fp32', 'tf32'])
def test_forward_events_images_decoder_api(kind):
    """Inner API of StateNetPhasedRecurrent called in an irregular order (BASELINE config 4's schedule):
    3 event passes, 1 image pass, 1 event pass, on a non-square MVSEC-like crop."""
    g, meta = load_case('cfg1_shipped')
    meta = dict(meta, H=64, W=88, B=1, wscale=1.5)
    model, cfg = build_product_model(meta, mma_kind=kind)
    model.to('cuda:0')
    net = model.statenetphasedrecurrent
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ocfg = O.NetCfg(cfg)
    gen = torch.Generator().manual_seed(77)
    ev = [torch.randn(1, 5, 64, 88, generator=gen) * (torch.rand(1, 5, 64, 88, generator=gen) < 0.1) for _ in range(4)]
    im = torch.rand(1, 1, 64, 88, generator=gen)
    s_ref = O.zero_super_states(ocfg, 1, 64, 88)
    s = model._zero_states(1, 64, 88)
    order = [('e', ev[0]), ('e', ev[1]), ('e', ev[2]), ('i', im), ('e', ev[3])]
    with torch.no_grad():
        for which, x in order:
            if which == 'e':
                s, _ = net.forward_events(x.to('cuda:0'), s, None, None)
                s_ref, _ = O.forward_events(sd, ocfg, x, s_ref, None)
            else:
                s, _ = net.forward_images(x.to('cuda:0'), s, None, None)
                s_ref, _ = O.forward_images(sd, ocfg, x, s_ref, None)
            d = net.forward_decoder(s)
            d_ref = O.forward_decoder(sd, ocfg, s_ref)
            assert max_rel_err(d.cpu().numpy(), d_ref.numpy()) <= REL_TOL[kind]


def test_states_round_trip_through_plain_nchw_tensors():
    """A caller may hand back states as ordinary contiguous NCHW tensors (e.g. after a checkpoint)."""
    g, meta = load_case('rect_b2')
    model, _ = build_product_model(meta, mma_kind='fp32')
    model.to('cuda:0')
    seq = case_inputs(meta)
    with torch.no_grad():
        p1, s1, l1 = model(seq[0], None, {'events0': None, 'image': None})
        p2, _, _ = model(seq[1], s1['image'], l1)
        s_plain = [t.contiguous().clone() for t in s1['image']]
        p2b, _, _ = model(seq[1], s_plain, l1)
    assert torch.equal(p2['image'], p2b['image'])


def test_weight_cache_tracks_parameter_updates():
    g, meta = load_case('cfg1_shipped')
    meta = dict(meta, H=32, W=32)
    model, _ = build_product_model(meta, mma_kind='fp32')
    model.to('cuda:0')
    item = O.synth_sequence(1, 32, 32, 1, 1, seed=3)[0]
    states = {'events0': None, 'image': None}
    with torch.no_grad():
        a = model(item, None, states)[0]['image'].clone()
        b = model(item, None, states)[0]['image'].clone()
        assert torch.equal(a, b)
        model.statenetphasedrecurrent.pred.conv2d.bias.add_(0.5)
        model.statenetphasedrecurrent.decoders[2].conv2d.weight.mul_(1.1)
        c = model(item, None, states)[0]['image']
    assert not torch.equal(a, c)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    ref = O.ergb2depth_recurrent(sd, dict(meta['config']), item, None, states)[0]['image']
    assert max_rel_err(c.cpu().numpy(), ref.numpy()) <= 1e-4


@pytest.mark.parametrize('name', ['rect_b2', 'lstm_state', 'k5_shipped'])
def test_cuda_graph_replay_is_bit_identical_to_eager(name):
    """config['cuda_graphs']=True replays each pass as one CUDA graph (ping-pong state buffers); same kernels,
    same order => bit-identical depth maps and states, across a state reset and a weight update."""
    g, meta = load_case(name)
    eager, _ = build_product_model(meta, mma_kind='tf32')
    eager.to('cuda:0')
    meta_g = dict(meta, config=dict(meta['config'], cuda_graphs=True))
    graphed, _ = build_product_model(meta_g, mma_kind='tf32')
    graphed.to('cuda:0')
    assert graphed.cuda_graphs and graphed.statenetphasedrecurrent.graph_capable()
    seq = case_inputs(meta)
    K = meta['config'].get('every_x_rgb_frame', 1)
    for rep in range(2):                      # second repetition: state reset (None) with graphs already captured
        sa = sb = None
        la = {f'events{k}': None for k in range(K)}
        la['image'] = None
        lb = dict(la)
        with torch.no_grad():
            for item in seq:                  # lock-step: graph-mode states alias ping-pong buffers (valid 1 step)
                pa, sa_d, la = eager(item, sa, la)
                pb, sb_d, lb = graphed(item, sb, lb)
                sa, sb = sa_d['image'], sb_d['image']
                assert list(pa) == list(pb)
                for k in pa:
                    assert torch.equal(pa[k], pb[k]), (name, rep, k)
                for x, y in zip(flat_supers(sa), flat_supers(sb)):
                    assert torch.equal(x, y), (name, rep)
            for m in (eager, graphed):        # weight update => graphs are re-captured with re-packed weights
                m.statenetphasedrecurrent.resblocks[0].conv1.weight.mul_(1.01)


def test_cuda_graph_replay_with_live_instance_norm():
    """norm='IN' in eval mode: the ConvLayers' InstanceNorm folds (running statistics) but the ResidualBlock's has none
    and computes instance statistics inside the captured back graph (ramnet_norm_fwd: scratch from the graph pool,
    float64 atomics -- so graph and eager agree to rounding, not bit for bit) -- and both match the reference golden."""
    g, meta = load_case('in_eval')
    eager, _ = build_product_model(meta, mma_kind='tf32')
    eager.to('cuda:0')
    graphed, _ = build_product_model(dict(meta, config=dict(meta['config'], cuda_graphs=True)), mma_kind='tf32')
    graphed.to('cuda:0')
    assert graphed._graphs_active()
    seq = case_inputs(meta)
    for rep in range(2):
        sa = sb = None
        la, lb = {'events0': None, 'image': None}, {'events0': None, 'image': None}
        with torch.no_grad():
            for l, item in enumerate(seq):
                pa, sa_d, la = eager(item, sa, la)
                pb, sb_d, lb = graphed(item, sb, lb)
                sa, sb = sa_d['image'], sb_d['image']
                for k in pa:
                    assert max_rel_err(pb[k].cpu().numpy(), pa[k].cpu().numpy()) <= 1e-5, (rep, l, k)
                    assert max_rel_err(pb[k].cpu().numpy(), g[f'pred/{l}/{k}']) <= 1e-3, (rep, l, k)


def test_unsupported_and_bad_shapes_raise():
    import rpg_ramnet_b200 as R
    g, meta = load_case('cfg1_shipped')
    model, _ = build_product_model(meta, mma_kind='fp32')
    model.to('cuda:0')
    bad = O.synth_sequence(1, 36, 52, 1, 1, seed=1)[0]      # 36 % 8 != 0: reference fails too (Appendix A)
    with torch.no_grad(), pytest.raises(R.RamnetError):
        model(bad, None, {'events0': None, 'image': None})


def test_smoke_entry():
    import __graft_entry__ as ge
    ge.smoke()
