"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python -m oracle.make_golden            (from the repo root; needs /root/reference)

TEST INFRASTRUCTURE ONLY.  The reference ships no golden vectors, so
parity is pinned on outputs of the reference itself: model forward (CPU fp32,
torch 2.11), scale_invariant_loss, events_to_voxel_grid and torch.optim.Adam.
Inputs are NOT stored: they are regenerated from seeds by
``oracle.ramnet_oracle.synth_sequence`` / ``synth_events`` (torch / numpy CPU
generators are deterministic), weights from ``torch.manual_seed(0)`` + the
reference's construction order (checksums stored).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from oracle import ramnet_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')

SHIPPED = dict(num_bins_rgb=1, num_bins_events=5, skip_type='sum', recurrent_block_type='conv',
               state_combination='convgru', num_encoders=3, base_num_channels=32,
               num_residual_blocks=2, use_upsample_conv=True, norm='none')


def _cfg(**kw):
    c = dict(SHIPPED)
    c.update(kw)
    return c


# name -> (arch, model config, B, H, W, L, input seed, weight scale)
CASES = {
    # BASELINE.json configs[0]: 128x128, B=1, seq=2, shipped block, K=1
    'cfg1_shipped':  ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1), 1, 128, 128, 2, 1, 1.0),
    'cfg1_stress2':  ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1), 1, 128, 128, 2, 1, 2.0),
    'rect_b2':       ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1), 2, 64, 96, 3, 4, 1.5),
    'k5_shipped':    ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=5), 1, 64, 64, 2, 5, 1.5),
    'lstm_state':    ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1, state_combination='convlstm'),
                      2, 64, 64, 2, 6, 1.5),
    'lstm_enc':      ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=2, state_combination='convlstm',
                                                  recurrent_block_type='convlstm'), 1, 64, 64, 2, 7, 1.5),
    'bn_eval':       ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1, norm='BN'), 2, 64, 64, 2, 8, 1.5),
    'baseline_rgb':  ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=2, state_combination='convlstm',
                                                  baseline='rgb', loss_composition='image'),
                      1, 64, 64, 2, 9, 1.5),
    'baseline_e':    ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=3, num_bins_rgb=5,
                                                  state_combination='convlstm',
                                                  baseline='e', loss_composition='image'),
                      1, 64, 64, 2, 10, 1.5),
    'baseline_ergb0': ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=3, num_bins_rgb=6, num_bins_events=6,
                                                   state_combination='convlstm',
                                                   baseline='ergb0', loss_composition='image'),
                       1, 64, 64, 2, 11, 1.5),
    'unet':          ('ERGB2Depth', _cfg(num_bins_rgb=6), 2, 64, 64, 1, 12, 1.5),
    'transposed':    ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1, use_upsample_conv=False),
                      1, 64, 64, 2, 13, 1.5),
    # live norm layers (round 2): train-mode batch / instance statistics + running-statistics update, and the
    # ResidualBlock's InstanceNorm2d (no running statistics: instance statistics in eval mode too)
    'bn_train':      ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1, norm='BN'), 2, 64, 64, 2, 14, 1.5),
    'in_train':      ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1, norm='IN'), 2, 64, 64, 2, 15, 1.5),
    'in_eval':       ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1, norm='IN'), 2, 64, 64, 2, 16, 1.5),
    'bn_train_tconv_lstm': ('ERGB2DepthRecurrent', _cfg(every_x_rgb_frame=1, norm='BN', use_upsample_conv=False,
                                                        state_combination='convlstm', recurrent_block_type='convlstm'),
                            2, 64, 64, 2, 17, 1.5),
    'unet_bn_train': ('ERGB2Depth', _cfg(num_bins_rgb=6, norm='BN'), 2, 64, 64, 2, 18, 1.5),
    'unet_transposed': ('ERGB2Depth', _cfg(num_bins_rgb=6, use_upsample_conv=False), 2, 64, 96, 1, 19, 1.5),
    'unet_concat':   ('ERGB2Depth', _cfg(num_bins_rgb=6, skip_type='concat'), 2, 64, 96, 1, 25, 1.5),
    'unet_concat_bn_train': ('ERGB2Depth', _cfg(num_bins_rgb=6, skip_type='concat', norm='BN'), 2, 64, 64, 2, 26, 1.5),
    'unet_transposed_in': ('ERGB2Depth', _cfg(num_bins_rgb=6, use_upsample_conv=False, norm='IN'), 1, 64, 64, 2, 20, 1.5),
}
TRAIN_MODE = {'bn_train', 'in_train', 'bn_train_tconv_lstm', 'unet_bn_train', 'unet_transposed_in', 'unet_concat_bn_train'}


scale_weights = O.scale_weights


def param_checksums(sd):
    names = sorted(sd.keys())
    return names, np.array([[float(sd[n].double().sum()), float(sd[n].double().abs().sum())] for n in names])


def seq_inputs(cfg, B, H, W, L, seed):
    K = cfg.get('every_x_rgb_frame', 1)
    bl = cfg.get('baseline', False)
    be = cfg['num_bins_rgb'] if bl in ('e', 'ergb0') else cfg['num_bins_events']
    return O.synth_sequence(B, H, W, L, K, seed, bins_events=be, bins_rgb=cfg['num_bins_rgb'])


def flat_supers(s):
    out = []
    for lvl in s:
        out.extend(lvl if isinstance(lvl, (list, tuple)) else [lvl])
    return out


def run_model_case(ns, name, spec):
    arch, cfg, B, H, W, L, seed, wscale = spec
    model = ref_import.build_model(ns, arch, cfg)
    scale_weights(model, wscale)
    train = name in TRAIN_MODE
    model.train(train)
    seq = seq_inputs(cfg, B, H, W, L, seed)
    K = cfg.get('every_x_rgb_frame', 1)
    out = {}
    names, cks = param_checksums(model.state_dict())
    out['param_names'] = np.array(names)
    out['param_checksums'] = cks
    prev_super = {'image': None}
    prev_lstm = {f'events{k}': None for k in range(K)}
    prev_lstm['image'] = None
    with torch.no_grad():
        for l, item in enumerate(seq):
            preds, supers, lstm = model(item, prev_super['image'], prev_lstm)
            for key, p in preds.items():
                out[f'pred/{l}/{key}'] = p.numpy().astype(np.float32)
            if supers.get('image') is not None:
                for key, s in supers.items():
                    for j, t in enumerate(flat_supers(s)):
                        out[f'super/{l}/{key}/{j}'] = t[:, ::8, ::4, ::4].numpy().astype(np.float32)
            prev_super, prev_lstm = supers, lstm
    if train:       # running statistics after the sequence (updated in place by every train-mode forward)
        for n, t in model.state_dict().items():
            if n.endswith(('running_mean', 'running_var', 'num_batches_tracked')):
                out['buf/' + n] = t.numpy().copy()
    out['meta'] = np.array(json.dumps(dict(arch=arch, config=cfg, B=B, H=H, W=W, L=L, seed=seed, wscale=wscale,
                                           train=train)))
    np.savez_compressed(os.path.join(OUT, f'model_{name}.npz'), **out)
    print(name, 'ok', {k: v.shape for k, v in out.items() if k.startswith('pred/0')})


def run_grad_case(ns, fname='grads_shipped.npz', train=False, seed=21, **cfg_kw):
    """fwd+bwd of the trainer's loss mix (SI only) -> loss value + per-tensor grad digests."""
    cfg = _cfg(every_x_rgb_frame=1, **cfg_kw)
    B, H, W, L = 2, 64, 64, 2
    model = ref_import.build_model(ns, 'ERGB2DepthRecurrent', cfg)
    scale_weights(model, 1.5)
    model.train(train)
    seq = seq_inputs(cfg, B, H, W, L, seed)
    comp, wts = ['image', 'events0'], [1.0, 1.0]
    prev_super, prev_lstm = {'image': None}, {'events0': None, 'image': None}
    terms, keys = [], []
    for item in seq:
        preds, supers, lstm = model(item, prev_super['image'], prev_lstm)
        for key, p in preds.items():
            if key in comp:
                if key not in keys:
                    keys.append(key)
                terms.append(wts[comp.index(key)] * ns.scale_invariant_loss(p, item['depth_' + key], 1.0, 1.0))
        prev_super, prev_lstm = supers, lstm
    loss = len(keys) * sum(terms) / float(L)     # lstm_trainer.py:253,279-281,381-382 aliasing
    loss.backward()
    out = {'loss': np.array(loss.item(), np.float64)}
    names = [n for n, _ in model.named_parameters()]
    out['names'] = np.array(names)
    out['grad_l2'] = np.array([float(p.grad.double().norm()) for _, p in model.named_parameters()])
    out['grad_sum'] = np.array([float(p.grad.double().sum()) for _, p in model.named_parameters()])
    for n, p in model.named_parameters():
        out['head/' + n] = p.grad.flatten()[:16].numpy().astype(np.float32)
    out['meta'] = np.array(json.dumps(dict(config=cfg, B=B, H=H, W=W, L=L, seed=seed, wscale=1.5, train=train,
                                           loss_composition=comp, loss_weights=wts)))
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print(fname, 'ok loss', loss.item())


def run_loss_cases(ns):
    g = torch.Generator().manual_seed(31)
    out = {}
    for i, (lam, w, nan_frac) in enumerate([(1.0, 1.0, 0.0), (0.5, 1.0, 0.1), (1.0, 0.7, 0.5), (0.85, 2.0, 0.02)]):
        p = torch.rand(2, 1, 24, 40, generator=g, requires_grad=True)
        t = torch.rand(2, 1, 24, 40, generator=g)
        t[torch.rand(t.shape, generator=g) < nan_frac] = float('nan')
        loss = ns.scale_invariant_loss(p, t, w, lam)
        loss.backward()
        out[f'{i}/pred'] = p.detach().numpy()
        out[f'{i}/target'] = t.numpy()
        out[f'{i}/loss'] = np.array(loss.item(), np.float64)
        out[f'{i}/grad'] = p.grad.numpy()
        out[f'{i}/params'] = np.array([lam, w])
    np.savez_compressed(os.path.join(OUT, 'si_loss.npz'), **out)
    print('si loss ok')


def run_adam_case():
    """torch.optim.Adam trajectory (base_trainer.py:36-37 builds it via getattr(optim, 'Adam'))."""
    g = torch.Generator().manual_seed(41)
    p = torch.nn.Parameter(torch.randn(4099, generator=g))
    opt = torch.optim.Adam([p], lr=3e-4, weight_decay=0)
    out = {'p0': p.detach().numpy().copy()}
    for s in range(3):
        gr = torch.randn(4099, generator=g) * (10.0 ** (s - 2))
        p.grad = gr.clone()
        opt.step()
        out[f'g{s}'] = gr.numpy()
        out[f'p{s + 1}'] = p.detach().numpy().copy()
    st = opt.state[p]
    out['m'] = st['exp_avg'].numpy()
    out['v'] = st['exp_avg_sq'].numpy()
    np.savez_compressed(os.path.join(OUT, 'adam.npz'), **out)
    print('adam ok')


def run_voxel_cases(ns):
    out = {}
    rng = np.random.default_rng(7)
    cases = {}
    W, H, B = 32, 24, 5
    cases['n1'] = (np.array([[0.5, 3, 4, 1]], np.float64), B, W, H)
    cases['n2_same_t'] = (np.array([[0.5, 3, 4, 1], [0.5, 31, 23, 0]], np.float64), B, W, H)      # dT == 0
    cases['n10'] = (O.synth_events(10, W, H, 1), B, W, H)
    cases['n2000'] = (O.synth_events(2000, W, H, 2), B, W, H)
    cases['n2000_hot'] = (O.synth_events(2000, W, H, 3, hot=True), B, W, H)
    ev = O.synth_events(500, W, H, 4)
    ev[:, 3] = np.where(ev[:, 3] == 0, -1.0, 1.0)                                           # +-1 polarity input
    cases['pm1'] = (ev, B, W, H)
    ev = O.synth_events(300, W, H, 5)
    ev[-40:, 0] = ev[-1, 0]                                                                 # many events on last stamp
    ev[:, 1] = np.where(rng.uniform(size=300) < 0.3, W - 1, ev[:, 1])                       # border pixels
    ev[:, 2] = np.where(rng.uniform(size=300) < 0.3, 0, ev[:, 2])
    cases['last_stamp_borders'] = (ev, B, W, H)
    cases['bins1'] = (O.synth_events(200, W, H, 6), 1, W, H)
    cases['bins2'] = (O.synth_events(200, W, H, 7), 2, W, H)
    cases['bins9_rect'] = (O.synth_events(5000, 96, 64, 8), 9, 96, 64)
    cases['n10000'] = (O.synth_events(10000, 96, 64, 9), 5, 96, 64)
    ev = O.synth_events(400, W, H, 10)
    ev[:, 1] += rng.uniform(0, 0.99, 400)                                                   # fractional x,y (trunc)
    ev[:, 2] += rng.uniform(0, 0.99, 400)
    cases['fractional_xy'] = (ev, B, W, H)
    for name, (ev, b, w, h) in cases.items():
        ref = ns.events_to_voxel_grid(ev.copy(), b, w, h)      # reference mutates its input -> copy
        out[name + '/events'] = ev
        out[name + '/grid'] = ref.astype(np.float32)
        out[name + '/shape'] = np.array([b, w, h])
    np.savez_compressed(os.path.join(OUT, 'voxel.npz'), **out)
    print('voxel ok', list(cases))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ns = ref_import.load()
    only = sys.argv[1:]
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        run_model_case(ns, name, spec)
    if not only or 'grads' in only:
        run_grad_case(ns)
    if not only or 'grads_norm' in only:
        run_grad_case(ns, 'grads_bn_train.npz', train=True, seed=22, norm='BN')
        run_grad_case(ns, 'grads_in_train.npz', train=True, seed=23, norm='IN')
        run_grad_case(ns, 'grads_bn_eval.npz', train=False, seed=24, norm='BN')
    if not only or 'misc' in only:
        run_loss_cases(ns)
        run_adam_case()
        run_voxel_cases(ns)


if __name__ == '__main__':
    main()
