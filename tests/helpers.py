"""Shared test helpers: golden loading, model construction, oracle driving."""
import contextlib
import io
import json
import os

import numpy as np
import torch

from oracle import ramnet_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

MODEL_CASES = ['cfg1_shipped', 'cfg1_stress2', 'rect_b2', 'k5_shipped', 'lstm_state', 'lstm_enc', 'bn_eval',
               'baseline_rgb', 'baseline_e', 'baseline_ergb0', 'unet', 'transposed',
               # live norm layers: train-mode BatchNorm / InstanceNorm statistics, ResidualBlock InstanceNorm in eval mode
               'bn_train', 'in_train', 'in_eval', 'bn_train_tconv_lstm', 'unet_bn_train',
               'unet_transposed', 'unet_transposed_in', 'unet_concat', 'unet_concat_bn_train']
GRAD_CASES = ['grads_shipped', 'grads_bn_train', 'grads_in_train', 'grads_bn_eval']


def load_case(name):
    g = np.load(os.path.join(GOLDEN, f'model_{name}.npz'))
    meta = json.loads(str(g['meta']))
    return g, meta


def build_product_model(meta, device_index=0, mma_kind=None):
    """Construct OUR module exactly like train.py:203-204 constructs the reference's."""
    import rpg_ramnet_b200 as R
    cfg = dict(meta['config'])
    cfg['gpu'] = device_index
    if mma_kind is not None:
        cfg['mma_kind'] = mma_kind
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = getattr(R, meta['arch'])(cfg)
    m.eval()
    O.scale_weights(m, meta['wscale'])
    m.train(bool(meta.get('train', False)))
    return m, cfg


def case_inputs(meta):
    cfg = meta['config']
    K = cfg.get('every_x_rgb_frame', 1)
    bl = cfg.get('baseline', False)
    be = cfg['num_bins_rgb'] if bl in ('e', 'ergb0') else cfg['num_bins_events']
    return O.synth_sequence(meta['B'], meta['H'], meta['W'], meta['L'], K, meta['seed'], bins_events=be,
                            bins_rgb=cfg['num_bins_rgb'])


def run_oracle_sequence(sd, meta, seq):
    """Drive the oracle over the sequence with state carry, like lstm_trainer.py:245-272.  Cases generated with the
    reference in train mode run under O.training_mode(): `sd`'s running statistics are updated in place."""
    if meta.get('train', False):
        with O.training_mode():
            return run_oracle_sequence(sd, dict(meta, train=False), seq)
    cfg = meta['config']
    K = cfg.get('every_x_rgb_frame', 1)
    outs = []
    if meta['arch'] == 'ERGB2Depth':
        for item in seq:
            outs.append((O.ergb2depth_unet(sd, cfg, item), None))
        return outs
    prev_super = None
    prev_lstm = {f'events{k}': None for k in range(K)}
    prev_lstm['image'] = None
    for item in seq:
        preds, supers, lstm = O.ergb2depth_recurrent(sd, cfg, item, prev_super, prev_lstm)
        outs.append((preds, supers))
        prev_super, prev_lstm = supers['image'], lstm
    return outs


def run_product_sequence(model, meta, seq):
    cfg = meta['config']
    K = cfg.get('every_x_rgb_frame', 1)
    outs = []
    prev_super = {'image': None}
    prev_lstm = {f'events{k}': None for k in range(K)}
    prev_lstm['image'] = None
    with torch.no_grad():
        for item in seq:
            preds, supers, lstm = model(item, prev_super['image'], prev_lstm)
            outs.append((preds, supers))
            prev_super, prev_lstm = supers, lstm
    return outs


def flat_supers(s):
    out = []
    for lvl in s:
        out.extend(lvl if isinstance(lvl, (list, tuple)) else [lvl])
    return out


def max_rel_err(a, b, floor=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
