"""CPU restatement of the RAM-Net forward / loss / optimiser hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Pinned against the
reference executed in the build container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``); the reference itself ships no golden vectors.

Every function is a plain functional restatement driven by a ``state_dict``
(name -> fp32 tensor, reference key names) and cites the reference lines it
follows.  Dense arithmetic (cross-correlation, bilinear x2) is delegated to
torch's CPU ``conv2d`` / ``interpolate`` exactly as the reference delegates it
to torch (third-party, pinned torch==1.6.0 in requirements.txt:90; the build
container has 2.11).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------
# a-1  events_to_voxel_grid            RAM_Net/utils/event_tensor_utils.py:71-117
# --------------------------------------------------------------------------
def voxel_grid(events: np.ndarray, num_bins: int, width: int, height: int) -> np.ndarray:
    """Bilinear-in-time event voting, numpy restatement.

    Follows event_tensor_utils.py:71-117 step for step but does NOT mutate its
    input (the reference overwrites columns 0 and 3, :95,:100).  Index
    arithmetic is int64 (truncation toward zero, :102), timestamps/weights are
    float64; ``np.add.at(float32_grid, idx, float64_vals)`` resolves to the
    float64 add loop, i.e. each step is grid = f32(f64(grid) + val), sequential
    in event order, left votes first, then right votes (:107-113) — probed:
    pre-casting val to float32 differs by 1 ulp on pixels with >= 2 votes.
    """
    assert events.ndim == 2 and events.shape[1] == 4
    assert num_bins > 0 and width > 0 and height > 0
    grid = np.zeros(num_bins * height * width, np.float32)
    if events.shape[0] == 0:
        return grid.reshape(num_bins, height, width)
    ev = np.asarray(events, dtype=np.float64)
    t0 = ev[0, 0]
    dT = ev[-1, 0] - t0
    if dT == 0:
        dT = 1.0
    ts = (num_bins - 1) * (ev[:, 0] - t0) / dT
    xs = ev[:, 1].astype(np.int64)
    ys = ev[:, 2].astype(np.int64)
    pol = ev[:, 3].copy()
    pol[pol == 0] = -1.0
    tis = ts.astype(np.int64)
    dts = ts - tis
    left = pol * (1.0 - dts)
    right = pol * dts
    ok = tis < num_bins
    np.add.at(grid, xs[ok] + ys[ok] * width + tis[ok] * width * height, left[ok])
    ok = (tis + 1) < num_bins
    np.add.at(grid, xs[ok] + ys[ok] * width + (tis[ok] + 1) * width * height, right[ok])
    return grid.reshape(num_bins, height, width)


def voxel_grid_votes(events: np.ndarray, num_bins: int, width: int, height: int):
    """The integer index stream and float64 vote values of `voxel_grid`,
    returned un-accumulated: (idx_left, val_left, idx_right, val_right) with
    index -1 where the vote is dropped.  Used for the bit-exact index check
    (cast the values to float32 to compare with the CUDA vote stream)."""
    ev = np.asarray(events, dtype=np.float64)
    n = ev.shape[0]
    if n == 0:
        z = np.zeros(0, np.int64)
        return z, np.zeros(0, np.float64), z, np.zeros(0, np.float64)
    t0 = ev[0, 0]
    dT = ev[-1, 0] - t0
    if dT == 0:
        dT = 1.0
    ts = (num_bins - 1) * (ev[:, 0] - t0) / dT
    xs = ev[:, 1].astype(np.int64)
    ys = ev[:, 2].astype(np.int64)
    pol = ev[:, 3].copy()
    pol[pol == 0] = -1.0
    tis = ts.astype(np.int64)
    dts = ts - tis
    base = xs + ys * width
    il = np.where(tis < num_bins, base + tis * width * height, -1)
    ir = np.where(tis + 1 < num_bins, base + (tis + 1) * width * height, -1)
    return il, pol * (1.0 - dts), ir, pol * dts



# --------------------------------------------------------------------------
# every dense contraction of the path goes through _conv2d so that the tests can also ask "what would exact arithmetic
# on TF32-ROUNDED OPERANDS give?" (the error model of tcgen05.mma kind::tf32 with fp32 accumulate).  The
# default is plain fp32 F.conv2d — the reference's arithmetic.  Inside `tf32_operands()` the input and the weight of
# every conv are rounded to TF32 (round-to-nearest-away on 13 mantissa bits, the product's cvt.rna.tf32) in the
# forward pass, and the gradient arriving at the conv output is rounded the same way in the backward pass (the
# product rounds dZ before its dgrad / wgrad GEMMs); rounding is straight-through for autograd.
# --------------------------------------------------------------------------
_TF32_OPERANDS = False


def rna_tf32(t: Tensor) -> Tensor:
    return ((t.detach().contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)


class _RoundFwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return rna_tf32(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y):
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        return rna_tf32(g)


class tf32_operands:
    """Context manager: conv operands (and conv-output gradients) rounded to TF32."""

    def __enter__(self):
        global _TF32_OPERANDS
        self.prev, _TF32_OPERANDS = _TF32_OPERANDS, True

    def __exit__(self, *exc):
        global _TF32_OPERANDS
        _TF32_OPERANDS = self.prev


def _conv2d(x, w, b=None, exact=False, **kw):
    """`exact`: a layer the product computes with fp32 FFMA even in TF32 mode (the 1x1 prediction conv)."""
    if _TF32_OPERANDS and x.dtype == torch.float32 and not exact:
        return _RoundBwd.apply(F.conv2d(_RoundFwd.apply(x), _RoundFwd.apply(w), b, **kw))
    return F.conv2d(x, w, b, **kw)

# --------------------------------------------------------------------------
# a-2/a-3  ConvLayer                    RAM_Net/model/submodules.py:8-35
# --------------------------------------------------------------------------
_TRAINING = False


class training_mode:
    """Context manager: norm layers behave as in model.train() -- BatchNorm2d / InstanceNorm2d(track_running_stats)
    normalise with batch / instance statistics and update the running statistics held in the state dict IN PLACE
    (momentum 0.1, unbiased variance), as nn.BatchNorm2d.forward / nn.InstanceNorm2d.forward do."""

    def __enter__(self):
        global _TRAINING
        self.prev, _TRAINING = _TRAINING, True

    def __exit__(self, *exc):
        global _TRAINING
        _TRAINING = self.prev


def _norm_eval(sd: StateDict, prefix: str, y: Tensor, kind: str) -> Tensor:
    """BatchNorm2d / InstanceNorm2d(track_running_stats=True) that follows a conv (submodules.py:21-24,29-30).
    Eval: running statistics.  Under `training_mode`: batch / instance statistics + running update."""
    rm, rv = sd[prefix + '.running_mean'].detach(), sd[prefix + '.running_var'].detach()
    if kind == 'BN':
        if _TRAINING and (prefix + '.num_batches_tracked') in sd:
            sd[prefix + '.num_batches_tracked'].add_(1)
        return F.batch_norm(y, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'], _TRAINING, 0.1, 1e-5)
    if kind == 'IN':  # affine=False default; running stats used in eval, instance stats (+ update) in train mode
        return F.instance_norm(y, rm, rv, None, None, _TRAINING, 0.1, 1e-5)
    return y


def conv_layer(sd: StateDict, prefix: str, x: Tensor, stride: int, padding: int,
               relu: bool = True, norm: Optional[str] = None, exact: bool = False) -> Tensor:
    """ConvLayer.forward (submodules.py:26-35): conv (+bias unless BN, :13)
    -> optional norm -> optional relu."""
    y = _conv2d(x, sd[prefix + '.conv2d.weight'], sd.get(prefix + '.conv2d.bias'), exact=exact,
                 stride=stride, padding=padding)
    if norm in ('BN', 'IN'):
        y = _norm_eval(sd, prefix + '.norm_layer', y, norm)
    return torch.relu(y) if relu else y


# --------------------------------------------------------------------------
# a-7  UpsampleConvLayer                submodules.py:69-97
# --------------------------------------------------------------------------
def upsample_conv_layer(sd: StateDict, prefix: str, x: Tensor, norm: Optional[str] = None) -> Tensor:
    """bilinear x2 (align_corners=False, :88) -> 5x5 s1 p2 conv -> norm -> relu."""
    up = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    y = _conv2d(up, sd[prefix + '.conv2d.weight'], sd.get(prefix + '.conv2d.bias'), stride=1, padding=2)
    if norm in ('BN', 'IN'):
        y = _norm_eval(sd, prefix + '.norm_layer', y, norm)
    return torch.relu(y)


def transposed_conv_layer(sd: StateDict, prefix: str, x: Tensor, norm: Optional[str] = None) -> Tensor:
    """TransposedConvLayer.forward (submodules.py:38-66): stride-2 5x5
    ConvTranspose2d, padding 2, output_padding 1 -> norm -> relu."""
    w = sd[prefix + '.transposed_conv2d.weight']
    if _TF32_OPERANDS and x.dtype == torch.float32:
        x, w = _RoundFwd.apply(x), _RoundFwd.apply(w)
    y = F.conv_transpose2d(x, w, sd.get(prefix + '.transposed_conv2d.bias'), stride=2, padding=2, output_padding=1)
    if _TF32_OPERANDS and y.dtype == torch.float32:
        y = _RoundBwd.apply(y)
    if norm in ('BN', 'IN'):
        y = _norm_eval(sd, prefix + '.norm_layer', y, norm)
    return torch.relu(y)


# --------------------------------------------------------------------------
# a-6  ResidualBlock                    submodules.py:182-215
# --------------------------------------------------------------------------
def residual_block(sd: StateDict, prefix: str, x: Tensor, norm: Optional[str] = None) -> Tensor:
    y = _conv2d(x, sd[prefix + '.conv1.weight'], sd.get(prefix + '.conv1.bias'), padding=1)
    if norm in ('BN', 'IN'):
        y = _norm_eval(sd, prefix + '.bn1', y, norm) if norm == 'BN' else F.instance_norm(y)
    y = torch.relu(y)
    y = _conv2d(y, sd[prefix + '.conv2.weight'], sd.get(prefix + '.conv2.bias'), padding=1)
    if norm in ('BN', 'IN'):
        y = _norm_eval(sd, prefix + '.bn2', y, norm) if norm == 'BN' else F.instance_norm(y)
    return torch.relu(y + x)


# --------------------------------------------------------------------------
# a-4  ConvGRU                          submodules.py:414-454
# --------------------------------------------------------------------------
def conv_gru(sd: StateDict, prefix: str, x: Tensor, h: Optional[Tensor]) -> Tensor:
    """u = s(Wu*[x,h]); r = s(Wr*[x,h]); o = tanh(Wo*[x, h.r]);
    h' = h(1-u) + o u   (submodules.py:446-452)."""
    if h is None:
        h = torch.zeros_like(x)
    xh = torch.cat([x, h], 1)
    u = torch.sigmoid(_conv2d(xh, sd[prefix + '.update_gate.weight'], sd[prefix + '.update_gate.bias'], padding=1))
    r = torch.sigmoid(_conv2d(xh, sd[prefix + '.reset_gate.weight'], sd[prefix + '.reset_gate.bias'], padding=1))
    o = torch.tanh(_conv2d(torch.cat([x, h * r], 1), sd[prefix + '.out_gate.weight'],
                            sd[prefix + '.out_gate.bias'], padding=1))
    return h * (1 - u) + o * u


# --------------------------------------------------------------------------
# a-5  ConvLSTM                         submodules.py:303-358
# --------------------------------------------------------------------------
def conv_lstm(sd: StateDict, prefix: str, x: Tensor,
              state: Optional[Sequence[Tensor]]) -> Tuple[Tensor, Tensor]:
    """One 3x3 conv to 4C; chunk order in, remember, out, cell (:344);
    c' = s(f) c + s(i) tanh(g); h' = s(o) tanh(c') (:355-356)."""
    if state is None:
        hsz = sd[prefix + '.Gates.weight'].shape[0] // 4
        h = x.new_zeros(x.shape[0], hsz, x.shape[2], x.shape[3])
        c = h.clone()
    else:
        h, c = state
    g = _conv2d(torch.cat([x, h], 1), sd[prefix + '.Gates.weight'], sd[prefix + '.Gates.bias'], padding=1)
    gi, gf, go, gc = g.chunk(4, 1)
    c2 = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gc)
    h2 = torch.sigmoid(go) * torch.tanh(c2)
    return h2, c2


# --------------------------------------------------------------------------
# a-9  StateNetPhasedRecurrent          RAM_Net/model/statenet.py:120-315
# --------------------------------------------------------------------------
class NetCfg:
    """The model-config keys of BaseERGB2Depth (model/model.py:16-77)."""

    def __init__(self, config: dict):
        self.num_bins_rgb = int(config['num_bins_rgb'])
        self.num_bins_events = int(config['num_bins_events'])
        self.skip_type = str(config.get('skip_type', 'sum'))
        self.state_combination = str(config.get('state_combination', 'sum'))
        self.num_encoders = int(config.get('num_encoders', 4))
        self.base = int(config.get('base_num_channels', 32))
        self.num_residual_blocks = int(config.get('num_residual_blocks', 2))
        self.recurrent_block_type = str(config.get('recurrent_block_type', 'convlstm'))
        self.norm = str(config['norm']) if 'norm' in config else None
        self.use_upsample_conv = bool(config.get('use_upsample_conv', True))
        self.K = config.get('every_x_rgb_frame', 1)
        self.baseline = config.get('baseline', False)
        self.loss_composition = config.get('loss_composition', False)


def _encode(sd: StateDict, cfg: NetCfg, modality: str, x: Tensor, prev_super, prev_lstm):
    """forward_events (statenet.py:204-239) / forward_images (:241-288)."""
    P = 'statenetphasedrecurrent.'
    head = 'head_events' if modality == 'events' else 'head_rgb'
    enc = 'encoders_events' if modality == 'events' else 'encoders_rgb'
    comb = 'state_combination_events' if modality == 'events' else 'state_combination_images'
    x = conv_layer(sd, P + head, x, 1, 2)                       # head: norm=None always (:139-145)
    if prev_lstm is None:
        prev_lstm = {'encoders': [None] * cfg.num_encoders, 'state_comb': [None] * cfg.num_encoders}
    supers, out_lstm = [], {'encoders': [], 'state_comb': []}
    for i in range(cfg.num_encoders):
        ep = f'{P}{enc}.{i}'
        if cfg.recurrent_block_type == 'conv':
            x = conv_layer(sd, ep, x, 2, 2, norm=cfg.norm)
            enc_state = None
        else:  # Recurrent2ConvLayer (submodules.py:122-142)
            x = conv_layer(sd, ep + '.conv', x, 2, 2, norm=cfg.norm)
            enc_state = conv_lstm(sd, ep + '.recurrent_block', x, prev_lstm['encoders'][i])
            x = enc_state[0]
        cp = f'{P}{comb}.{i}'
        sc = cfg.state_combination
        is_baseline = bool(cfg.baseline) and modality == 'images'
        if sc == 'convlstm' and not is_baseline:
            # statenet.py:222-229 / :263-270: LSTM state is the previous super state [h, c]
            st = conv_lstm(sd, cp + '.recurrent_block', x, prev_super[i])
            super_state, comb_state = st, st
            new_x = x
        else:
            if sc in ('sum', 'conv'):
                # state_sum / state_conv (statenet.py:23-28) return ONE tensor which the caller
                # tuple-unpacks (:231, :272): ill-formed in the reference unless B == 2.
                raise NotImplementedError("state_combination=%r is ill-formed in the reference" % sc)
            elif sc == 'convgru':
                val = conv_gru(sd, cp + '.recurrent_block', x, prev_super[i])
                out, comb_state = val, val                                # RecurrentConvLayer.forward :116-120
            elif sc == 'convlstm':  # baseline only (:280-283): state from prev_states_lstm
                st = conv_lstm(sd, cp + '.recurrent_block', x, prev_lstm['state_comb'][i])
                out, comb_state = st[0], st
            else:
                raise KeyError(sc)
            if is_baseline:
                new_x, super_state = out, out      # baselines feed the recurrent output upward (:276-284)
            else:
                new_x, super_state = x, out        # RAM-Net feeds the encoder output upward (:260-275)
        x = new_x
        supers.append(super_state)
        out_lstm['encoders'].append(enc_state)
        out_lstm['state_comb'].append(comb_state)
    return supers, out_lstm


def forward_events(sd, cfg, x, prev_super, prev_lstm):
    return _encode(sd, cfg, 'events', x, prev_super, prev_lstm)


def forward_images(sd, cfg, x, prev_super, prev_lstm):
    return _encode(sd, cfg, 'images', x, prev_super, prev_lstm)


def forward_decoder(sd: StateDict, cfg: NetCfg, supers, return_logits: bool = False):
    """statenet.py:290-315: resblocks on S[-1]; dec0(x); dec_i(x + S[n-i-1]); pred; sigmoid."""
    P = 'statenetphasedrecurrent.'
    tup = (not bool(cfg.baseline)) and cfg.state_combination == 'convlstm'
    pick = (lambda s: s[0]) if tup else (lambda s: s)
    x = pick(supers[-1])
    for i in range(cfg.num_residual_blocks):
        x = residual_block(sd, f'{P}resblocks.{i}', x, cfg.norm)
    for i in range(cfg.num_encoders):
        if i > 0:
            if cfg.skip_type != 'sum':
                raise NotImplementedError('only skip_type="sum" is well-formed in the reference StateNet')
            x = x + pick(supers[cfg.num_encoders - i - 1])
        if cfg.use_upsample_conv:
            x = upsample_conv_layer(sd, f'{P}decoders.{i}', x, cfg.norm)
        else:
            x = transposed_conv_layer(sd, f'{P}decoders.{i}', x, cfg.norm)
    logits = conv_layer(sd, P + 'pred', x, 1, 0, relu=False, norm=cfg.norm, exact=True)
    return (torch.sigmoid(logits), logits) if return_logits else torch.sigmoid(logits)


# --------------------------------------------------------------------------
# a-10  ERGB2DepthRecurrent.forward     RAM_Net/model/model.py:141-219
# --------------------------------------------------------------------------
def zero_super_states(cfg: NetCfg, B: int, H: int, W: int):
    """model.py:146-159."""
    out = []
    for i in range(cfg.num_encoders):
        h, w, c = int(H / 2 ** (i + 1)), int(W / 2 ** (i + 1)), int(cfg.base * 2 ** (i + 1))
        z = torch.zeros(B, c, h, w)
        if not bool(cfg.baseline) and cfg.state_combination == 'convlstm':
            out.append([z, z.clone()])
        else:
            out.append(z)
    return out


def ergb2depth_recurrent(sd: StateDict, config: dict, item: dict, prev_super, prev_lstm: dict,
                         return_logits: bool = False):
    cfg = NetCfg(config)
    preds, supers_d, lstm_d, logits_d = {}, {}, {}, {}
    if prev_super is None:
        B, _, H, W = item['image'].shape
        prev_super = zero_super_states(cfg, B, H, W)
    bl = cfg.baseline
    events_as_images = bl == 'ergb0' or (bl == 'e' and cfg.loss_composition == 'image')
    last = None
    if (not bool(bl)) or events_as_images:
        if events_as_images:
            n, last = cfg.K - 1, prev_lstm['image']
        else:
            n, last = cfg.K, prev_lstm['events{}'.format(cfg.K - 1)]
        for k in range(n):
            key = 'events{}'.format(k)
            if bl == 'ergb0' or bl == 'e':
                s, l = forward_images(sd, cfg, item[key], prev_super, last)
            else:
                s, l = forward_events(sd, cfg, item[key], prev_super, last)
            r = forward_decoder(sd, cfg, s, True)
            preds[key], logits_d[key] = r
            supers_d[key], lstm_d[key] = s, l
            prev_super, last = s, l
    if (not bool(bl)) or bl == 'rgb' or (bl == 'e' and cfg.loss_composition != 'image'):
        last = prev_lstm['image']
    s, l = forward_images(sd, cfg, item['image'], prev_super, last)
    preds['image'], logits_d['image'] = forward_decoder(sd, cfg, s, True)
    supers_d['image'], lstm_d['image'] = s, l
    if return_logits:
        return preds, supers_d, lstm_d, logits_d
    return preds, supers_d, lstm_d


# --------------------------------------------------------------------------
# a-11  ERGB2Depth / UNet               model/model.py:79-111, model/unet.py:87-131
# --------------------------------------------------------------------------
def ergb2depth_unet(sd: StateDict, config: dict, item: dict, return_logits: bool = False):
    cfg = NetCfg(config)
    P = 'unet.'
    x = conv_layer(sd, P + 'head', item['image'], 1, 2)
    head = x
    blocks = []
    for i in range(cfg.num_encoders):
        x = conv_layer(sd, f'{P}encoders.{i}', x, 2, 2, norm=cfg.norm)
        blocks.append(x)
    for i in range(cfg.num_residual_blocks):
        x = residual_block(sd, f'{P}resblocks.{i}', x, cfg.norm)
    for i in range(cfg.num_encoders):
        skip = blocks[cfg.num_encoders - i - 1]                 # skip on EVERY decoder (unet.py:126-127)
        x = torch.cat([x, skip], 1) if cfg.skip_type == 'concat' else x + skip      # unet.py:11-16
        if cfg.use_upsample_conv:
            x = upsample_conv_layer(sd, f'{P}decoders.{i}', x, cfg.norm)
        else:
            x = transposed_conv_layer(sd, f'{P}decoders.{i}', x, cfg.norm)
    xh = torch.cat([x, head], 1) if cfg.skip_type == 'concat' else x + head
    logits = conv_layer(sd, P + 'pred', xh, 1, 0, relu=False, norm=cfg.norm, exact=True)   # unet.py:129
    pred = torch.sigmoid(logits)
    return ({'image': pred}, {'image': logits}) if return_logits else {'image': pred}


# --------------------------------------------------------------------------
# a-12  scale_invariant_loss            RAM_Net/model/loss.py:6-9
# --------------------------------------------------------------------------
def si_loss(pred: Tensor, target: Tensor, weight: float = 1.0, n_lambda: float = 1.0) -> Tensor:
    d = pred - target
    ok = ~torch.isnan(d)
    dv = d[ok]
    return weight * ((dv ** 2).mean() - n_lambda * dv.mean() ** 2)


def si_loss_grad(pred: Tensor, target: Tensor, weight: float = 1.0, n_lambda: float = 1.0) -> Tensor:
    """Analytic d loss / d pred = (2w/n)(d - lambda * mean(d)) on valid pixels, 0 at NaN."""
    d = pred - target
    ok = ~torch.isnan(d)
    n = ok.sum().to(pred.dtype)
    dz = torch.where(ok, d, torch.zeros_like(d))
    mean = dz.sum() / n
    return torch.where(ok, (2.0 * weight / n) * (d - n_lambda * mean), torch.zeros_like(d))


def spatial_gradient(p: Tensor) -> Tensor:
    """kornia.filters.spatial_gradient(p, mode='sobel', order=1, normalized=True) restated: [B,C,H,W] -> [B*C,2,H,W]
    (x derivative, y derivative), 3x3 Sobel / 8, replicate padding.  tests/test_oracle_golden.py cross-checks it against
    OpenCV's Sobel (an independent implementation of the same published operator), which is as far as it can be pinned
    without kornia."""
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]], dtype=p.dtype) / 8.0
    k = torch.stack([kx, kx.t()]).unsqueeze(1)                      # [2,1,3,3]
    b, c, h, w = p.shape
    return F.conv2d(F.pad(p.reshape(b * c, 1, h, w), [1, 1, 1, 1], mode='replicate'), k)   # [b*c,2,h,w]


def multi_scale_grad_loss(pred: Tensor, target: Tensor, start_scale: int = 1, num_scales: int = 4) -> Tensor:
    """MultiScaleGradient.forward (model/loss.py:33-63).  kornia.filters.spatial_gradient (kornia==0.4.0,
    requirements.txt:32, NOT vendored and not installed here) restated from its published behaviour: normalised
    3x3 Sobel (/8), replicate padding, output [B,C,2,H,W] (loss.py:54).  PARITY UNPINNED for the kornia part: no
    reference output exists to check it against in this container."""
    diff = pred - target
    B = target.shape[0]
    loss = 0
    for s in range(num_scales):
        p = F.avg_pool2d(diff, start_scale * 2 ** s, start_scale * 2 ** s)
        g = spatial_gradient(p)
        ok = ~torch.isnan(g)
        loss = loss + torch.abs(g[ok]).sum() / ok.sum() * B * 2
    return loss / num_scales


def sequence_loss(preds_per_step: List[Dict[str, Tensor]], targets_per_step: List[Dict[str, Tensor]],
                  loss_composition: Sequence[str], loss_weights: Sequence[float],
                  weight: float = 1.0, n_lambda: float = 1.0) -> Tensor:
    """Loss mixing of LSTMTrainer.forward_pass_sequence without the grad/mse terms
    (trainer/lstm_trainer.py:253,275-288,381-382,189-224): every key is bound to the SAME
    loss list and calculate_total_batch_loss runs once per key, hence the K_keys factor."""
    L = len(preds_per_step)
    terms, keys_seen = [], []
    for preds, tg in zip(preds_per_step, targets_per_step):
        for key, p in preds.items():
            if key in loss_composition:
                w = loss_weights[list(loss_composition).index(key)]
                if key not in keys_seen:
                    keys_seen.append(key)
                terms.append(w * si_loss(p, tg['depth_' + key], weight, n_lambda))
    return len(keys_seen) * sum(terms) / float(L)


# --------------------------------------------------------------------------
# a-14  Adam                            base/base_trainer.py:36-37 -> torch.optim.Adam
# --------------------------------------------------------------------------
def adam_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, step: int,
              lr: float = 3e-4, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8,
              weight_decay: float = 0.0):
    """torch.optim.Adam (non-amsgrad, L2 weight decay added to the gradient), fp32.
    step is 1-based.  Returns (p, m, v)."""
    p, g, m, v = (np.asarray(a, np.float32) for a in (p, g, m, v))
    if weight_decay != 0.0:
        g = g + np.float32(weight_decay) * p
    m = np.float32(b1) * m + np.float32(1 - b1) * g
    v = np.float32(b2) * v + np.float32(1 - b2) * g * g
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / np.float32(math.sqrt(bc2)) + np.float32(eps)
    p = p - np.float32(step_size) * (m / denom)
    return p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)


# --------------------------------------------------------------------------
# synthetic inputs: the generators live in the product package
# (bench.py's GPU arm must not import oracle/); re-exported for the tests.
# --------------------------------------------------------------------------
from rpg_ramnet_b200.utils.synthetic import synth_events, synth_sequence  # noqa: E402,F401


def scale_weights(model, s: float):
    """'stress init' (SURVEY §7): every 4-D conv weight x s; BN stats randomised
    so that eval-mode BN is not the identity."""
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        for n, p in model.state_dict().items():
            if n.endswith('weight') and p.dim() == 4 and s != 1.0:
                p.mul_(s)
            if n.endswith('running_mean'):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            if n.endswith('running_var'):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            if '.norm_layer.weight' in n or '.bn1.weight' in n or '.bn2.weight' in n:
                p.copy_(0.75 + 0.5 * torch.rand(p.shape, generator=g))
            if '.norm_layer.bias' in n or '.bn1.bias' in n or '.bn2.bias' in n:
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
