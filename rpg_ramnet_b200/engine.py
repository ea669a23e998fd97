"""Host-side executor: packs module parameters for the CUDA kernels and issues the fused graph.

This is the "plumbing" half of the drop-in: it decides WHICH C-ABI calls make up a RAM-Net pass
and in what order, owns the packed-weight cache, and keeps activations pixel-major (NHWC) between
kernels.  All arithmetic happens inside libramnet_sm100a.so.

Fusion map (reference op -> kernel), see SURVEY.md §2b:
  ConvLayer / head (Cin<=8)        -> ramnet_head_conv            (conv+bias+ReLU, NCHW in, NHWC out)
  ConvLayer 5x5 s2                 -> ramnet_conv_fwd EPI_BIAS_RELU
  ConvGRU (cat, 3 convs, 8 pointwise) -> 2 x ramnet_conv_fwd: EPI_GRU_RU over [x|h], EPI_GRU_OUT over [x|h*r]
  ConvLSTM (cat, conv, 9 pointwise)-> 1 x ramnet_conv_fwd EPI_LSTM over [x|h], gate-interleaved columns
  ResidualBlock                    -> 2 x ramnet_conv_fwd (EPI_BIAS_RELU, EPI_BIAS_RES_RELU)
  skip_sum + bilinear x2           -> ramnet_upsample2x_add
  UpsampleConvLayer conv           -> ramnet_conv_fwd EPI_BIAS_RELU
  pred + sigmoid                   -> ramnet_pred_sigmoid
  eval-mode BatchNorm              -> folded into the packed weights/bias (no kernel)
"""
import os
from typing import Optional

import torch

from . import ops
from ._lib import RamnetError

_KINDS = {'fp32': ops.MMA_FP32, 'tf32': ops.MMA_TF32}
DEFAULT_MMA_KIND = os.environ.get('RAMNET_MMA_KIND', 'tf32')


def resolve_mma_kind(name: Optional[str]) -> int:
    name = (name or DEFAULT_MMA_KIND).lower()
    if name not in _KINDS:
        raise RamnetError(f"mma_kind must be one of {sorted(_KINDS)}, got {name!r}")
    return _KINDS[name]


class Packed:
    __slots__ = ('w', 'b', 'Cout', 'ksize', 'stride', 'token')


def _token(tensors, extra):
    return tuple((t.data_ptr(), t._version, t.device) for t in tensors if t is not None) + tuple(extra)


def _fold_norm(w, b, norm_mod, norm_kind, training):
    """Eval-mode BatchNorm2d / InstanceNorm2d(track_running_stats=True) folded into (w, b)."""
    if norm_kind not in ('BN', 'IN') or norm_mod is None:
        return w, b
    if training:
        raise RamnetError("norm='%s' in training mode (batch statistics) is not implemented; "
                          "call model.eval() or use norm='none' (all shipped configs)" % norm_kind)
    if getattr(norm_mod, 'running_mean', None) is None:
        raise RamnetError('InstanceNorm2d without running statistics is not implemented')
    scale = torch.rsqrt(norm_mod.running_var.float() + norm_mod.eps)
    shift = -norm_mod.running_mean.float() * scale
    if getattr(norm_mod, 'weight', None) is not None:
        scale = scale * norm_mod.weight.float()
        shift = shift * norm_mod.weight.float() + norm_mod.bias.float()
    w2 = w * scale.view(-1, 1, 1, 1)
    b2 = shift if b is None else b * scale + shift
    return w2, b2


class WeightCache:
    """Packed weights per fused layer, rebuilt when a source parameter changes
    (optimizer step, load_state_dict, .to())."""

    def __init__(self):
        self._cache = {}

    def get(self, key, sources, extra, builder) -> Packed:
        tok = _token(sources, extra)
        hit = self._cache.get(key)
        if hit is not None and hit.token == tok:
            return hit
        with torch.no_grad():
            p = builder()
        p.token = tok
        self._cache[key] = p
        return p

    def clear(self):
        self._cache.clear()


def _norm_sources(nm):
    if nm is None:
        return []
    return [getattr(nm, a, None) for a in ('running_mean', 'running_var', 'weight', 'bias')]


def pack_conv(cache: WeightCache, key, conv, kind: int, norm_mod=None, norm_kind=None, training=False) -> Packed:
    """nn.Conv2d (+ folded eval norm) -> Packed for ramnet_conv_fwd."""
    def build():
        w, b = conv.weight.detach().float(), None if conv.bias is None else conv.bias.detach().float()
        w, b = _fold_norm(w, b, norm_mod, norm_kind, training)
        p = Packed()
        p.w = ops.pack_weights(w, kind)
        p.b = None if b is None else b.contiguous()
        p.Cout, p.ksize, p.stride = w.shape[0], w.shape[2], conv.stride[0]
        return p
    return cache.get(key, [conv.weight, conv.bias] + _norm_sources(norm_mod), (kind, training), build)


def pack_head(cache: WeightCache, key, conv) -> Packed:
    """Head conv keeps nn.Conv2d's [Cout,Cin,5,5] layout (ramnet_head_conv reads it directly)."""
    def build():
        p = Packed()
        p.w = conv.weight.detach().float().contiguous()
        p.b = None if conv.bias is None else conv.bias.detach().float().contiguous()
        p.Cout, p.ksize, p.stride = conv.weight.shape[0], conv.weight.shape[2], 1
        return p
    return cache.get(key, [conv.weight, conv.bias], (), build)


def pack_gru(cache: WeightCache, key, gru, kind: int):
    """ConvGRU -> (RU pack with Cout=2C: [reset | update], OUT pack)."""
    def build_ru():
        w = torch.cat([gru.reset_gate.weight.detach(), gru.update_gate.weight.detach()], 0).float()
        b = torch.cat([gru.reset_gate.bias.detach(), gru.update_gate.bias.detach()], 0).float()
        p = Packed()
        p.w, p.b = ops.pack_weights(w, kind), b.contiguous()
        p.Cout, p.ksize, p.stride = w.shape[0], w.shape[2], 1
        return p
    ru = cache.get(key + '/ru', [gru.reset_gate.weight, gru.reset_gate.bias, gru.update_gate.weight,
                                 gru.update_gate.bias], (kind,), build_ru)
    out = pack_conv(cache, key + '/out', gru.out_gate, kind)
    return ru, out


def pack_lstm(cache: WeightCache, key, lstm, kind: int) -> Packed:
    """ConvLSTM Gates -> columns interleaved 4c+g (in, remember, out, cell)."""
    def build():
        w = lstm.Gates.weight.detach().float()
        C = w.shape[0] // 4
        p = Packed()
        p.w = ops.pack_weights(w, kind, lstm_interleave=True)
        p.b = lstm.Gates.bias.detach().float().view(4, C).t().contiguous().view(-1)
        p.Cout, p.ksize, p.stride = w.shape[0], w.shape[2], 1
        return p
    return cache.get(key, [lstm.Gates.weight, lstm.Gates.bias], (kind,), build)


# ------------------------------------------------------------------------------------------
# fused blocks
# ------------------------------------------------------------------------------------------
def run_conv(x, p: Packed, epilogue, kind, x1=None, aux0=None, aux1=None, round_out=False):
    return ops.conv_fwd(x, x1, p.w, p.b, p.Cout, p.ksize, p.stride, epilogue, kind, aux0=aux0, aux1=aux1,
                        round_tf32=round_out and kind == ops.MMA_TF32)


def run_gru(x, h, ru: Packed, out: Packed, kind):
    """submodules.py:436-454 as two fused convolutions; returns h'."""
    if h is None:
        h = ops.zeros_nhwc(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device)
    if h.shape != x.shape:
        raise RamnetError(f'ConvGRU: state shape {tuple(h.shape)} does not match input {tuple(x.shape)} '
                          '(H and W must be divisible by 2**num_encoders)')
    u, rh = run_conv(x, ru, ops.EPI_GRU_RU, kind, x1=h, aux0=h, round_out=True)
    return run_conv(x, out, ops.EPI_GRU_OUT, kind, x1=rh, aux0=h, aux1=u, round_out=True)


def run_lstm(x, state, p: Packed, kind):
    """submodules.py:318-358 as one fused convolution; returns (h', c')."""
    if state is None:
        h = ops.zeros_nhwc(x.shape[0], p.Cout // 4, x.shape[2], x.shape[3], x.device)
        c = ops.zeros_nhwc(x.shape[0], p.Cout // 4, x.shape[2], x.shape[3], x.device)
    else:
        h, c = ops.as_nhwc(state[0]), ops.as_nhwc(state[1])
    if h.shape[2:] != x.shape[2:]:
        raise RamnetError(f'ConvLSTM: state shape {tuple(h.shape)} does not match input {tuple(x.shape)}')
    return run_conv(x, p, ops.EPI_LSTM, kind, x1=h, aux0=c, round_out=True)
