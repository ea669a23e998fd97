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


# bumped by anything that rewrites parameters behind autograd's back (the fused Adam kernel)
_WEIGHT_EPOCH = 0


def bump_weight_epoch():
    global _WEIGHT_EPOCH
    _WEIGHT_EPOCH += 1


def _token(tensors, extra):
    return tuple((t.data_ptr(), t._version, t.device) for t in tensors if t is not None) + tuple(extra) + (_WEIGHT_EPOCH,)


def _fold_norm(w, b, norm_mod, norm_kind, training):
    """Eval-mode BatchNorm2d / InstanceNorm2d(track_running_stats=True) folded into (w, b)."""
    if norm_kind not in ('BN', 'IN') or norm_mod is None:
        return w, b
    if norm_mod.training or getattr(norm_mod, 'running_mean', None) is None:
        raise RamnetError("internal: a norm layer with live statistics reached the weight folding (engine.norm_is_live "
                          "routes those through ops.norm_fwd)")
    scale = torch.rsqrt(norm_mod.running_var.float() + norm_mod.eps)
    shift = -norm_mod.running_mean.float() * scale
    if getattr(norm_mod, 'weight', None) is not None:
        scale = scale * norm_mod.weight.float()
        shift = shift * norm_mod.weight.float() + norm_mod.bias.float()
    w2 = w * scale.view(-1, 1, 1, 1)
    b2 = shift if b is None else b * scale + shift
    return w2, b2


def _norm_params(nm):
    return (getattr(nm, 'weight', None), getattr(nm, 'bias', None)) if nm is not None else (None, None)


def norm_is_live(norm_mod, norm_kind, *grad_tensors) -> bool:
    """True when the layer's norm cannot be folded into the conv weights and runs as its own kernels
    (ops.norm_fwd): batch / instance statistics (module in train mode, or an InstanceNorm2d built without running
    statistics: the ResidualBlock's, submodules.py:192-194), or gradients are wanted through it."""
    if norm_kind not in ('BN', 'IN') or norm_mod is None:
        return False
    return bool(norm_mod.training) or getattr(norm_mod, 'running_mean', None) is None or \
        needs_grad(*grad_tensors, *_norm_params(norm_mod))


def norm_act(z, norm_mod, norm_kind, act, res=None, round_out=False):
    """act(norm(z) (+ res)) with live statistics; differentiable when anything upstream wants gradients."""
    batch_stats = bool(norm_mod.training) or getattr(norm_mod, 'running_mean', None) is None
    gamma, beta = _norm_params(norm_mod)
    if needs_grad(z, res, gamma, beta):
        from .autograd import NormActFn
        return NormActFn.apply(z, res, gamma, beta, norm_mod, norm_kind, act, batch_stats, round_out)
    rm, rv = getattr(norm_mod, 'running_mean', None), getattr(norm_mod, 'running_var', None)
    y, _ = ops.norm_fwd(z, norm_kind, act, None if gamma is None else gamma.detach().float().contiguous(),
                        None if beta is None else beta.detach().float().contiguous(), res, rm, rv,
                        norm_mod.momentum if norm_mod.momentum is not None else 0.1, norm_mod.eps, batch_stats, round_out)
    if rm is not None and batch_stats and norm_kind == 'BN' and getattr(norm_mod, 'num_batches_tracked', None) is not None:
        norm_mod.num_batches_tracked.add_(1)
    return y


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


def pack_conv(cache: WeightCache, key, conv, kind: int, norm_mod=None, norm_kind=None, training=False,
              hpack_ok=False) -> Packed:
    """nn.Conv2d (+ folded eval norm) -> Packed for ramnet_conv_fwd.  `hpack_ok`: the caller's epilogue is one the
    horizontal-tap-packed kernel implements (bias / relu / residual / fused pred), so eligible layers may use it."""
    def build():
        w, b = conv.weight.detach().float(), None if conv.bias is None else conv.bias.detach().float()
        w, b = _fold_norm(w, b, norm_mod, norm_kind, training)
        p = Packed()
        hp = hpack_ok and ops.hpack_eligible(w.shape[0], w.shape[2], conv.stride[0], kind)
        if hpack_ok and ops.s2seg_eligible(w.shape[1], w.shape[0], w.shape[2], conv.stride[0], kind):
            p.w = ops.pack_weights_s2seg(w)
        else:
            p.w = ops.pack_weights_hpack(w) if hp else ops.pack_weights(w, kind)
        p.b = None if b is None else b.contiguous()
        p.Cout, p.ksize, p.stride = w.shape[0], w.shape[2], conv.stride[0]
        return p
    hp_key = hpack_ok and (ops.hpack_eligible(conv.weight.shape[0], conv.weight.shape[2], conv.stride[0], kind) or
                           ops.s2seg_eligible(conv.weight.shape[1], conv.weight.shape[0], conv.weight.shape[2], conv.stride[0], kind))
    return cache.get(key, [conv.weight, conv.bias] + _norm_sources(norm_mod), (kind, training, hp_key), build)


def pack_upconv(cache: WeightCache, key, conv, norm_mod=None, norm_kind=None, training=False) -> Packed:
    """Decoder nn.Conv2d (+ folded eval norm) -> collapsed up-conv taps for ops.conv_up_fwd (RAMNET_FLAG_UPCONV)."""
    def build():
        w, b = conv.weight.detach().float(), None if conv.bias is None else conv.bias.detach().float()
        w, b = _fold_norm(w, b, norm_mod, norm_kind, training)
        p = Packed()
        p.w = ops.pack_weights_upconv(w)
        p.b = None if b is None else b.contiguous()
        p.Cout, p.ksize, p.stride = w.shape[0], w.shape[2], 1
        return p
    return cache.get(key + '/up', [conv.weight, conv.bias] + _norm_sources(norm_mod), ('upconv', training), build)


def pack_head(cache: WeightCache, key, conv, tc: bool = False) -> Packed:
    """Head conv keeps nn.Conv2d's [Cout,Cin,5,5] layout (ramnet_head_conv reads it directly)."""
    def build():
        p = Packed()
        p.w = conv.weight.detach().float().contiguous()
        p.b = None if conv.bias is None else conv.bias.detach().float().contiguous()
        p.Cout, p.ksize, p.stride = conv.weight.shape[0], conv.weight.shape[2], 1
        if tc:      # tensor-core path: [r][Cout][dx*Cin + ci]
            p.w = ops.pack_weights_head(p.w)
        return p
    return cache.get(key, [conv.weight, conv.bias], (tc,), build)


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
def run_conv(x, p: Packed, epilogue, kind, x1=None, aux0=None, aux1=None, round_out=False, out0=None, out1=None):
    return ops.conv_fwd(x, x1, p.w, p.b, p.Cout, p.ksize, p.stride, epilogue, kind, aux0=aux0, aux1=aux1,
                        round_tf32=round_out and kind == ops.MMA_TF32, out0=out0, out1=out1)


def run_gru(x, h, ru: Packed, out: Packed, kind, out_h=None):
    """submodules.py:436-454 as two fused convolutions; returns h' (written into `out_h` when given)."""
    if h is None:
        h = ops.zeros_nhwc(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device)
    if h.shape != x.shape:
        raise RamnetError(f'ConvGRU: state shape {tuple(h.shape)} does not match input {tuple(x.shape)} '
                          '(H and W must be divisible by 2**num_encoders)')
    u, rh = run_conv(x, ru, ops.EPI_GRU_RU, kind, x1=h, aux0=h, round_out=True)
    return run_conv(x, out, ops.EPI_GRU_OUT, kind, x1=rh, aux0=h, aux1=u, round_out=True, out0=out_h)


def run_lstm(x, state, p: Packed, kind, out_state=None):
    """submodules.py:318-358 as one fused convolution; returns (h', c') (written into `out_state` when given)."""
    if state is None:
        h = ops.zeros_nhwc(x.shape[0], p.Cout // 4, x.shape[2], x.shape[3], x.device)
        c = ops.zeros_nhwc(x.shape[0], p.Cout // 4, x.shape[2], x.shape[3], x.device)
    else:
        h, c = ops.as_nhwc(state[0]), ops.as_nhwc(state[1])
    if h.shape[2:] != x.shape[2:]:
        raise RamnetError(f'ConvLSTM: state shape {tuple(h.shape)} does not match input {tuple(x.shape)}')
    o0, o1 = (None, None) if out_state is None else out_state
    return run_conv(x, p, ops.EPI_LSTM, kind, x1=h, aux0=c, round_out=True, out0=o0, out1=o1)


# ------------------------------------------------------------------------------------------
# layer-level entry points: inference launches the kernels directly, training (grad mode) routes
# through the autograd.Function wrappers whose backward is our own kernels (autograd.py)
# ------------------------------------------------------------------------------------------
def needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def _no_fold_in_training(norm_mod, norm_kind):
    if norm_kind in ('BN', 'IN') and norm_mod is not None:
        raise RamnetError('gradients through folded BatchNorm/InstanceNorm are not implemented '
                          "(all shipped configs use norm='none')")


def head_layer(cache, key, conv, x, tf32):
    if needs_grad(conv.weight, conv.bias):
        from .autograd import HeadConvFn
        return HeadConvFn.apply(x.float().contiguous(), conv.weight, conv.bias, tf32)
    Cout, Cin = conv.weight.shape[0], conv.weight.shape[1]
    if tf32 and conv.weight.shape[2] == 5 and ops.head_tc_ok(Cin, Cout):
        hp = pack_head(cache, key, conv, tc=True)
        return ops.head_conv_tc(ops.head_im2row(x.float()), hp.w, hp.b, Cin, Cout, round_tf32=True)
    hp = pack_head(cache, key, conv)
    return ops.head_conv(x.float(), hp.w, hp.b, round_tf32=tf32)


def conv_layer(cache, key, conv, kind, x, epilogue, x1=None, res=None, norm_mod=None, norm_kind=None, training=False,
               round_out=False):
    if norm_is_live(norm_mod, norm_kind, x, x1, res, conv.weight, conv.bias):
        # conv (+ bias) -> norm kernels (statistics, normalise + activation (+ residual)): submodules.py:26-35, 200-215
        if epilogue not in (ops.EPI_BIAS, ops.EPI_BIAS_RELU, ops.EPI_BIAS_RES_RELU):
            raise RamnetError('internal: live norm after a fused epilogue that has no norm in the reference')
        p = pack_conv(cache, key + '/raw', conv, kind, hpack_ok=x1 is None)
        if needs_grad(x, x1, conv.weight, conv.bias):
            from .autograd import ConvFn
            z = ConvFn.apply(x, x1, None, conv.weight, conv.bias, p.w, ops.EPI_BIAS, kind, p.stride, False)
        else:
            z = run_conv(x, p, ops.EPI_BIAS, kind, x1=x1)
        return norm_act(z, norm_mod, norm_kind, None if epilogue == ops.EPI_BIAS else 'relu',
                        res=res if epilogue == ops.EPI_BIAS_RES_RELU else None,
                        round_out=round_out and kind == ops.MMA_TF32)
    p = pack_conv(cache, key, conv, kind, norm_mod, norm_kind, training,
                  hpack_ok=(epilogue in (ops.EPI_BIAS, ops.EPI_BIAS_RELU, ops.EPI_BIAS_RES_RELU) and x1 is None and
                            not (conv.stride[0] == 2 and epilogue == ops.EPI_BIAS_RES_RELU)))
    if needs_grad(x, x1, res, conv.weight, conv.bias):
        _no_fold_in_training(norm_mod, norm_kind)
        from .autograd import ConvFn
        return ConvFn.apply(x, x1, res, conv.weight, conv.bias, p.w, epilogue, kind, p.stride,
                            round_out and kind == ops.MMA_TF32)
    return run_conv(x, p, epilogue, kind, x1=x1, aux0=res, round_out=round_out)


def gru_layer(cache, key, gru, kind, x, h, out_h=None):
    ru, out = pack_gru(cache, key, gru, kind)
    params = (gru.reset_gate.weight, gru.reset_gate.bias, gru.update_gate.weight, gru.update_gate.bias,
              gru.out_gate.weight, gru.out_gate.bias)
    if needs_grad(x, h, *params):
        from .autograd import GruFn
        if h is None:
            h = ops.zeros_nhwc(x.shape[0], x.shape[1], x.shape[2], x.shape[3], x.device)
        if h.shape != x.shape:
            raise RamnetError(f'ConvGRU: state shape {tuple(h.shape)} does not match input {tuple(x.shape)}')
        return GruFn.apply(x, h, *params, ru, out, kind)
    return run_gru(x, h, ru, out, kind, out_h=out_h)


def lstm_layer(cache, key, lstm, kind, x, state, out_state=None):
    p = pack_lstm(cache, key, lstm, kind)
    st = None if state is None else (state[0], state[1])
    if needs_grad(x, lstm.Gates.weight, lstm.Gates.bias, *(st or ())):
        from .autograd import LstmFn
        C = p.Cout // 4
        if st is None:
            h = ops.zeros_nhwc(x.shape[0], C, x.shape[2], x.shape[3], x.device)
            c = ops.zeros_nhwc(x.shape[0], C, x.shape[2], x.shape[3], x.device)
        else:
            h, c = ops.as_nhwc(st[0]), ops.as_nhwc(st[1])
        if h.shape[2:] != x.shape[2:]:
            raise RamnetError(f'ConvLSTM: state shape {tuple(h.shape)} does not match input {tuple(x.shape)}')
        return LstmFn.apply(x, h, c, lstm.Gates.weight, lstm.Gates.bias, p, kind)
    return run_lstm(x, state, p, kind, out_state=out_state)


def transposed_conv_layer(cache, key, tconv, kind, x, skip=None, norm_mod=None, norm_kind=None, training=False):
    """TransposedConvLayer.forward (submodules.py:38-66): ConvTranspose2d(k=5, stride 2, padding 2, output_padding 1)
    + bias + ReLU.  A transposed convolution IS the data gradient of the stride-2 convolution with the same weight
    tensor, so it runs as zero-insertion + the forward tensor-core kernel on tap-flipped, channel-transposed weights."""
    w = tconv.weight                      # [Cin, Cout, k, k] == nn.Conv2d layout of the conv it is the adjoint of
    has_norm = norm_kind in ('BN', 'IN') and norm_mod is not None
    live = norm_is_live(norm_mod, norm_kind, x, skip, w, tconv.bias)
    if tuple(tconv.stride) != (2, 2) or tuple(tconv.output_padding) != (1, 1) or \
            tuple(tconv.padding) != (w.shape[2] // 2,) * 2:
        raise RamnetError('TransposedConvLayer: only stride 2, padding k//2, output_padding 1 is implemented')
    Cin, Cout = w.shape[0], w.shape[1]

    def build():
        wf, bf = w.detach().float(), None if tconv.bias is None else tconv.bias.detach().float()
        if has_norm and not live:         # eval-mode norm: per-OUTPUT-channel scale (dim 1 of a ConvTranspose2d weight)
            w2, bf = _fold_norm(wf.transpose(0, 1), bf, norm_mod, norm_kind, False)
            wf = w2.transpose(0, 1).contiguous()
        p = Packed()
        p.w = ops.pack_weights_dgrad(wf, kind, 0, Cout)
        p.b = None if bf is None else bf.contiguous()
        p.Cout, p.ksize, p.stride = Cout, w.shape[2], 1
        return p
    p = cache.get(key, [w, tconv.bias] + (_norm_sources(norm_mod) if has_norm else []), (kind, 'tconv', live), build)
    if live:
        if needs_grad(x, skip, w, tconv.bias):
            from .autograd import TransposedConvFn
            z = TransposedConvFn.apply(x, skip, w, tconv.bias, p.w, kind, False)
        else:
            N, _, H, W = x.shape
            up = ops.zero_insert2x(x, 2 * H, 2 * W, skip=skip)
            z = ops.conv_fwd(up, None, p.w, p.b, Cout, p.ksize, 1, ops.EPI_BIAS, kind)
        return norm_act(z, norm_mod, norm_kind, 'relu')
    if needs_grad(x, skip, w, tconv.bias):
        from .autograd import TransposedConvFn
        return TransposedConvFn.apply(x, skip, w, tconv.bias, p.w, kind)
    N, _, H, W = x.shape
    up = ops.zero_insert2x(x, 2 * H, 2 * W, skip=skip)      # skip sum fused into the zero insertion
    return ops.conv_fwd(up, None, p.w, p.b, Cout, p.ksize, 1, ops.EPI_BIAS_RELU, kind)


def upsample_add(x, skip, tf32):
    if needs_grad(x, skip):
        from .autograd import UpsampleAddFn
        return UpsampleAddFn.apply(x, skip, tf32)
    return ops.upsample2x_add(x, skip, round_tf32=tf32)


def pred_layer(x, pred_conv, norm_mod, norm_kind, training, return_logits=False, skip=None, concat=False):
    """pred (+ norm) + sigmoid over x (+ skip); concat: over cat([x, skip]) (UNet with skip_type='concat', unet.py:129)."""
    w = pred_conv.weight
    b = pred_conv.bias
    if norm_is_live(norm_mod, norm_kind, x, skip, w, b):      # pred conv -> norm -> sigmoid (statenet.py:116-117,313)
        if needs_grad(x, skip, w, b):
            from .autograd import PredLogitsFn
            z = PredLogitsFn.apply(x, skip, w, b, concat)
        else:
            z = ops.pred_logits(x, skip, w, None if b is None else b.detach().float(), concat)
        logits = None
        if return_logits:                 # debugging aid (inference): the normalised logits, running statistics untouched
            gamma, beta = _norm_params(norm_mod)
            bs = bool(norm_mod.training) or getattr(norm_mod, 'running_mean', None) is None
            logits, _ = ops.norm_fwd(z.detach(), norm_kind, None, None if gamma is None else gamma.detach().float().contiguous(),
                                     None if beta is None else beta.detach().float().contiguous(), None,
                                     None if bs else norm_mod.running_mean, None if bs else norm_mod.running_var,
                                     0.0, norm_mod.eps, bs)
        depth = norm_act(z, norm_mod, norm_kind, 'sigmoid')
        return (depth, logits) if return_logits else depth
    if needs_grad(x, skip, w, b):
        _no_fold_in_training(norm_mod, norm_kind)
        if return_logits:
            raise RamnetError('return_logits is an inference-only debugging aid')
        from .autograd import PredFn
        return PredFn.apply(x, skip, w, b, concat)
    wf, bf = _fold_norm(w.detach().float(), None if b is None else b.detach().float(), norm_mod, norm_kind, training)
    return ops.pred_sigmoid(x, skip, wf, bf, want_logits=return_logits, concat=concat)


# ------------------------------------------------------------------------------------------
# CUDA-graph runner (inference): one captured graph per (pass type, state direction)
# ------------------------------------------------------------------------------------------
class GraphRunner:
    """Replays a RAM-Net pass as TWO CUDA-graph launches instead of ~30 Python-issued kernel launches (the reference
    issues ~70 per pass): the FRONT (head -> encoders -> state update, statenet.py:204-288) and the BACK
    (residual blocks -> decoders -> depth, statenet.py:290-315).  The back of a pass reads nothing but the super states
    the front wrote, and the front of the NEXT pass reads those same states and writes the other buffer set -- so the
    next front runs on a second stream UNDER the current back (round 2).  Every conv kernel is a persistent grid of
    one CTA (pair) per SM, so the overlap does not share SMs: it fills the SMs a kernel leaves idle in its last wave
    and during its epilogue tail (at levels 1-2 a layer is 64-128 work items on 74 CTA pairs).
    RAMNET_PASS_OVERLAP=0 keeps everything on the caller's stream.

    Recurrent state ping-pongs between two persistent buffer sets so no state copy is ever made: the front for
    direction s reads set s and writes set 1-s.

    Aliasing contract: the state tensors returned by a pass ARE these persistent buffers; they stay
    valid until the second-next pass overwrites them (callers that carry only the latest state —
    trainer/lstm_trainer.py:380, test.py:380 — are unaffected).  A caller that hands back the OLDER of the two
    sets (state kept for two or more passes, already overwritten) gets a RamnetError instead of silently reading
    recycled memory; clone the states to keep them longer.  Depth maps are returned as fresh tensors.

    `nxt` (run): the pass that WILL follow, when the caller knows it (ERGB2DepthRecurrent.forward does, within one
    item): its front is launched before this pass's back.  Device inputs are copied into the staging slot after
    everything already queued on the caller's stream (they may just have been produced there) unless the owner sets
    `inputs_static` (inputs resident and not written by queued work: the copy then does not wait for the caller's
    stream and the first front of an item can run under the last back of the previous one); host inputs come over
    the copy stream and never wait for it."""

    def __init__(self, net, B, H, W, device):
        self.net, self.B, self.H, self.W, self.device = net, B, H, W, device
        self.pair = (not bool(net.baseline)) and net.state_combination == 'convlstm'
        self.sets = [self._alloc_states(), self._alloc_states()]
        self.x_in, self.graphs = {}, {}
        self.param_sig = None
        self.pools = {'front': None, 'back': None}   # fronts never overlap fronts, backs never overlap backs
        # host inputs: staged H2D on a copy stream into a 2-slot ring per pass type, so the copy of the next pass
        # runs under the kernels of the current one (the reference copies on the compute stream, model.py:177,200)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.overlap = os.environ.get('RAMNET_PASS_OVERLAP', '1') != '0'
        self.front_stream = torch.cuda.Stream(device=device) if self.overlap else None
        self.inputs_static = False
        self.slot_ctr = {}
        self.last_dst = None        # buffer set the most recent pass wrote (the only one a caller may hand back)
        self.back_done = [None, None]   # event: the last back that read set k has finished (before a front rewrites it)
        self.pending = None         # front launched ahead for the announced next pass

    def _alloc_states(self):
        out = []
        n = self.net
        for i in range(n.num_encoders):
            c, h, w = n.base_num_channels << (i + 1), self.H >> (i + 1), self.W >> (i + 1)
            if self.pair:
                out.append([ops.zeros_nhwc(self.B, c, h, w, self.device), ops.zeros_nhwc(self.B, c, h, w, self.device)])
            else:
                out.append(ops.zeros_nhwc(self.B, c, h, w, self.device))
        return out

    @staticmethod
    def _flat(states):
        flat = []
        for s in states:
            flat.extend(s if isinstance(s, (list, tuple)) else [s])
        return flat

    def _which_set(self, prev):
        if prev is None:
            return None
        flat = self._flat(prev)
        for k in (0, 1):
            mine = self._flat(self.sets[k])
            if len(mine) == len(flat) and all(a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape)
                                             for a, b in zip(flat, mine)):
                return k
        return None

    def _sig(self):
        return tuple((p.data_ptr(), p._version) for p in self.net.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in self.net.buffers()) + \
            (self.net.training, self.net._kind(), _WEIGHT_EPOCH)

    def run(self, which, x, prev_super, nxt=None):
        cur = torch.cuda.current_stream(self.device)
        sig = self._sig()
        if sig != self.param_sig:           # weights changed (optimizer step / load_state_dict): re-capture
            if self.pending is not None:
                raise RamnetError('cuda_graphs: the weights changed between two passes of one item')
            self.graphs.clear()
            self.param_sig = sig
        src = self._which_set(prev_super)
        fr, self.pending = self.pending, None
        if fr is not None:
            if fr['which'] != which or fr['x'] is not x or src != fr['src']:
                raise RamnetError('cuda_graphs: the pass announced to the previous run() call (nxt=) is not the one being '
                                  'run; its front has already been launched on the announced input and states')
        else:
            if src is not None and self.last_dst is not None and src != self.last_dst:
                raise RamnetError('cuda_graphs: the states passed in are the runner\'s buffer set that the previous pass has '
                                  'already overwritten (states returned by a pass are valid for one further pass only); '
                                  'clone() states that must live longer')
            foreign = src is None
            if foreign:                     # foreign or initial state: bring it into set 0 (on the caller's stream)
                src = 0
                if self.back_done[0] is not None:
                    cur.wait_event(self.back_done[0])
                mine = self._flat(self.sets[0])
                if prev_super is None:
                    for t in mine:
                        t.zero_()
                else:
                    for t, f in zip(mine, self._flat(prev_super)):
                        t.copy_(f)
            fr = self._launch_front(which, x, src, after_caller=foreign)
        dst = fr['dst']
        if nxt is not None and self.overlap:
            self.pending = self._launch_front(nxt[0], nxt[1], dst)
        if fr['event'] is not None:
            cur.wait_event(fr['event'])
        key = ('back', dst)
        entry = self.graphs.get(key)
        if entry is None:
            entry = self.graphs[key] = self._capture_back(dst)
        graph, pred = entry
        graph.replay()
        if self.overlap:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.back_done[dst] = ev
        self.last_dst = dst
        return self.sets[dst], pred.clone()

    def _launch_front(self, which, x, src, after_caller=False):
        """Stages the input and replays the front graph of a pass reading buffer set `src` (on the front stream when
        passes overlap).  Returns what run() needs to finish the pass."""
        cur = torch.cuda.current_stream(self.device)
        dst = 1 - src
        slot = self.stage(which, x)
        if slot is None:                    # ring full of inputs staged ahead but never consumed: drop them
            for sl in self.x_in[which]:
                sl['staged_for'] = None
            slot = self.stage(which, x)
        key = ('front', which, src, slot['idx'])
        entry = self.graphs.get(key)
        if entry is None:
            entry = self.graphs[key] = self._capture_front(which, slot['buf'], src, dst)
        fs = self.front_stream if self.overlap else cur
        if fs is not cur and after_caller:
            fs.wait_stream(cur)             # the state was copied in on the caller's stream
        if slot['ready'] is not None:
            fs.wait_event(slot['ready'])
            slot['ready'] = None
        if fs is not cur and self.back_done[dst] is not None:
            fs.wait_event(self.back_done[dst])      # the back of the second-last pass still reads the set written here
        event = None
        with torch.cuda.stream(fs):
            entry[0].replay()
            slot['done'].record(fs)         # the staging slot may be overwritten once this front has read it
            if fs is not cur:
                event = torch.cuda.Event()
                event.record(fs)
        slot['staged_for'] = None
        return {'which': which, 'x': x, 'src': src, 'dst': dst, 'event': event}

    def stage(self, which, x):
        """Brings the input of the next `which` pass into a staging slot and returns the slot.  Host tensors go
        over the copy stream (asynchronously when pinned); device tensors are copied on the stream the front runs on.
        Idempotent per tensor: forward() stages the inputs of later passes ahead of time."""
        ring = self.x_in.get(which)
        if ring is None or ring[0]['buf'].shape != x.shape:
            ring = [{'buf': torch.empty(tuple(x.shape), dtype=torch.float32, device=self.device), 'idx': i,
                     'ready': None, 'done': torch.cuda.Event(), 'staged_for': None} for i in range(2)]
            self.x_in[which] = ring
            self.slot_ctr[which] = 0
            self.graphs = {k: v for k, v in self.graphs.items() if not (k[0] == 'front' and k[1] == which)}
        for slot in ring:
            if slot['staged_for'] is x:
                return slot
        slot = ring[self.slot_ctr[which] & 1]
        if slot['staged_for'] is not None:      # both slots hold inputs that have not been consumed yet
            return None
        self.slot_ctr[which] += 1
        slot['staged_for'] = x
        if x.is_cuda:
            cur = torch.cuda.current_stream(self.device)
            if self.overlap:                    # same stream as the fronts: ordered after the slot's last reader
                fs = self.front_stream
                if not self.inputs_static:
                    fs.wait_stream(cur)         # x may have been produced by work queued on the caller's stream
                with torch.cuda.stream(fs):
                    slot['buf'].copy_(x, non_blocking=True)
                x.record_stream(fs)
            else:
                slot['buf'].copy_(x, non_blocking=True)
            slot['ready'] = None
        else:
            cs = self.copy_stream
            cs.wait_event(slot['done'])         # last reader of this slot (no-op before its first use)
            with torch.cuda.stream(cs):
                slot['buf'].copy_(x, non_blocking=True)
                slot['ready'] = torch.cuda.Event()
                slot['ready'].record(cs)
        return slot

    def _capture(self, kind, fn, dst):
        """Warm-up (weight cache, kernel attributes) + capture of `fn`, with the live contents of buffer set `dst`
        preserved.  Captures happen a handful of times per runner: the device is drained around them so that neither
        the warm-up nor the restore can race with a pass in flight on the other stream."""
        torch.cuda.synchronize(self.device)
        keep = [t.clone() for t in self._flat(self.sets[dst])]
        side = torch.cuda.Stream(device=self.device)
        # overlapping passes: the tile planner minimises SM time, not the makespan of a kernel alone on the GPU
        # RAMNET_DYNAMIC=1 | front | back (default off): work items drawn from a global counter (RAMNET_FLAG_DYNAMIC) so that
        # CTAs that start late -- their SM still ran the other stream's kernel -- take less of the work.  Validated
        # (bit-identical) but measured slower in this two-stream schedule: profiles/r02_dynamic_items.txt
        hint = ops.FLAG_SM_TIME if self.overlap else 0
        dyn = os.environ.get('RAMNET_DYNAMIC', '0')
        if self.overlap and (dyn == '1' or dyn == kind):
            hint |= ops.FLAG_DYNAMIC
        with ops.plan_flags(hint), torch.cuda.stream(side):
            for _ in range(2):
                fn()
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with ops.plan_flags(hint), torch.cuda.graph(graph, pool=self.pools[kind]):
            out = fn()
        if self.pools[kind] is None:
            self.pools[kind] = graph.pool()
        for t, k in zip(self._flat(self.sets[dst]), keep):
            t.copy_(k)
        torch.cuda.synchronize(self.device)
        return graph, out

    def _capture_front(self, which, xin, src, dst):
        net = self.net
        graph, _ = self._capture('front', lambda: net._encode(which, xin, self.sets[src], None, out_states=self.sets[dst]), dst)
        return graph, None

    def _capture_back(self, dst):
        net = self.net
        return self._capture('back', lambda: net.forward_decoder(self.sets[dst]), dst)
