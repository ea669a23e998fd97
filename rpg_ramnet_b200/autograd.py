"""torch.autograd.Function wrappers: the CUDA graph's fused ops with hand-written backward kernels.

The reference trains by calling loss.backward() through ATen/cuDNN autograd (trainer/lstm_trainer.py:450,
full BPTT, states never detached).  Here every fused forward op is one autograd node whose backward
issues our own kernels: pointwise adjoint of the fused epilogue (which also folds in the bias gradient) ->
weight gradient (tap-packed tcgen05 kernel) -> data gradient (the forward tcgen05 kernel on dZ with tap-flipped,
channel-transposed weights).  autograd only does the graph bookkeeping (BPTT order).

Parameter gradients do not travel through autograd's AccumulateGrad.  BPTT runs every layer's backward once per pass —
L * (K + 1) times per step — and round 1 paid per call for a zero-filled dW tensor, the split sum + scatter of the
weight-gradient partial tiles, a column-sum pass over dZ for the bias and an AccumulateGrad add.  Now:
  * bias gradients are accumulated by the pointwise adjoint itself (atomics) straight into `bias.grad`;
  * weight gradients accumulate as PARTIAL TILES in a per-layer workspace across the passes (`ops.WgradAccumulator`:
    the kernel's epilogue adds to what is there) and the split sum + scatter runs ONCE per step, into `weight.grad`,
    from a callback autograd runs when the backward pass ends (so `.grad` is complete when `loss.backward()`
    returns, whatever optimizer follows; with FusedAdam `.grad` is a view of its flat buffer).
Shapes the tap-packed kernel does not cover, FP32 mode, or RAMNET_WGRAD_DEFER=0 fall back to one full
`ramnet_conv_wgrad` per call, accumulated into `.grad` directly.
"""
import os

import torch

from . import ops


def _nhwc(g):
    return ops.as_nhwc(g)


# ------------------------------------------------------------------------------------------------------------------
# gradient sinks
# ------------------------------------------------------------------------------------------------------------------
def _grad_of(p):
    """`p.grad` as a dense fp32 tensor we may accumulate into in place (created on first use)."""
    if p.grad is None or p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
        g = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
        if p.grad is not None:
            g.copy_(p.grad)
        p.grad = g
    return p.grad


# Weight gradients feed nothing but the optimizer, so they run on a second stream beside the data-gradient chain that
# BPTT serialises (round 2; RAMNET_WGRAD_STREAM=0 keeps them on the backward stream).  Both are persistent full-GPU
# kernels: the gain is the SMs one leaves idle in its last wave, as for the inference passes (engine.GraphRunner).
_WGRAD_STREAMS = {}


def _wgrad_stream(device):
    if os.environ.get('RAMNET_WGRAD_STREAM', '1') == '0':
        return None
    s = _WGRAD_STREAMS.get(device)
    if s is None:
        s = _WGRAD_STREAMS[device] = torch.cuda.Stream(device=device)
    return s


class _Deferred:
    """Per-process registry of the weight-gradient accumulators that hold partial tiles of the running backward pass."""
    pending = []            # (accumulator, targets) with targets = [(parameter, row_begin, row_end)] in dW row order
    queued = False
    forked = []             # (backward stream, side stream) pairs to join when the backward pass ends
    streams = []            # streams backward nodes wrote parameter gradients on (the model's front stream, the caller's)

    @classmethod
    def _queue(cls):
        if not cls.queued:
            cls.queued = True
            torch.autograd.Variable._execution_engine.queue_callback(cls.flush)

    @classmethod
    def note(cls, acc, targets):
        if not any(a is acc for a, _ in cls.pending):
            cls.pending.append((acc, targets))
        cls._queue()

    @classmethod
    def note_stream(cls, cur):
        if not any(c == cur for c in cls.streams):
            cls.streams.append(cur)
        cls._queue()

    @classmethod
    def fork(cls, cur, side):
        """The side stream now holds work that `.grad` depends on: flush() joins it back into `cur`."""
        if not any(c == cur and s_ == side for c, s_ in cls.forked):
            cls.forked.append((cur, side))
        cls._queue()

    @classmethod
    def flush(cls):
        """End of the backward pass: one split sum + scatter per layer, accumulated into the parameters' .grad."""
        pending, cls.pending, cls.queued = cls.pending, [], False
        forked, cls.forked = cls.forked, []
        streams, cls.streams = cls.streams, []
        for cur, side in forked:
            cur.wait_stream(side)
        if streams:       # our nodes write .grad themselves (no AccumulateGrad), so the engine does not join their streams
            amb = torch.cuda.current_stream(streams[0].device)
            for cur in streams:
                if cur != amb:
                    amb.wait_stream(cur)
        with torch.no_grad():
            for acc, targets in pending:
                if len(targets) == 1:
                    acc.finalize(_grad_of(targets[0][0]))
                    continue
                rows = targets[-1][2]
                p0 = targets[0][0]
                tmp = torch.zeros((rows,) + tuple(p0.shape[1:]), dtype=torch.float32, device=p0.device)
                acc.finalize(tmp)
                for p, lo, hi in targets:          # a fused conv over several nn.Parameters (ConvGRU reset | update gates)
                    _grad_of(p).add_(tmp[lo:hi])


def _accumulator(weight, tag, builder):
    """The layer's WgradAccumulator, cached on the parameter object per (tag = shapes of the call)."""
    accs = getattr(weight, '_ramnet_wgrad_accs', None)
    if accs is None:
        accs = {}
        try:
            weight._ramnet_wgrad_accs = accs
        except AttributeError:
            return None, True
    if tag in accs:
        return accs[tag], False
    acc = builder()                    # runs the first partial launch itself (probe) or returns None
    accs[tag] = acc
    return acc, True


def _weight_grad(weights, dz, x0, x1, Cout, k, stride, kind, head=None):
    """dW of one fused conv.  `weights`: [(parameter, row_begin, row_end)].  Deferred when possible, else a full
    conv_wgrad accumulated into .grad now."""
    side = _wgrad_stream(dz.device)
    cur = torch.cuda.current_stream(dz.device)
    _Deferred.note_stream(cur)
    if side is not None:
        side.wait_stream(cur)                      # dz was produced on the backward stream just now
        with torch.cuda.stream(side):
            _weight_grad_on_stream(weights, dz, x0, x1, Cout, k, stride, kind, head)
        for t in (dz, x0, x1):                     # freed by the backward stream while the side stream may still read them
            if t is not None:
                t.record_stream(side)
        _Deferred.fork(cur, side)
        return
    _weight_grad_on_stream(weights, dz, x0, x1, Cout, k, stride, kind, head)


def _weight_grad_on_stream(weights, dz, x0, x1, Cout, k, stride, kind, head):
    w0 = weights[0][0]
    tag = (tuple(dz.shape), tuple(x0.shape), None if x1 is None else tuple(x1.shape), stride, kind, head)
    if head is None:
        acc, fresh = _accumulator(w0, tag, lambda: ops.wgrad_accumulator(x0, x1, dz, Cout, k, stride, kind))
    else:
        acc, fresh = _accumulator(w0, tag, lambda: ops.head_wgrad_accumulator(x0, dz, head))
    if acc is not None:
        if not fresh:
            acc.add(dz, x0, x1)
        _Deferred.note(acc, weights)
        return
    with torch.no_grad():
        if len(weights) == 1 and head is None:
            ops.conv_wgrad(dz, x0, x1, Cout, k, stride, _grad_of(w0), None, kind)
        elif head is not None:
            ops.head_conv_wgrad_tc(x0, dz, _grad_of(w0), None, head)
        else:
            tmp = torch.zeros((Cout,) + tuple(w0.shape[1:]), dtype=torch.float32, device=w0.device)
            ops.conv_wgrad(dz, x0, x1, Cout, k, stride, tmp, None, kind)
            for p, lo, hi in weights:
                _grad_of(p).add_(tmp[lo:hi])


def _bias_sink(bias, C):
    """Where the pointwise adjoint accumulates the bias gradient: bias.grad itself when the kernel can fold the column
    sums in (256 % (C/4) == 0), else None (the caller then sums dz separately)."""
    if bias is None:
        return None
    return _grad_of(bias) if ops.colsum_fusable(C) else None


def _bias_fallback(bias, dz):
    """Bias gradient for channel counts the fused column sum does not cover."""
    if bias is not None and not ops.colsum_fusable(dz.shape[1]):
        with torch.no_grad():
            _grad_of(bias).add_(dz.sum(dim=(0, 2, 3)))


class HeadConvFn(torch.autograd.Function):
    """ConvLayer head: relu(conv5x5(x_nchw) + b) -> NHWC."""

    @staticmethod
    def forward(ctx, x, weight, bias, round_tf32):
        Cout, Cin, k, _ = weight.shape
        b = None if bias is None else bias.detach()
        ctx.tc = bool(round_tf32) and k == 5 and ops.head_tc_ok(Cin, Cout)
        if ctx.tc:      # TF32 mode: tensor-core path over the horizontally unrolled input (kept for the weight gradient)
            x = ops.head_im2row(x)
            wp = _cached_pack(weight, 'head', lambda: ops.pack_weights_head(weight))
            y = ops.head_conv_tc(x, wp, b, Cin, Cout, True)
        else:
            y = ops.head_conv(x, weight.detach(), b, round_tf32)
        ctx.save_for_backward(x, y)
        ctx.weight, ctx.bias = weight, bias
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        dz = ops.relu_bwd(_nhwc(dy), y, round_tf32=ctx.tc, db=_bias_sink(bias, y.shape[1]))
        _bias_fallback(bias, dz)
        if ctx.tc:
            _weight_grad([(weight, 0, weight.shape[0])], dz, x, None, weight.shape[0], 5, 1, ops.MMA_TF32, head=weight.shape[1])
        else:
            with torch.no_grad():
                ops.head_conv_wgrad(x, dz, _grad_of(weight), None)
        return None, None, None, None


# Data-gradient weight packs live on the parameter object itself and are rebuilt when it changes (same token as
# engine.WeightCache: in-place version counter + the epoch the fused Adam bumps).  BPTT calls every layer's backward
# L*(K+1) times per step with the same weights, so all but the first call reuse the pack.  (Keyed on the object, not
# on data_ptr: a new tensor that recycles a freed address must never see a stale pack.)
def _cached_pack(weight, tag, builder, also=()):
    from . import engine
    tok = (weight.data_ptr(), weight._version, engine._WEIGHT_EPOCH) + tuple((w.data_ptr(), w._version) for w in also)
    packs = getattr(weight, '_ramnet_dgrad_packs', None)
    if packs is None:
        packs = {}
        try:
            weight._ramnet_dgrad_packs = packs
        except AttributeError:
            return builder()
    hit = packs.get(tag)
    if hit is not None and hit[0] == tok:
        return hit[1]
    packed = builder()
    packs[tag] = (tok, packed)
    return packed


def _dgrad(dz, weight, kind, stride, ci_begin, ci_count, in_hw, add=None, key_weight=None, also=()):
    """Data gradient w.r.t. input channels [ci_begin, ci_begin+ci_count) of a conv with nn.Conv2d weight `weight`.
    `add`: a gradient already produced for the same tensor; the contribution is accumulated onto it IN PLACE by the
    kernel's epilogue (EPI_BIAS_ADD) instead of a separate elementwise add.  `key_weight` / `also`: the parameter
    object(s) the pack cache hangs on when `weight` is a temporary (the concatenated ConvGRU reset | update weight)."""
    Cout, _, k, _ = weight.shape
    H, W = int(in_hw[0]), int(in_hw[1])
    kw = weight if key_weight is None else key_weight
    if (stride == 2 and kind == ops.MMA_TF32 and k in (3, 5) and H % 2 == 0 and W % 2 == 0 and Cout % 32 == 0
            and ci_count % 32 == 0):
        # sub-pixel decomposition: four stride-1 convolutions of dZ, one per input parity (no zero insertion)
        wp = _cached_pack(kw, ('s2', ci_begin, ci_count), lambda: ops.pack_weights_dgrad_s2(weight, ci_begin, ci_count), also)
        dx = ops.conv_dgrad_s2(dz, wp, ci_count, k, H, W)
        return dx if add is None else add.add_(dx)
    wp = _cached_pack(kw, (kind, ci_begin, ci_count), lambda: ops.pack_weights_dgrad(weight, kind, ci_begin, ci_count), also)
    if stride == 2:
        dz = ops.zero_insert2x(dz, H, W)
    if add is not None and kind == ops.MMA_TF32:
        return ops.conv_fwd(dz, None, wp, None, ci_count, k, 1, ops.EPI_BIAS_ADD, kind, aux0=add, out0=add)
    dx = ops.conv_fwd(dz, None, wp, None, ci_count, k, 1, ops.EPI_BIAS, kind)
    return dx if add is None else add.add_(dx)


class ConvFn(torch.autograd.Function):
    """conv(+bias)(+residual)(+relu) over the virtual concat [x0 | x1]."""

    @staticmethod
    def forward(ctx, x0, x1, res, weight, bias, packed_w, epilogue, kind, stride, round_out):
        y = ops.conv_fwd(x0, x1, packed_w, None if bias is None else bias.detach(), weight.shape[0], weight.shape[2],
                         stride, epilogue, kind, aux0=res, round_tf32=round_out)
        ctx.save_for_backward(x0, x1, y)
        ctx.weight, ctx.bias = weight, bias
        ctx.cfg = (epilogue, kind, stride, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x0, x1, y = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        epilogue, kind, stride, has_res = ctx.cfg
        dy = _nhwc(dy)
        Cout, Ct, k, _ = weight.shape
        if epilogue == ops.EPI_BIAS:
            dz = dy
            if bias is not None:
                with torch.no_grad():
                    _grad_of(bias).add_(dz.sum(dim=(0, 2, 3)))
        else:
            dz = ops.relu_bwd(dy, y, round_tf32=(kind == ops.MMA_TF32), db=_bias_sink(bias, Cout))
            _bias_fallback(bias, dz)
        _weight_grad([(weight, 0, Cout)], dz, x0, x1, Cout, k, stride, kind)
        C0 = x0.shape[1]
        dx0 = _dgrad(dz, weight, kind, stride, 0, C0, x0.shape[2:]) if ctx.needs_input_grad[0] else None
        dx1 = None
        if x1 is not None and ctx.needs_input_grad[1]:
            dx1 = _dgrad(dz, weight, kind, stride, C0, Ct - C0, x1.shape[2:])
        dres = dz if (has_res and ctx.needs_input_grad[2]) else None
        if dres is not None and _wgrad_stream(dz.device) is not None:
            dres = dz.clone()      # autograd may accumulate into a returned gradient in place; the side stream still reads dz
        return dx0, dx1, dres, None, None, None, None, None, None, None


class GruFn(torch.autograd.Function):
    """ConvGRU.forward (submodules.py:436-454) as two fused convolutions; backward = the same two GEMMs
    transposed plus two pointwise gate adjoints."""

    @staticmethod
    def forward(ctx, x, h, w_r, b_r, w_u, b_u, w_o, b_o, ru_pack, out_pack, kind):
        N, C, H, W = x.shape
        r = ops.empty_nhwc(N, C, H, W, x.device)
        o = ops.empty_nhwc(N, C, H, W, x.device)
        tf32 = kind == ops.MMA_TF32
        u, rh = ops.conv_fwd(x, h, ru_pack.w, ru_pack.b, 2 * C, 3, 1, ops.EPI_GRU_RU, kind, aux0=h, round_tf32=tf32, stash=r)
        hn = ops.conv_fwd(x, rh, out_pack.w, out_pack.b, C, 3, 1, ops.EPI_GRU_OUT, kind, aux0=h, aux1=u, round_tf32=tf32,
                          stash=o)
        ctx.save_for_backward(x, h, u, r, rh, o)
        ctx.params = (w_r, b_r, w_u, b_u, w_o, b_o)
        ctx.kind = kind
        return hn

    @staticmethod
    def backward(ctx, dhn):
        x, h, u, r, rh, o = ctx.saved_tensors
        w_r, b_r, w_u, b_u, w_o, b_o = ctx.params
        kind = ctx.kind
        N, C, H, W = x.shape
        tf32 = kind == ops.MMA_TF32
        fuse = ops.colsum_fusable(C)
        # bias gradients of the fused [reset | update] conv land in one [2C] scratch (two parameters own its halves)
        db_ru = torch.zeros(2 * C, dtype=torch.float32, device=x.device) if (fuse and (b_r is not None or b_u is not None)) else None
        dzo, dzru, dh = ops.gru_out_bwd(_nhwc(dhn), h, u, o, round_tf32=tf32,
                                        db_o=_bias_sink(b_o, C) if fuse else None, db_ru=db_ru)
        _weight_grad([(w_o, 0, C)], dzo, x, rh, C, 3, 1, kind)
        dx = _dgrad(dzo, w_o, kind, 1, 0, C, (H, W))
        drh = _dgrad(dzo, w_o, kind, 1, C, C, (H, W))
        ops.gru_ru_bwd(drh, h, r, dzru, dh, round_tf32=tf32, db_ru=db_ru)
        with torch.no_grad():
            if db_ru is not None:
                if b_r is not None:
                    _grad_of(b_r).add_(db_ru[:C])
                if b_u is not None:
                    _grad_of(b_u).add_(db_ru[C:])
            elif not fuse:
                if b_o is not None:
                    _grad_of(b_o).add_(dzo.sum(dim=(0, 2, 3)))
                s = dzru.sum(dim=(0, 2, 3))
                if b_r is not None:
                    _grad_of(b_r).add_(s[:C])
                if b_u is not None:
                    _grad_of(b_u).add_(s[C:])
        _weight_grad([(w_r, 0, C), (w_u, C, 2 * C)], dzru, x, h, 2 * C, 3, 1, kind)
        # the concatenated [reset | update] weight only exists to be re-packed for the data gradient: build it when a
        # pack is actually missing (once per optimizer step), not on every backward call
        w_ru = _LazyCat(w_r, w_u)
        dx = _dgrad(dzru, w_ru, kind, 1, 0, C, (H, W), add=dx, key_weight=w_r, also=(w_u,))
        dh = _dgrad(dzru, w_ru, kind, 1, C, C, (H, W), add=dh, key_weight=w_r, also=(w_u,))
        return (dx if ctx.needs_input_grad[0] else None, dh if ctx.needs_input_grad[1] else None,
                None, None, None, None, None, None, None, None, None)


class _LazyCat:
    """torch.cat([w_r, w_u]) materialised only if a packer reads it (`.detach()` / `.shape`)."""

    def __init__(self, a, b):
        self.a, self.b = a, b
        self.shape = (a.shape[0] + b.shape[0],) + tuple(a.shape[1:])

    def detach(self):
        return torch.cat([self.a.detach(), self.b.detach()], 0)


class LstmFn(torch.autograd.Function):
    """ConvLSTM.forward (submodules.py:318-358) as one fused convolution with gate-interleaved columns; backward =
    one pointwise gate adjoint + the transposed GEMMs on the unpermuted Gates.weight."""

    @staticmethod
    def forward(ctx, x, h, c, weight, bias, pack, kind):
        N, _, H, W = x.shape
        C = weight.shape[0] // 4
        gates = torch.empty((N, H, W, C, 4), dtype=torch.float32, device=x.device)
        hn, cn = ops.conv_fwd(x, h, pack.w, pack.b, 4 * C, 3, 1, ops.EPI_LSTM, kind, aux0=c,
                              round_tf32=(kind == ops.MMA_TF32), stash=gates)
        ctx.save_for_backward(x, h, c, cn, gates)
        ctx.weight, ctx.bias = weight, bias
        ctx.kind = kind
        ctx.mark_non_differentiable()
        return hn, cn

    @staticmethod
    def backward(ctx, dhn, dcn):
        x, h, c, cn, gates = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        kind = ctx.kind
        N, Cx, H, W = x.shape
        C = weight.shape[0] // 4
        dz, dc = ops.lstm_bwd(None if dhn is None else _nhwc(dhn), None if dcn is None else _nhwc(dcn), gates, c, cn,
                              round_tf32=(kind == ops.MMA_TF32), db=None if bias is None else _grad_of(bias))
        _weight_grad([(weight, 0, 4 * C)], dz, x, h, 4 * C, 3, 1, kind)
        dx = _dgrad(dz, weight, kind, 1, 0, Cx, (H, W)) if ctx.needs_input_grad[0] else None
        dh = _dgrad(dz, weight, kind, 1, Cx, C, (H, W)) if ctx.needs_input_grad[1] else None
        return dx, dh, (dc if ctx.needs_input_grad[2] else None), None, None, None, None


class TransposedConvFn(torch.autograd.Function):
    """TransposedConvLayer.forward (submodules.py:38-66, norm-free): relu(ConvTranspose2d(k, stride 2, padding k//2,
    output_padding 1)(x + skip) + b).  A transposed convolution is the data gradient of the stride-2 convolution that
    shares its weight tensor W [Cin, Cout, k, k] (nn.Conv2d layout of a Cout -> Cin conv), so
      forward  = zero insertion + the stride-1 kernel on tap-flipped, channel-transposed weights,
      d input  = that stride-2 convolution applied to dZ (the forward kernel, plain packing of W),
      d weight = the stride-2 conv's weight gradient with the roles swapped: "dZ" operand = x + skip, "X" operand = dZ."""

    @staticmethod
    def forward(ctx, x, skip, weight, bias, packed_w, kind, relu=True):
        Cin, Cout, k, _ = weight.shape
        N, _, H, W = x.shape
        up, xs = ops.zero_insert2x(x, 2 * H, 2 * W, skip=skip, want_sum=True)
        y = ops.conv_fwd(up, None, packed_w, None if bias is None else bias.detach(), Cout, k, 1,
                         ops.EPI_BIAS_RELU if relu else ops.EPI_BIAS, kind)
        ctx.save_for_backward(xs, y)
        ctx.weight, ctx.bias, ctx.kind, ctx.has_skip, ctx.relu = weight, bias, kind, skip is not None, relu
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, y = ctx.saved_tensors
        weight, bias, kind = ctx.weight, ctx.bias, ctx.kind
        Cin, Cout, k, _ = weight.shape
        if ctx.relu:
            dz = ops.relu_bwd(_nhwc(dy), y, round_tf32=(kind == ops.MMA_TF32), db=_bias_sink(bias, Cout))
            _bias_fallback(bias, dz)
        else:                               # a norm layer follows (its adjoint has already rounded dz)
            dz = _nhwc(dy)
            if bias is not None:
                with torch.no_grad():
                    _grad_of(bias).add_(dz.sum(dim=(0, 2, 3)))
        _weight_grad([(weight, 0, Cin)], xs, dz, None, Cin, k, 2, kind)
        dx = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            wp = _cached_pack(weight, ('tconv_dx', kind), lambda: ops.pack_weights(weight, kind))
            dx = ops.conv_fwd(dz, None, wp, None, Cin, k, 2, ops.EPI_BIAS, kind)
        return (dx if ctx.needs_input_grad[0] else None, dx if (ctx.has_skip and ctx.needs_input_grad[1]) else None,
                None, None, None, None, None)


class UpsampleAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, skip, round_tf32):
        ctx.has_skip = skip is not None
        return ops.upsample2x_add(x, skip, round_tf32)

    @staticmethod
    def backward(ctx, dy):
        g = ops.upsample2x_bwd(_nhwc(dy))
        return g, (g if ctx.has_skip else None), None


class PredFn(torch.autograd.Function):
    """1x1 pred conv + sigmoid over x (+ skip), or over cat([x, skip]) when `concat` (unet.py:11-12,129)."""

    @staticmethod
    def forward(ctx, x, skip, weight, bias, concat=False):
        depth = ops.pred_sigmoid(x, skip, weight.detach(), None if bias is None else bias.detach(), concat=concat)
        ctx.save_for_backward(x, skip, depth)
        ctx.weight, ctx.bias, ctx.concat = weight, bias, concat
        return depth

    @staticmethod
    def backward(ctx, ddepth):
        x, skip, depth = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        dx, dskip, dw, db = ops.pred_bwd(ddepth, depth, x, weight, skip, ctx.concat)
        with torch.no_grad():
            _grad_of(weight).add_(dw.view(weight.shape))
            if bias is not None:
                _grad_of(bias).add_(db)
        return dx, dskip, None, None, None


class NormActFn(torch.autograd.Function):
    """act(norm(z) (+ res)) for a live normalisation layer (train-mode BatchNorm2d / InstanceNorm2d, the
    ResidualBlock's InstanceNorm2d without running statistics, or an eval-mode norm that gradients flow through):
    submodules.py:29-33, 60-64, 91-95, 203-214.  z is the bias-only output of the conv before it (ConvFn with
    EPI_BIAS keeps it for its own backward; no second copy).  Running statistics are updated in place by the forward
    kernel; dgamma / dbeta accumulate straight into the parameters' .grad."""

    @staticmethod
    def forward(ctx, z, res, gamma, beta, norm_mod, kind, act, batch_stats, round_out):
        g = None if gamma is None else gamma.detach().float().contiguous()
        b = None if beta is None else beta.detach().float().contiguous()
        rm, rv = getattr(norm_mod, 'running_mean', None), getattr(norm_mod, 'running_var', None)
        y, stats = ops.norm_fwd(z, kind, act, g, b, res, rm, rv, norm_mod.momentum if norm_mod.momentum is not None else 0.1,
                                norm_mod.eps, batch_stats, round_out)
        if rm is not None and batch_stats and kind == 'BN' and getattr(norm_mod, 'num_batches_tracked', None) is not None:
            norm_mod.num_batches_tracked.add_(1)
        ctx.save_for_backward(z, y, stats)
        ctx.gamma, ctx.beta = gamma, beta
        ctx.cfg = (kind, act, batch_stats, round_out, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, y, stats = ctx.saved_tensors
        kind, act, batch_stats, round_out, has_res = ctx.cfg
        gamma, beta = ctx.gamma, ctx.beta
        want_dres = has_res and ctx.needs_input_grad[1]
        _Deferred.note_stream(torch.cuda.current_stream(z.device))
        dz, dres = ops.norm_bwd(_nhwc(dy), y, z, stats, kind, act, None if gamma is None else gamma.detach().float().contiguous(),
                                batch_stats, round_out, want_dres,
                                None if gamma is None else _grad_of(gamma), None if beta is None else _grad_of(beta))
        return dz, dres, None, None, None, None, None, None, None


class PredLogitsFn(torch.autograd.Function):
    """1x1 pred conv without its activation (a norm layer sits between them, statenet.py:116-117 with norm set)."""

    @staticmethod
    def forward(ctx, x, skip, weight, bias, concat=False):
        logits = ops.pred_logits(x, skip, weight.detach(), None if bias is None else bias.detach(), concat)
        ctx.save_for_backward(x, skip)
        ctx.weight, ctx.bias, ctx.concat = weight, bias, concat
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        x, skip = ctx.saved_tensors
        weight, bias = ctx.weight, ctx.bias
        dx, dskip, dw, db = ops.pred_logits_bwd(dlogits, x, weight, skip, ctx.concat)
        with torch.no_grad():
            _grad_of(weight).add_(dw.view(weight.shape))
            if bias is not None:
                _grad_of(bias).add_(db)
        return dx, dskip, None, None, None
