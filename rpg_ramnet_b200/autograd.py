"""torch.autograd.Function wrappers: the CUDA graph's fused ops with hand-written backward kernels.

The reference trains by calling loss.backward() through ATen/cuDNN autograd (trainer/lstm_trainer.py:450,
full BPTT, states never detached).  Here every fused forward op is one autograd node whose backward
issues our own kernels: pointwise adjoint of the fused epilogue -> weight/bias gradient
(ramnet_conv_wgrad, accumulated in nn.Conv2d layout) -> data gradient (the forward tcgen05 kernel on dZ
with tap-flipped, channel-transposed weights).  autograd only does the graph bookkeeping (BPTT order,
gradient accumulation into .grad), so torch.optim / our fused Adam see ordinary .grad tensors.
"""
import torch

from . import ops


def _nhwc(g):
    return ops.as_nhwc(g)


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
        ctx.has_bias = bias is not None
        ctx.wshape = weight.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dz = ops.relu_bwd(_nhwc(dy), y, round_tf32=ctx.tc)
        dw = torch.zeros(ctx.wshape, dtype=torch.float32, device=y.device)
        db = torch.zeros(ctx.wshape[0], dtype=torch.float32, device=y.device) if ctx.has_bias else None
        if ctx.tc:
            ops.head_conv_wgrad_tc(x, dz, dw, db, ctx.wshape[1])
        else:
            ops.head_conv_wgrad(x, dz, dw, db)
        return None, dw, db, None


# Data-gradient weight packs live on the parameter object itself and are rebuilt when it changes (same token as
# engine.WeightCache: in-place version counter + the epoch the fused Adam bumps).  BPTT calls every layer's backward
# L*(K+1) times per step with the same weights, so all but the first call reuse the pack.  (Keyed on the object, not
# on data_ptr: a new tensor that recycles a freed address must never see a stale pack.)
def _cached_pack(weight, tag, builder):
    from . import engine
    tok = (weight.data_ptr(), weight._version, engine._WEIGHT_EPOCH)
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


def _dgrad(dz, weight, kind, stride, ci_begin, ci_count, in_hw):
    """Data gradient w.r.t. input channels [ci_begin, ci_begin+ci_count) of a conv with nn.Conv2d weight `weight`."""
    Cout, _, k, _ = weight.shape
    H, W = int(in_hw[0]), int(in_hw[1])
    if (stride == 2 and kind == ops.MMA_TF32 and k in (3, 5) and H % 2 == 0 and W % 2 == 0 and Cout % 32 == 0
            and ci_count % 32 == 0):
        # sub-pixel decomposition: four stride-1 convolutions of dZ, one per input parity (no zero insertion)
        wp = _cached_pack(weight, ('s2', ci_begin, ci_count), lambda: ops.pack_weights_dgrad_s2(weight, ci_begin, ci_count))
        return ops.conv_dgrad_s2(dz, wp, ci_count, k, H, W)
    wp = _cached_pack(weight, (kind, ci_begin, ci_count), lambda: ops.pack_weights_dgrad(weight, kind, ci_begin, ci_count))
    if stride == 2:
        dz = ops.zero_insert2x(dz, H, W)
    return ops.conv_fwd(dz, None, wp, None, ci_count, k, 1, ops.EPI_BIAS, kind)


class ConvFn(torch.autograd.Function):
    """conv(+bias)(+residual)(+relu) over the virtual concat [x0 | x1]."""

    @staticmethod
    def forward(ctx, x0, x1, res, weight, bias, packed_w, epilogue, kind, stride, round_out):
        y = ops.conv_fwd(x0, x1, packed_w, None if bias is None else bias.detach(), weight.shape[0], weight.shape[2],
                         stride, epilogue, kind, aux0=res, round_tf32=round_out)
        ctx.save_for_backward(x0, x1, y, weight)
        ctx.cfg = (epilogue, kind, stride, bias is not None, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x0, x1, y, weight = ctx.saved_tensors
        epilogue, kind, stride, has_bias, has_res = ctx.cfg
        dy = _nhwc(dy)
        dz = dy if epilogue == ops.EPI_BIAS else ops.relu_bwd(dy, y, round_tf32=(kind == ops.MMA_TF32))
        Cout, Ct, k, _ = weight.shape
        dw = torch.zeros_like(weight, dtype=torch.float32)
        db = torch.zeros(Cout, dtype=torch.float32, device=y.device) if has_bias else None
        ops.conv_wgrad(dz, x0, x1, Cout, k, stride, dw, db, kind)
        C0 = x0.shape[1]
        dx0 = _dgrad(dz, weight, kind, stride, 0, C0, x0.shape[2:]) if ctx.needs_input_grad[0] else None
        dx1 = None
        if x1 is not None and ctx.needs_input_grad[1]:
            dx1 = _dgrad(dz, weight, kind, stride, C0, Ct - C0, x1.shape[2:])
        dres = dz if (has_res and ctx.needs_input_grad[2]) else None
        return dx0, dx1, dres, dw, db, None, None, None, None, None


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
        ctx.save_for_backward(x, h, u, r, rh, o, w_r, w_u, w_o)
        ctx.kind = kind
        return hn

    @staticmethod
    def backward(ctx, dhn):
        x, h, u, r, rh, o, w_r, w_u, w_o = ctx.saved_tensors
        kind = ctx.kind
        N, C, H, W = x.shape
        tf32 = kind == ops.MMA_TF32
        dzo, dzru, dh = ops.gru_out_bwd(_nhwc(dhn), h, u, o, round_tf32=tf32)
        dw_o = torch.zeros_like(w_o, dtype=torch.float32)
        db_o = torch.zeros(C, dtype=torch.float32, device=x.device)
        ops.conv_wgrad(dzo, x, rh, C, 3, 1, dw_o, db_o, kind)
        dx = _dgrad(dzo, w_o, kind, 1, 0, C, (H, W))
        drh = _dgrad(dzo, w_o, kind, 1, C, C, (H, W))
        ops.gru_ru_bwd(drh, h, r, dzru, dh, round_tf32=tf32)
        w_ru = torch.cat([w_r.detach(), w_u.detach()], 0)
        dw_ru = torch.zeros_like(w_ru, dtype=torch.float32)
        db_ru = torch.zeros(2 * C, dtype=torch.float32, device=x.device)
        ops.conv_wgrad(dzru, x, h, 2 * C, 3, 1, dw_ru, db_ru, kind)
        dx = dx + _dgrad(dzru, w_ru, kind, 1, 0, C, (H, W))
        dh = dh + _dgrad(dzru, w_ru, kind, 1, C, C, (H, W))
        return (dx if ctx.needs_input_grad[0] else None, dh if ctx.needs_input_grad[1] else None,
                dw_ru[:C].contiguous(), db_ru[:C].contiguous(), dw_ru[C:].contiguous(), db_ru[C:].contiguous(),
                dw_o, db_o, None, None, None)


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
        ctx.save_for_backward(x, h, c, cn, gates, weight)
        ctx.kind = kind
        ctx.mark_non_differentiable()
        return hn, cn

    @staticmethod
    def backward(ctx, dhn, dcn):
        x, h, c, cn, gates, weight = ctx.saved_tensors
        kind = ctx.kind
        N, Cx, H, W = x.shape
        C = weight.shape[0] // 4
        dz, dc = ops.lstm_bwd(None if dhn is None else _nhwc(dhn), None if dcn is None else _nhwc(dcn), gates, c, cn,
                              round_tf32=(kind == ops.MMA_TF32))
        dw = torch.zeros_like(weight, dtype=torch.float32)
        db = torch.zeros(4 * C, dtype=torch.float32, device=x.device)
        ops.conv_wgrad(dz, x, h, 4 * C, 3, 1, dw, db, kind)
        dx = _dgrad(dz, weight, kind, 1, 0, Cx, (H, W)) if ctx.needs_input_grad[0] else None
        dh = _dgrad(dz, weight, kind, 1, Cx, C, (H, W)) if ctx.needs_input_grad[1] else None
        return dx, dh, (dc if ctx.needs_input_grad[2] else None), dw, db, None, None


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
    """1x1 pred conv + sigmoid."""

    @staticmethod
    def forward(ctx, x, skip, weight, bias):
        depth = ops.pred_sigmoid(x, skip, weight.detach(), None if bias is None else bias.detach())
        ctx.save_for_backward(x, skip, depth, weight)
        ctx.has_bias = bias is not None
        return depth

    @staticmethod
    def backward(ctx, ddepth):
        x, skip, depth, weight = ctx.saved_tensors
        dx, dw, db = ops.pred_bwd(ddepth, depth, x, weight, skip)
        return dx, (dx if skip is not None else None), dw.view(weight.shape), (db if ctx.has_bias else None)
