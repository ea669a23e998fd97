"""scale_invariant_loss on the device without dynamic shapes (mirrors RAM_Net/model/loss.py:6-9).

Same call signature.  Forward = one reduction kernel (sum d, sum d^2, count of non-NaN) + a
1-thread finaliser; backward = one streaming kernel writing (2w/n)(d - lambda*mean(d)) (0 at NaN).
No boolean-mask gather, no host sync.
"""
import torch

from .. import ops


class _ScaleInvariantLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_input, y_target, weight, n_lambda, process_group):
        pred, target = y_input.detach().float().contiguous(), y_target.detach().float().contiguous()
        stats = ops.si_loss_stats(pred, target)
        if process_group is not False:          # exact global-batch loss under data parallelism
            from ..distributed import all_reduce_loss_stats
            all_reduce_loss_stats(stats, None if process_group is True else process_group)
        ctx.save_for_backward(pred, target, stats)
        ctx.weight, ctx.n_lambda = float(weight), float(n_lambda)
        return ops.si_loss_value(stats, float(weight), float(n_lambda))

    @staticmethod
    def backward(ctx, grad_out):
        pred, target, stats = ctx.saved_tensors
        g = ops.si_loss_grad(pred, target, stats, ctx.weight, ctx.n_lambda, 1.0)
        return g * grad_out, None, None, None, None


def scale_invariant_loss(y_input, y_target, weight=1.0, n_lambda=1.0, process_group=False):
    """`process_group`: False (default) = local statistics, as the single-process reference; True or a
    torch.distributed group = all-reduce (sum d, sum d^2, n) first, so every rank gets the loss and
    gradient of the GLOBAL batch (gradients are then summed across ranks, not averaged)."""
    return _ScaleInvariantLoss.apply(y_input, y_target, weight, n_lambda, process_group)


class _MultiScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, target, start_scale, num_scales):
        pred, tgt = prediction.detach().float().contiguous(), target.detach().float().contiguous()
        stats = ops.msg_loss_stats(pred, tgt, start_scale, num_scales)
        ctx.save_for_backward(pred, tgt, stats)
        ctx.cfg = (start_scale, num_scales)
        return ops.msg_loss_value(stats, pred.shape[0], num_scales)

    @staticmethod
    def backward(ctx, grad_out):
        pred, tgt, stats = ctx.saved_tensors
        g = ops.msg_loss_grad(pred, tgt, stats, ctx.cfg[0], ctx.cfg[1], 1.0)
        return g * grad_out, None, None, None


class MultiScaleGradient(torch.nn.Module):
    """Mirrors model/loss.py:22-70 (`MultiScaleGradient(start_scale=1, num_scales=4)`), device-side, no dynamic shapes:
    per scale one reduction kernel (sum |Sobel|, count of non-NaN) and one gradient kernel.  `preview=True` (TensorBoard
    visualisation, lstm_trainer.py:165) is host-side plotting and is not part of the path."""

    def __init__(self, start_scale=1, num_scales=4):
        super().__init__()
        self.start_scale, self.num_scales = start_scale, num_scales

    def forward(self, prediction, target, preview=False):
        if preview:
            raise NotImplementedError('preview=True is a TensorBoard visualisation aid of the reference trainer')
        return _MultiScaleGrad.apply(prediction, target, self.start_scale, self.num_scales)


multi_scale_grad_loss_fn = MultiScaleGradient()


def multi_scale_grad_loss(prediction, target, preview=False):
    return multi_scale_grad_loss_fn.forward(prediction, target, preview)
