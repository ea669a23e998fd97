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
