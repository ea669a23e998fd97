"""Losses of the RAM-Net trainer on the device, without dynamic shapes (mirrors RAM_Net/model/loss.py, every public
name: `scale_invariant_loss` :6-9, `scale_invariant_log_loss` :12-15, `mse_loss` :18-19, `MultiScaleGradient` :22-63,
`multi_scale_grad_loss_fn` :66, `multi_scale_grad_loss` :69-70 — `trainer/lstm_trainer.py:5` imports the last two
names and `mse_loss`).

Same call signatures.  Forward = one reduction kernel per term (sum d, sum d^2, count of non-NaN d) + a 1-thread
finaliser; backward = one streaming kernel writing (2w/n)(d - lambda*mean(d)) (0 at NaN) already multiplied by
autograd's grad_output (read on the device).  No boolean-mask gather, no host sync.

Data parallelism: the reference normalises by the count of valid pixels of the WHOLE batch and subtracts
the whole batch's mean, so per-rank losses do not average to the single-process loss.  `process_group`:
  None  (default) -> global statistics whenever torch.distributed is initialised with world_size > 1 (the flat
                     gradient all-reduce of FusedAdam SUMS, which is only right for global-statistics gradients),
                     local statistics otherwise (= the single-process reference);
  False           -> local statistics always;   True / a group -> all-reduce over the default / that group.
`SILossBatch` gathers the statistics of all L x keys terms of a sequence in one [T,3] buffer so that a training step
needs ONE all-reduce of 3T doubles between forward and backward instead of one per term.
"""
import torch
import torch.nn.functional as F

from .. import ops


def _resolve_group(process_group):
    """-> (exchange: bool, group or None)."""
    import torch.distributed as dist
    if process_group is False:
        return False, None
    group = None if process_group in (None, True) else process_group
    if not (dist.is_available() and dist.is_initialized()):
        return False, None
    return dist.get_world_size(group) > 1, group


class _ScaleInvariantLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_input, y_target, weight, n_lambda, process_group, log_space):
        pred, target = y_input.detach().float().contiguous(), y_target.detach().float().contiguous()
        stats = ops.si_loss_stats(pred, target, log_space=log_space)
        exchange, group = _resolve_group(process_group)
        if exchange:                            # exact global-batch loss under data parallelism
            from ..distributed import all_reduce_loss_stats
            all_reduce_loss_stats(stats, group)
        ctx.save_for_backward(pred, target, stats)
        ctx.cfg = (float(weight), float(n_lambda), bool(log_space))
        return ops.si_loss_value(stats, float(weight), float(n_lambda))

    @staticmethod
    def backward(ctx, grad_out):
        pred, target, stats = ctx.saved_tensors
        w, lam, log_space = ctx.cfg
        g = ops.si_loss_grad(pred, target, stats, w, lam, 1.0, scale_dev=grad_out, log_space=log_space)
        return g, None, None, None, None, None


def scale_invariant_loss(y_input, y_target, weight=1.0, n_lambda=1.0, process_group=None):
    """model/loss.py:6-9: w * (mean(d^2) - lambda * mean(d)^2) over the non-NaN d = y_input - y_target."""
    return _ScaleInvariantLoss.apply(y_input, y_target, weight, n_lambda, process_group, False)


def scale_invariant_log_loss(y_input, y_target, n_lambda=1.0, process_group=None):
    """model/loss.py:12-15: the same statistic on d = log(y_input) - log(y_target) (metric-depth inputs)."""
    return _ScaleInvariantLoss.apply(y_input, y_target, 1.0, n_lambda, process_group, True)


def mse_loss(y_input, y_target, process_group=None):
    """model/loss.py:18-19: F.mse_loss over the pixels whose target is not NaN = the scale-invariant statistic with
    lambda = 0.  (The reference masks on isnan(target) only; predictions are sigmoid outputs, hence finite, so the
    mask on isnan(input - target) used here selects the same pixels.)"""
    return _ScaleInvariantLoss.apply(y_input, y_target, 1.0, 0.0, process_group, False)


class _SIBatchFn(torch.autograd.Function):
    """All terms of a SILossBatch as one autograd node: [T] loss values out, one gradient kernel per term back."""

    @staticmethod
    def forward(ctx, batch, *preds):
        ctx.batch = batch
        T = len(batch.terms)
        exchange, group = _resolve_group(batch.process_group)
        stats = batch.stats[:T]
        if exchange:                            # ONE all-reduce of 3T doubles for the whole sequence
            import torch.distributed as dist
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
        out = torch.empty(T, dtype=torch.float32, device=stats.device)
        for i, (_, _, w, lam) in enumerate(batch.terms):
            ops.si_loss_value(stats[i], w, lam, out=out[i:i + 1])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        batch = ctx.batch
        grad_out = grad_out.contiguous().float()
        grads = []
        for i, (pred, target, w, lam) in enumerate(batch.terms):
            grads.append(ops.si_loss_grad(pred, target, batch.stats[i], w, lam, 1.0, scale_dev=grad_out[i:i + 1]))
        return (None, *grads)


class SILossBatch:
    """Scale-invariant loss terms of one training sequence with a single statistics exchange.

        batch = SILossBatch(capacity=L * len(keys), device=dev)
        for item in sequence: ... batch.add(pred, target, weight, n_lambda)        # statistics kernel, no sync
        terms = batch.finish()        # [T] float32 losses (ONE all-reduce under data parallelism), autograd-connected

    Equivalent to T calls of scale_invariant_loss (tests/test_dist_gloo.py, tests/test_gpu_train.py)."""

    def __init__(self, capacity, device, process_group=None):
        self.stats = torch.zeros((capacity, 3), dtype=torch.float64, device=device)
        self.process_group = process_group
        self.terms, self.inputs = [], []

    def add(self, y_input, y_target, weight=1.0, n_lambda=1.0):
        i = len(self.terms)
        if i >= self.stats.shape[0]:
            raise ops._lib.RamnetError(f'SILossBatch: capacity {self.stats.shape[0]} exceeded')
        pred, target = y_input.detach().float().contiguous(), y_target.detach().float().contiguous()
        ops.si_loss_stats(pred, target, out=self.stats[i])
        self.terms.append((pred, target, float(weight), float(n_lambda)))
        self.inputs.append(y_input)
        return i

    def finish(self):
        out = _SIBatchFn.apply(self, *self.inputs)
        self.inputs = []
        return out


class _MultiScaleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, target, start_scale, num_scales, process_group):
        pred, tgt = prediction.detach().float().contiguous(), target.detach().float().contiguous()
        stats, signs = ops.msg_loss_stats(pred, tgt, start_scale, num_scales, want_signs=True)
        n_batch = pred.shape[0]
        exchange, group = _resolve_group(process_group)
        if exchange:        # global (sum |g|, count) per scale and the global batch size (loss.py:57 multiplies by B)
            import torch.distributed as dist
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
            n_batch *= dist.get_world_size(group)
        ctx.save_for_backward(signs, stats)       # 2 bytes per pooled pixel instead of the two maps
        ctx.cfg = (start_scale, num_scales, n_batch, tuple(pred.shape))
        return ops.msg_loss_value(stats, n_batch, num_scales)

    @staticmethod
    def backward(ctx, grad_out):
        signs, stats = ctx.saved_tensors
        g = ops.msg_loss_grad(signs, ctx.cfg[3], stats, ctx.cfg[0], ctx.cfg[1], 1.0, n_batch=ctx.cfg[2], scale_dev=grad_out)
        return g, None, None, None, None


class MultiScaleGradient(torch.nn.Module):
    """Mirrors model/loss.py:22-63 (`MultiScaleGradient(start_scale=1, num_scales=4)`), device-side, no dynamic shapes:
    one pooling launch over all scales, one reduction launch (sum |Sobel|, count of non-NaN, signs) and two gather launches
    for the gradient (no atomics).  `preview=True`
    (lstm_trainer.py:162-165, TensorBoard only) returns, per scale, the Sobel magnitude of the pooled difference
    (device kernel) resized to (2H, 2W) with the reference's bicubic `torch.nn.Upsample(align_corners=True)`."""

    def __init__(self, start_scale=1, num_scales=4, process_group=None):
        super().__init__()
        self.start_scale, self.num_scales, self.process_group = start_scale, num_scales, process_group

    def forward(self, prediction, target, preview=False):
        if preview:
            _, _, H, W = target.shape
            record = []
            for s in range(self.num_scales):
                mag = ops.msg_sobel_preview(prediction, target, self.start_scale * (2 ** s))
                # logging-only resize (loss.py:42,47); not part of the training arithmetic
                record.append(F.interpolate(mag, size=(2 * H, 2 * W), mode='bicubic', align_corners=True))
            return record
        return _MultiScaleGrad.apply(prediction, target, self.start_scale, self.num_scales, self.process_group)


multi_scale_grad_loss_fn = MultiScaleGradient()


def multi_scale_grad_loss(prediction, target, preview=False):
    return multi_scale_grad_loss_fn.forward(prediction, target, preview)
