"""Depth metrics on the device (mirrors RAM_Net/model/metric.py:8-57, same function names).

The reference's trainer pulls every prediction and target map to the host and runs numpy / sklearn on
them each step (trainer/lstm_trainer.py:100-106,291-294).  Here one streaming kernel
(ramnet_depth_metrics) reduces the masked error sums per sample in float64 on the device and only
those [N, 8] doubles cross PCIe.  Every function takes [N, 1, H, W] CUDA tensors (y_input, y_target)
and returns a Python float, like the reference's numpy scalars; `eval_metrics` evaluates a list of
them from ONE kernel launch and one small read-back.

NaN convention (as upstream): target pixels without ground truth are NaN and are masked out; the
prediction is finite (a sigmoid output), so the reference's two masks (`~isnan(|t - p|)` and
`~isnan(t)`, metric.py:9-10) coincide.
"""
import numpy as np
import torch

from .. import ops


def _dev(x, like=None):
    """LSTMTrainer._eval_metrics hands numpy arrays to the metric functions (lstm_trainer.py:100-106); tensors may live
    on the host or the device.  Everything is reduced on the GPU: host inputs are uploaded (the optional
    `_eval_metrics` override of INTEGRATION.md §5 avoids that round trip)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not x.is_cuda:
        dev = like.device if (like is not None and like.is_cuda) else torch.device('cuda', torch.cuda.current_device())
        x = x.to(dev, non_blocking=True)
    return x

__all__ = ['abs_rel_diff', 'squ_rel_diff', 'rms_linear', 'scale_invariant_error', 'mean_error', 'median_error', 'mse',
           'eval_metrics']


def _from_sums(name, s):
    """s: [N, 8] float64 sums on the host (see ramnet_depth_metrics)."""
    tot = s.sum(0)
    n = tot[0]
    if name == 'abs_rel_diff':                 # metric.py:8-10
        return float(tot[1] / n)
    if name == 'squ_rel_diff':                 # :12-15
        return float(tot[2] / n)
    if name == 'rms_linear':                   # :17-20
        return float((tot[3] / n) ** 0.5)
    if name == 'scale_invariant_error':        # :22-25 (on |t - p|, as upstream)
        return float(tot[3] / n - (tot[4] / n) ** 2)
    if name == 'mean_error':                   # :27-29
        return float(tot[4] / n)
    if name == 'mse':                          # :35-54, C == 1: per-sample masked MSE averaged over the batch
        return float((s[:, 6] / s[:, 5]).mean())
    raise KeyError(name)


def _median_error(y_input, y_target):
    """metric.py:31-33.  A selection, not a sum: sorted on the device, only the middle values are read back."""
    y_input = _dev(y_input)
    y_target = _dev(y_target, y_input)
    d = (y_target.detach().float() - y_input.detach().float()).abs().flatten()
    d = d[~torch.isnan(d)]
    n = d.numel()
    if n == 0:
        return float('nan')
    v, _ = torch.sort(d)
    return float((v[(n - 1) // 2] + v[n // 2]) / 2)


def eval_metrics(y_input, y_target, names):
    """All of `names` (functions or their names) for one prediction / target pair: the device twin of
    LSTMTrainer._eval_metrics (trainer/lstm_trainer.py:100-106).  Returns a list of floats."""
    names = [n if isinstance(n, str) else n.__name__ for n in names]
    y_input = _dev(y_input)
    y_target = _dev(y_target, y_input)
    sums = None
    out = []
    for n in names:
        if n == 'median_error':
            out.append(_median_error(y_input, y_target))
            continue
        if sums is None:
            sums = ops.depth_metric_sums(y_input, y_target).cpu()
        out.append(_from_sums(n, sums))
    return out


def abs_rel_diff(y_input, y_target, eps=1e-6):
    y_input = _dev(y_input)
    return _from_sums('abs_rel_diff', ops.depth_metric_sums(y_input, _dev(y_target, y_input), eps).cpu())


def squ_rel_diff(y_input, y_target, eps=1e-6):
    y_input = _dev(y_input)
    return _from_sums('squ_rel_diff', ops.depth_metric_sums(y_input, _dev(y_target, y_input), eps).cpu())


def rms_linear(y_input, y_target):
    return eval_metrics(y_input, y_target, ['rms_linear'])[0]


def scale_invariant_error(y_input, y_target):
    return eval_metrics(y_input, y_target, ['scale_invariant_error'])[0]


def mean_error(y_input, y_target):
    return eval_metrics(y_input, y_target, ['mean_error'])[0]


def median_error(y_input, y_target):
    return _median_error(y_input, y_target)


def mse(y_input, y_target):
    return eval_metrics(y_input, y_target, ['mse'])[0]
