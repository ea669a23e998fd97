"""Data-parallel plumbing for the RAM-Net path: one process per GPU, batch sharded.

The forward path needs no collective.  Two exchanges exist on the training side:
  * the scale-invariant loss normalises by the GLOBAL count of valid pixels and subtracts the
    GLOBAL mean (model/loss.py:7-9), so equivalence with a single-process batch needs an all-reduce
    of (sum d, sum d^2, n) per loss term between ramnet_si_loss_stats and ramnet_si_loss_grad;
  * one all-reduce (sum) of the flat fp32 gradient buffer before ramnet_adam_step.
Both are thin wrappers over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist


def shard_item(item: Dict[str, torch.Tensor], rank: int, world_size: int) -> Dict[str, torch.Tensor]:
    """Rank r takes samples [r*B/W, (r+1)*B/W) of every tensor of a data-loader item."""
    out = {}
    for k, v in item.items():
        B = v.shape[0]
        if B % world_size:
            raise ValueError(f'batch {B} of {k!r} is not divisible by world size {world_size}')
        per = B // world_size
        out[k] = v[rank * per:(rank + 1) * per]
    return out


def all_reduce_loss_stats(stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """In-place sum of the float64 [3] statistics (sum d, sum d^2, n) over the data-parallel group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def si_loss_from_stats(stats: torch.Tensor, weight: float = 1.0, n_lambda: float = 1.0) -> torch.Tensor:
    """w * (mean(d^2) - lambda * mean(d)^2) from (sum d, sum d^2, n)  (model/loss.py:9)."""
    n = stats[2]
    return weight * (stats[1] / n - n_lambda * (stats[0] / n) ** 2)


def si_grad_from_stats(pred: torch.Tensor, target: torch.Tensor, stats: torch.Tensor, weight: float = 1.0,
                       n_lambda: float = 1.0) -> torch.Tensor:
    """d loss / d pred with GLOBAL statistics: (2w/n)(d - lambda*mean(d)), 0 at NaN.  Host-side reference
    of ramnet_si_loss_grad used by the CPU tests; the product path calls the CUDA kernel."""
    d = pred - target
    ok = ~torch.isnan(d)
    n, mean = stats[2].to(pred.dtype), (stats[0] / stats[2]).to(pred.dtype)
    return torch.where(ok, (2.0 * weight / n) * (d - n_lambda * mean), torch.zeros_like(d))


def all_reduce_flat_grads(flat_grad: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum of the flat fp32 gradient buffer (59.5 MB for the shipped block) — gradients computed with
    global loss statistics are summed, not averaged."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad
