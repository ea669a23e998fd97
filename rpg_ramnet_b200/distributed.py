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


def dp_world_size(group: Optional[dist.ProcessGroup] = None) -> int:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(None if group in (None, True) else group)
    return 1


class BucketedAllReduce:
    """Sum of the flat gradient buffer in `n_buckets` contiguous slices on a side stream.

    `launch()` enqueues every slice's all-reduce on the communication stream (after the gradients written on the
    current stream so far) and yields `(lo, hi, ready)` per slice in order; `ready()` makes the current stream wait for
    THAT slice only, so the consumer of slice k (its Adam launch) runs while slices k+1.. are still on the wire.  With
    one rank (or one bucket and no process group) nothing is enqueued and `ready` is a no-op.  Works on CPU tensors
    (gloo, synchronous) for the host-logic tests.  Inside CUDA-graph capture the side stream forks from and joins back
    into the capturing stream through the recorded events, so the captured graph carries the same overlap."""

    def __init__(self, flat: torch.Tensor, n_buckets: int = 4, group: Optional[dist.ProcessGroup] = None):
        self.flat, self.group = flat, (None if group in (None, True) else group)
        n = flat.numel()
        n_buckets = max(1, min(int(n_buckets), max(1, n // 4)))
        # 16-byte aligned cuts so every slice keeps the vector alignment the Adam kernel checks
        cuts = sorted({0, n} | {((n * k // n_buckets) // 4) * 4 for k in range(1, n_buckets)})
        self.spans = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        self.stream = torch.cuda.Stream(device=flat.device) if flat.is_cuda else None

    def launch(self):
        if dp_world_size(self.group) <= 1:
            for lo, hi in self.spans[:1]:
                yield 0, self.flat.numel(), (lambda: None)
            return
        if self.stream is None:                 # CPU (gloo): synchronous
            for lo, hi in self.spans:
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                yield lo, hi, (lambda: None)
            return
        cur = torch.cuda.current_stream(self.flat.device)
        self.stream.wait_stream(cur)            # gradients are final on the compute stream
        events = []
        with torch.cuda.stream(self.stream):
            for lo, hi in self.spans:
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                events.append(ev)
        for (lo, hi), ev in zip(self.spans, events):
            yield lo, hi, (lambda ev=ev: cur.wait_event(ev))
