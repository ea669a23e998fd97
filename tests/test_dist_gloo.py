"""CPU, world_size=2, gloo: the data-parallel host logic (batch sharding + exact global SI loss via the
3-float statistics all-reduce + summed gradients) equals the single-process full-batch result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ramnet_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from rpg_ramnet_b200 import distributed as D
    g = torch.Generator().manual_seed(5)
    pred = torch.rand(4, 1, 16, 24, generator=g)
    target = torch.rand(4, 1, 16, 24, generator=g)
    target[0, :, 2:9, 3:11] = float('nan')          # unequal NaN counts per shard
    item = D.shard_item({'pred': pred, 'target': target}, rank, world)
    p, t = item['pred'], item['target']
    d = p - t
    ok = ~torch.isnan(d)
    stats = torch.stack([d[ok].double().sum(), (d[ok].double() ** 2).sum(), ok.sum().double()])
    local = stats.clone()
    D.all_reduce_loss_stats(stats)
    loss = D.si_loss_from_stats(stats, 1.0, 0.85)
    grad = D.si_grad_from_stats(p, t, stats, 1.0, 0.85)
    flat = torch.full((10,), float(rank + 1))
    D.all_reduce_flat_grads(flat)
    ret[rank] = dict(loss=float(loss), grad=grad.numpy(), local_n=float(local[2]), flat=flat.numpy())
    dist.destroy_process_group()


def test_exact_global_si_loss_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    g = torch.Generator().manual_seed(5)
    pred = torch.rand(4, 1, 16, 24, generator=g)
    target = torch.rand(4, 1, 16, 24, generator=g)
    target[0, :, 2:9, 3:11] = float('nan')
    ref_loss = O.si_loss(pred.double(), target.double(), 1.0, 0.85).item()
    ref_grad = O.si_loss_grad(pred, target, 1.0, 0.85).numpy()
    assert ret[0]['local_n'] != ret[1]['local_n']            # shards really differ
    for r in range(world):
        assert abs(ret[r]['loss'] - ref_loss) <= 1e-12
        np.testing.assert_allclose(ret[r]['grad'], ref_grad[2 * r:2 * r + 2], rtol=1e-5, atol=1e-9)
        np.testing.assert_array_equal(ret[r]['flat'], np.full(10, 3.0, np.float32))
    # the per-shard losses do NOT average to the global loss (why the statistics all-reduce exists)
    l0 = O.si_loss(pred[:2].double(), target[:2].double(), 1.0, 0.85).item()
    l1 = O.si_loss(pred[2:].double(), target[2:].double(), 1.0, 0.85).item()
    assert abs(0.5 * (l0 + l1) - ref_loss) > 1e-6


def test_shard_item_rejects_ragged_batches():
    import pytest
    from rpg_ramnet_b200 import distributed as D
    with pytest.raises(ValueError):
        D.shard_item({'image': torch.zeros(3, 1, 4, 4)}, 0, 2)


def _bucket_worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from rpg_ramnet_b200 import distributed as D
    from rpg_ramnet_b200.model.loss import _resolve_group
    flat = torch.arange(1003, dtype=torch.float32) * (rank + 1)
    red = D.BucketedAllReduce(flat, n_buckets=4)
    spans = []
    for lo, hi, ready in red.launch():
        ready()
        spans.append((lo, hi))
    # default loss semantics under data parallelism: global statistics (None), local only on request (False)
    ret[rank] = dict(flat=flat.numpy(), spans=spans, auto=_resolve_group(None)[0], off=_resolve_group(False)[0],
                     world=D.dp_world_size())
    dist.destroy_process_group()


def test_bucketed_flat_gradient_all_reduce_world2():
    """The bucketed all-reduce of FusedAdam: contiguous 16-byte aligned slices that cover the buffer exactly once, each
    summed over the ranks; the loss's process_group default resolves to the global-statistics exchange when
    world_size > 1 (summed gradients are only right for global-statistics losses)."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, port, ret), nprocs=world, join=True)
    want = np.arange(1003, dtype=np.float32) * 3.0
    for r in range(world):
        np.testing.assert_array_equal(ret[r]['flat'], want)
        spans = ret[r]['spans']
        assert len(spans) == 4 and spans[0][0] == 0 and spans[-1][1] == 1003
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert all(lo % 4 == 0 for lo, _ in spans)
        assert ret[r]['auto'] is True and ret[r]['off'] is False and ret[r]['world'] == 2


def test_single_process_defaults_are_local():
    from rpg_ramnet_b200 import distributed as D
    from rpg_ramnet_b200.model.loss import _resolve_group
    assert D.dp_world_size() == 1
    assert _resolve_group(None) == (False, None) and _resolve_group(True) == (False, None)
    flat = torch.ones(10)
    out = list(D.BucketedAllReduce(flat, 4).launch())
    assert [(lo, hi) for lo, hi, _ in out] == [(0, 10)]
