#!/usr/bin/env python
"""MultiScaleGradient loss (model/loss.py:22-70) on the device: forward (pool + Sobel statistics + value) and backward
(gather gradient) at the bench shape, device time per call.  python tools/msg_bench.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
peak = json.load(open(mp))['hbm_gbs'] if os.path.exists(mp) else 6537.0


def timed(fn, iters=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for N in (4, 32):
    p, t = torch.rand(N, 1, 256, 512, device=dev), torch.rand(N, 1, 256, 512, device=dev)
    t[:, :, :32] = float('nan')
    stats, signs = ops.msg_loss_stats(p, t, 1, 4, want_signs=True)
    one = torch.ones((), device=dev)
    tf = timed(lambda: ops.msg_loss_value(ops.msg_loss_stats(p, t, 1, 4, want_signs=True)[0], N, 4))
    tb = timed(lambda: ops.msg_loss_grad(signs, tuple(p.shape), stats, 1, 4, 1.0, n_batch=N, scale_dev=one))
    px = N * 256 * 512
    # algorithmic bytes: forward reads pred + target once (8 B / pixel); backward writes the gradient once (4 B / pixel)
    print(f'batch {N:2d} x 256x512, 4 scales: forward {tf:6.1f} us ({8.0 * px / tf / 1e3:6.0f} GB/s, {8.0 * px / tf / 1e3 / peak:.2f} of HBM), '
          f'backward {tb:6.1f} us ({4.0 * px / tb / 1e3:6.0f} GB/s)')
