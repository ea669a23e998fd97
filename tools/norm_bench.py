#!/usr/bin/env python
"""Live norm layers (ramnet_norm_fwd / ramnet_norm_bwd) on the layer shapes of the shipped block at batch 4, 256x512:
device time per call (CUDA events, inputs rotated through > L2), achieved GB/s of algorithmic bytes against the
measured HBM peak, next to torch CPU F.batch_norm on one shape.
    python tools/norm_bench.py            (also the target of the ncu launch list: profiles/r02_norm_kernels.txt)
Algorithmic bytes: forward = read z twice (statistics, apply) + write y = 12 B / element (+4 with a residual);
backward = read (dy, y, z) twice + write dz = 28 B / element."""
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops  # noqa: E402

dev = torch.device('cuda', 0)
mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
peak = json.load(open(mp))['hbm_gbs'] if os.path.exists(mp) else 6537.0
SHAPES = [('enc0 / gru level 0', 4, 64, 128, 256), ('enc1', 4, 128, 64, 128), ('enc2 / resblocks', 4, 256, 32, 64),
          ('dec2', 4, 32, 256, 512), ('pred (C = 1)', 4, 1, 256, 512)]
ROT = 6           # > 126 MB of L2 across the rotation for the big shapes


def timed(fn, iters=30):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(20_000_000)       # GPU head start: the launches queue up and run back to back
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


print(f'# HBM peak {peak:.0f} GB/s (MEASURED_PEAKS.json); us per call = statistics + finalize + apply launches')
print(f'{"layer":22s} {"shape":>18s} {"kind":>4s} {"fwd us":>8s} {"GB/s":>7s} {"frac":>6s} {"bwd us":>8s} {"GB/s":>7s} {"frac":>6s}')
for name, N, C, H, W in SHAPES:
    for kind in ('BN', 'IN'):
        zs = [ops.empty_nhwc(N, C, H, W, dev).normal_() for _ in range(ROT)]
        dys = [ops.empty_nhwc(N, C, H, W, dev).normal_() for _ in range(ROT)]
        gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        act = 'sigmoid' if C == 1 else 'relu'
        y, stats = ops.norm_fwd(zs[0], kind, act, gamma, beta, None, rm, rv, 0.1, 1e-5, True, True)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        n = N * C * H * W
        tf = timed(lambda i: ops.norm_fwd(zs[i % ROT], kind, act, gamma, beta, None, rm, rv, 0.1, 1e-5, True, True))
        tb = timed(lambda i: ops.norm_bwd(dys[i % ROT], y, zs[i % ROT], stats, kind, act, gamma, True, True, False, dg, db))
        gf, gb = 12.0 * n / tf / 1e3, 28.0 * n / tb / 1e3
        print(f'{name:22s} {str((N, C, H, W)):>18s} {kind:>4s} {tf:8.1f} {gf:7.0f} {gf / peak:6.2f} {tb:8.1f} {gb:7.0f} {gb / peak:6.2f}')
        del zs, dys
z = torch.randn(4, 64, 128, 256)
rm, rv = torch.zeros(64), torch.ones(64)
F.batch_norm(z, rm, rv, None, None, True, 0.1, 1e-5)
t0 = time.perf_counter()
for _ in range(5):
    torch.relu(F.batch_norm(z, rm, rv, None, None, True, 0.1, 1e-5))
print(f'# torch CPU F.batch_norm + relu, (4, 64, 128, 256), {torch.get_num_threads()} threads: {(time.perf_counter() - t0) / 5 * 1e6:.0f} us')
