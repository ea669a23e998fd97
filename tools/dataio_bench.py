#!/usr/bin/env python
"""Loader wire format / trainer metrics on the device vs the numpy reference path.
    python tools/dataio_bench.py   -> one JSON line per op: us per call, GB/s of algorithmic bytes, CPU (numpy) us"""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops
from rpg_ramnet_b200.model import metric as M
from oracle import dataio_oracle as D          # CPU baseline leg only

dev = torch.device('cuda', 0)
peak = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6650.0


def gpu_us(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def cpu_us(fn, iters=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    return (time.perf_counter() - t0) * 1e6 / iters


B, H, W = 32, 256, 512
rng = np.random.default_rng(0)
vox = (rng.standard_normal((B, 5, H, W)) * (rng.random((B, 5, H, W)) < 0.1)).astype(np.float32)
depth = rng.uniform(0, 120, (B, 1, H, W)).astype(np.float32)
depth[:, :, :32] = np.nan
pred = rng.uniform(0, 1, (B, 1, H, W)).astype(np.float32)
tgt = D.depth_to_log_label(depth, 80.0, 3.70378)
vg, dg, pg, tg = (torch.from_numpy(a).to(dev) for a in (vox, depth, pred, tgt))
names = ['mse', 'abs_rel_diff', 'scale_invariant_error']
rows = [
    ('voxel_normalize (batch of 32 grids, one launch pair)', lambda: ops.voxel_normalize_(vg),
     lambda: [D.normalize_voxel_grid(vox[i]) for i in range(B)], vox.nbytes * 3),
    ('depth_to_label [32,1,256,512]', lambda: ops.depth_to_label(dg, 80.0, 3.70378),
     lambda: D.depth_to_log_label(depth, 80.0, 3.70378), depth.nbytes * 2),
    ('eval_metrics mse+abs_rel+si (incl. the [N,8] read-back)', lambda: M.eval_metrics(pg, tg, names),
     lambda: [D.METRICS[n](pred, tgt) for n in names], pred.nbytes * 2),
]
for name, g, c, nbytes in rows:
    tg_us, tc_us = gpu_us(g), cpu_us(c)
    print(json.dumps({'op': name, 'gpu_us': round(tg_us, 1), 'algorithmic_GBps': round(nbytes / tg_us / 1e3, 1),
                      'frac_of_measured_hbm_peak': round(nbytes / tg_us / 1e3 / peak, 3), 'cpu_numpy_us': round(tc_us, 1),
                      'cpu_cores': 1}))
