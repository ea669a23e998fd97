#!/usr/bin/env python
"""Timing of the head convolution variants at the bench workload (B=4, 256x512): direct fp32 kernel vs
im2row + tensor-core 5x1 conv.   python tools/head_bench.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops

dev = torch.device('cuda', 0)
B, H, W = 4, 256, 512


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for Cin in (5, 1):
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(32, Cin, 5, 5, device=dev) * 0.1
    b = torch.zeros(32, device=dev)
    wp = ops.pack_weights_head(w)
    xe = ops.head_im2row(x)
    t_direct = timeit(lambda: ops.head_conv(x, w, b, True))
    t_im2row = timeit(lambda: ops.head_im2row(x))
    t_conv = timeit(lambda: ops.head_conv_tc(xe, wp, b, Cin, 32, True))
    print(f'Cin={Cin}: direct fp32 {t_direct:.1f} us | im2row {t_im2row:.1f} us + tcgen05 5x1 conv {t_conv:.1f} us')
