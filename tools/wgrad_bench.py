#!/usr/bin/env python
"""Per-layer timing of the weight-gradient kernel at the bench workload (B=4, 256x512).
    python tools/wgrad_bench.py [--kind tf32|fp32] [--iters 10]"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops
from layer_bench import LAYERS, MULT

ap = argparse.ArgumentParser()
ap.add_argument('--kind', default='tf32'); ap.add_argument('--iters', type=int, default=10); ap.add_argument('--only', default=None)
args = ap.parse_args()
kind = {'tf32': ops.MMA_TF32, 'fp32': ops.MMA_FP32}[args.kind]
dev = torch.device('cuda', 0)
B = 4
tot_ms = tot_fl = 0.0
print(f'{"layer":34s} {"GFLOP":>8s} {"us":>9s} {"TFLOP/s":>8s}')
for name, H, W, C0, C1, Cout, k, stride, epi in LAYERS:
    if args.only and args.only not in name:
        continue
    x0 = ops.empty_nhwc(B, C0, H, W, dev).normal_()
    x1 = ops.empty_nhwc(B, C1, H, W, dev).normal_() if C1 else None
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    dz = ops.empty_nhwc(B, Cout, Ho, Wo, dev).normal_()
    dw = torch.zeros(Cout, C0 + C1, k, k, device=dev)
    for _ in range(2):
        ops.conv_wgrad(dz, x0, x1, Cout, k, stride, dw, None, kind)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        ops.conv_wgrad(dz, x0, x1, Cout, k, stride, dw, None, kind)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / args.iters
    fl = 2.0 * B * Ho * Wo * Cout * (C0 + C1) * k * k
    m = MULT.get(name, 1)
    tot_ms += m * us / 1e3; tot_fl += m * fl
    print(f'{name:34s} {fl / 1e9:8.2f} {us:9.1f} {fl / us / 1e6:8.1f}')
print(f'{"sum over one pass":34s} {tot_fl / 1e9:8.2f} {tot_ms * 1e3:9.1f} {tot_fl / tot_ms / 1e9:8.1f}')
