#!/usr/bin/env python
"""Per-layer timing of the RAM-Net convolutions at the bench workload (B=4, 256x512).

    python tools/layer_bench.py [--kind tf32|fp32] [--iters 20] [--batch 4]

Prints one line per distinct conv of a pass: algorithmic GFLOP, average CUDA-event time over
`iters` launches (inputs rotated through buffers larger than L2), TFLOP/s and the fraction of the
measured bf16 peak.  Development aid; its output is what profiles/*layers*.txt record.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops  # noqa: E402

# name, H_in, W_in, C0, C1, Cout, k, stride, epilogue
LAYERS = [
    ('enc0 5x5s2 32->64', 256, 512, 32, 0, 64, 5, 2, ops.EPI_BIAS_RELU),
    ('enc1 5x5s2 64->128', 128, 256, 64, 0, 128, 5, 2, ops.EPI_BIAS_RELU),
    ('enc2 5x5s2 128->256', 64, 128, 128, 0, 256, 5, 2, ops.EPI_BIAS_RELU),
    ('gru0 RU 3x3 64+64->128', 128, 256, 64, 64, 128, 3, 1, ops.EPI_GRU_RU),
    ('gru0 OUT 3x3 64+64->64', 128, 256, 64, 64, 64, 3, 1, ops.EPI_GRU_OUT),
    ('gru1 RU 3x3 128+128->256', 64, 128, 128, 128, 256, 3, 1, ops.EPI_GRU_RU),
    ('gru1 OUT 3x3 128+128->128', 64, 128, 128, 128, 128, 3, 1, ops.EPI_GRU_OUT),
    ('gru2 RU 3x3 256+256->512', 32, 64, 256, 256, 512, 3, 1, ops.EPI_GRU_RU),
    ('gru2 OUT 3x3 256+256->256', 32, 64, 256, 256, 256, 3, 1, ops.EPI_GRU_OUT),
    ('res 3x3 256->256 (x4)', 32, 64, 256, 0, 256, 3, 1, ops.EPI_BIAS_RELU),
    ('dec0 5x5 256->128', 64, 128, 256, 0, 128, 5, 1, ops.EPI_BIAS_RELU),
    ('dec1 5x5 128->64', 128, 256, 128, 0, 64, 5, 1, ops.EPI_BIAS_RELU),
    ('dec2 5x5 64->32', 256, 512, 64, 0, 32, 5, 1, ops.EPI_BIAS_RELU),
]
# up-conv mode (bilinear x2 + 5x5 conv in one launch on the low-resolution input): name, H_lo, W_lo, Cin, Cout, epilogue
UP_LAYERS = [
    ('dec0 upconv 256->128', 32, 64, 256, 128, ops.EPI_BIAS_RELU),
    ('dec1 upconv 128->64', 64, 128, 128, 64, ops.EPI_BIAS_RELU),
    ('dec2 upconv 64->32', 128, 256, 64, 32, ops.EPI_BIAS_RELU),
    ('dec2 upconv 64->32 +pred', 128, 256, 64, 32, ops.EPI_BIAS_RELU_PRED),
]
MULT = {'res 3x3 256->256 (x4)': 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--kind', default='tf32')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--only', default=None)
    args = ap.parse_args()
    kind = {'tf32': ops.MMA_TF32, 'fp32': ops.MMA_FP32}[args.kind]
    dev = torch.device('cuda', 0)
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(
        os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {'bf16_tflops': 1590.0}
    peak = peaks['bf16_tflops']
    B = args.batch
    total_ms = total_fl = 0.0
    print(f'# kind={args.kind} batch={B} iters={args.iters} peak(bf16 burst)={peak} TF')
    print(f'{"layer":34s} {"GFLOP":>8s} {"us":>9s} {"TFLOP/s":>8s} {"frac":>6s}')
    for name, H, W, C0, C1, Cout, k, stride, epi in LAYERS:
        if args.only and args.only not in name:
            continue
        nbuf = max(2, int(300e6 // (B * H * W * (C0 + C1) * 4)) + 1)
        nbuf = min(nbuf, 8)
        xs = [ops.empty_nhwc(B, C0, H, W, dev).normal_() for _ in range(nbuf)]
        x1s = [ops.empty_nhwc(B, C1, H, W, dev).normal_() for _ in range(nbuf)] if C1 else [None] * nbuf
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        w = torch.randn(Cout, C0 + C1, k, k, device=dev) * 0.02
        hp = epi == ops.EPI_BIAS_RELU and ops.hpack_eligible(Cout, k, stride, kind)
        if epi == ops.EPI_BIAS_RELU and not C1 and ops.s2seg_eligible(C0, Cout, k, stride, kind):
            wp = ops.pack_weights_s2seg(w)
        else:
            wp = ops.pack_weights_hpack(w) if hp else ops.pack_weights(w, kind)
        b = torch.zeros(Cout, device=dev)
        Cs = Cout // 2 if epi == ops.EPI_GRU_RU else Cout
        aux0 = ops.empty_nhwc(B, Cs, Ho, Wo, dev).normal_() if epi in (ops.EPI_GRU_RU, ops.EPI_GRU_OUT) else None
        aux1 = ops.empty_nhwc(B, Cs, Ho, Wo, dev).uniform_() if epi == ops.EPI_GRU_OUT else None

        def run(i):
            return ops.conv_fwd(xs[i % nbuf], x1s[i % nbuf], wp, b, Cout, k, stride, epi, kind, aux0=aux0, aux1=aux1,
                                round_tf32=True)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        fl = 2.0 * B * Ho * Wo * Cout * (C0 + C1) * k * k
        tf = fl / us / 1e6
        mult = MULT.get(name, 1)
        total_ms += mult * us / 1e3
        total_fl += mult * fl
        print(f'{name:34s} {fl / 1e9:8.2f} {us:9.1f} {tf:8.1f} {tf / peak:6.3f}')
    if total_ms > 0:
      print(f'{"sum over one pass (convs only)":34s} {total_fl / 1e9:8.2f} {total_ms * 1e3:9.1f} '
            f'{total_fl / total_ms / 1e9:8.1f} {total_fl / total_ms / 1e9 / peak:6.3f}')
    for name, H, W, Cin, Cout, epi in UP_LAYERS:
        if args.only and args.only not in name:
            continue
        nbuf = 4
        xs = [ops.empty_nhwc(B, Cin, H, W, dev).normal_() for _ in range(nbuf)]
        w = torch.randn(Cout, Cin, 5, 5, device=dev) * 0.02
        wp = ops.pack_weights_upconv(w)
        b = torch.zeros(Cout, device=dev)
        pw, pb = torch.randn(Cout, device=dev), torch.zeros(1, device=dev)

        def run(i):
            if epi == ops.EPI_BIAS_RELU_PRED:
                return ops.conv_up_fwd(xs[i % nbuf], wp, b, Cout, epi, aux0=pw, aux1=pb)
            return ops.conv_up_fwd(xs[i % nbuf], wp, b, Cout, epi)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        fl = 2.0 * B * (2 * H) * (2 * W) * Cout * Cin * 25
        print(f'{name:34s} {fl / 1e9:8.2f} {us:9.1f} {fl / us / 1e6:8.1f} {fl / us / 1e6 / peak:6.3f}')


if __name__ == '__main__':
    main()
