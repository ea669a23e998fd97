#!/usr/bin/env python
"""Kernel-boundary cost inside a CUDA graph: a chain of n identical conv launches captured in one graph and replayed,
us per launch, next to the same launch issued eagerly behind a GPU head start.  With RAMNET_PROF=1 (separate run) the
kernel prints its own busy cycles; (in-graph us per launch) - (in-kernel us) is what a kernel boundary costs.
    python tools/gap_bench.py [--n 16] [--only res]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops  # noqa: E402
from tools.layer_bench import LAYERS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--n', type=int, default=16)
    ap.add_argument('--only', default=None)
    ap.add_argument('--batch', type=int, default=4)
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    kind, B, n = ops.MMA_TF32, args.batch, args.n
    print(f'{"layer":30s} {"graph us/launch":>16s} {"eager (head start) us":>22s}')
    for name, H, W, C0, C1, Cout, k, stride, epi in LAYERS:
        if args.only and args.only not in name:
            continue
        nbuf = 4
        xs = [ops.empty_nhwc(B, C0, H, W, dev).normal_() for _ in range(nbuf)]
        x1s = [ops.empty_nhwc(B, C1, H, W, dev).normal_() for _ in range(nbuf)] if C1 else [None] * nbuf
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        w = torch.randn(Cout, C0 + C1, k, k, device=dev) * 0.02
        if epi == ops.EPI_BIAS_RELU and not C1 and ops.s2seg_eligible(C0, Cout, k, stride, kind):
            wp = ops.pack_weights_s2seg(w)
        elif epi == ops.EPI_BIAS_RELU and ops.hpack_eligible(Cout, k, stride, kind):
            wp = ops.pack_weights_hpack(w)
        else:
            wp = ops.pack_weights(w, kind)
        b = torch.zeros(Cout, device=dev)
        Cs = Cout // 2 if epi == ops.EPI_GRU_RU else Cout
        aux0 = ops.empty_nhwc(B, Cs, Ho, Wo, dev).normal_() if epi in (ops.EPI_GRU_RU, ops.EPI_GRU_OUT) else None
        aux1 = ops.empty_nhwc(B, Cs, Ho, Wo, dev).uniform_() if epi == ops.EPI_GRU_OUT else None
        outs = [ops.empty_nhwc(B, Cs, Ho, Wo, dev) for _ in range(2)]
        outs1 = [ops.empty_nhwc(B, Cs, Ho, Wo, dev) for _ in range(2)] if epi == ops.EPI_GRU_RU else [None, None]

        def run(i):
            return ops.conv_fwd(xs[i % nbuf], x1s[i % nbuf], wp, b, Cout, k, stride, epi, kind, aux0=aux0, aux1=aux1,
                                round_tf32=True, out0=outs[i & 1], out1=outs1[i & 1])
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for i in range(3):
                run(i)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(n):
                run(i)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        g_us = e0.elapsed_time(e1) * 1e3 / (10 * n)
        torch.cuda._sleep(int(20e-3 * 1.9e9))
        e0.record()
        for i in range(n):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        e_us = e0.elapsed_time(e1) * 1e3 / n
        print(f'{name:30s} {g_us:16.1f} {e_us:22.1f}')


if __name__ == '__main__':
    main()
