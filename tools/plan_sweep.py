#!/usr/bin/env python
"""Sweep the halo kernel's tile configurations per layer (RAMNET_HALO_FORCE, read by the library at every launch) and print
the planner's own choice next to the best measured one.  Development aid for the cost model; output -> profiles/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops  # noqa: E402
from tools.layer_bench import LAYERS  # noqa: E402

SHAPES = [(1, 1), (2, 1), (4, 1), (2, 2), (1, 2)]


def time_layer(run, iters=10):
    for i in range(2):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


def main():
    dev = torch.device('cuda', 0)
    B = 4
    only = sys.argv[1] if len(sys.argv) > 1 else None
    kind = ops.MMA_TF32
    tot_def = tot_best = 0.0
    for name, H, W, C0, C1, Cout, k, stride, epi in LAYERS:
        if only and only not in name:
            continue
        if stride == 2:
            continue        # the parity-plane path has its own planner (RAMNET_S2_FORCE, profiles/r02_s2seg.txt)
        nbuf = min(8, max(2, int(300e6 // (B * H * W * (C0 + C1) * 4)) + 1))
        xs = [ops.empty_nhwc(B, C0, H, W, dev).normal_() for _ in range(nbuf)]
        x1s = [ops.empty_nhwc(B, C1, H, W, dev).normal_() for _ in range(nbuf)] if C1 else [None] * nbuf
        w = torch.randn(Cout, C0 + C1, k, k, device=dev) * 0.02
        wp = ops.pack_weights(w, kind)          # plain pack: hpack layers are outside this sweep
        b = torch.zeros(Cout, device=dev)
        Cs = Cout // 2 if epi == ops.EPI_GRU_RU else Cout
        aux0 = ops.empty_nhwc(B, Cs, H, W, dev).normal_() if epi in (ops.EPI_GRU_RU, ops.EPI_GRU_OUT) else None
        aux1 = ops.empty_nhwc(B, Cs, H, W, dev).uniform_() if epi == ops.EPI_GRU_OUT else None

        def run(i):
            return ops.conv_fwd(xs[i % nbuf], x1s[i % nbuf], wp, b, Cout, k, stride, epi, kind, aux0=aux0, aux1=aux1,
                                round_tf32=True)
        os.environ.pop('RAMNET_HALO_FORCE', None)
        t_def = time_layer(run)
        res = []
        for ptx, pty in SHAPES:
            for bn in (256, 128, 64, 32):
                if Cout % bn or ptx * pty * bn > 512:
                    continue
                for pair in (1, 0):
                    os.environ['RAMNET_HALO_FORCE'] = f'{ptx},{pty},{bn},0,0,{pair}'
                    try:
                        t = time_layer(run)
                    except Exception as e:      # configuration does not fit
                        continue
                    res.append((t, ptx, pty, bn, pair))
        os.environ.pop('RAMNET_HALO_FORCE', None)
        res.sort()
        mult = 4 if name.startswith('res') else 1
        tot_def += mult * t_def
        tot_best += mult * min(t_def, res[0][0])
        top = '  '.join(f'{t:.1f}us@{ptx}x{pty},bn{bn},p{pair}' for t, ptx, pty, bn, pair in res[:4])
        print(f'{name:30s} planner {t_def:6.1f} us | best {top}', flush=True)
    print(f'sum planner {tot_def:.1f} us, sum best {tot_best:.1f} us')


if __name__ == '__main__':
    main()
