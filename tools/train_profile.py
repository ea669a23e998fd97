#!/usr/bin/env python
"""Per-shape breakdown of one training step (fwd + bwd) of the bench workload: CUDA-event time of every conv /
dgrad / wgrad launch, aggregated by shape.   python tools/train_profile.py [--kind tf32]"""
import argparse, collections, contextlib, io, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rpg_ramnet_b200 as R
from rpg_ramnet_b200 import ops
from rpg_ramnet_b200.utils.synthetic import synth_sequence
import bench

ap = argparse.ArgumentParser(); ap.add_argument('--kind', default='tf32'); ap.add_argument('--L', type=int, default=2)
args = ap.parse_args()
dev = torch.device('cuda', 0)
model = bench.build_model(torch, 0, args.kind, cuda_graphs=False).train()
items = [{k: v.to(dev) for k, v in it.items()} for it in synth_sequence(bench.B, bench.H, bench.W, args.L, 1, seed=2)]
def step():
    model.zero_grad()
    ps, pl, terms = None, {'events0': None, 'image': None}, []
    for it in items:
        preds, s, pl = model(it, ps, pl)
        ps = s['image']
        terms += [R.scale_invariant_loss(preds[k], it['depth_' + k]) for k in preds]
    sum(terms).backward()
step(); step()
ops.PROFILE = []
step(); torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for p in prof:
    tag = p[4] if len(p) > 4 and p[4] else p[0]
    a = agg[tag]; a[0] += 1; a[1] += p[2].elapsed_time(p[3]); a[2] += p[1]
tot = sum(a[1] for a in agg.values())
print(f'# one training step, L={args.L}, batch {bench.B}, {bench.W}x{bench.H}, kind={args.kind}: {tot:.1f} ms in profiled kernels')
for tag, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{a[1]:8.2f} ms {100*a[1]/tot:5.1f}%  n={a[0]:3d}  {a[2]/max(a[1],1e-9)/1e9:7.1f} TF/s  {tag}')
