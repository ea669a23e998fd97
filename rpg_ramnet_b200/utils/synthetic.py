"""Seeded synthetic inputs of the reference's wire format: item dicts as
data_loader/dataset.py yields them ({'events{k}': [B,bins,H,W], 'image': [B,1,H,W], 'depth_*'})
and raw event arrays [N,4] = [t, x, y, p].  Used by bench.py, smoke() and the tests."""
from typing import List

import numpy as np
import torch


def synth_sequence(B: int, H: int, W: int, L: int, K: int, seed: int, bins_events: int = 5,
                   bins_rgb: int = 1, with_targets: bool = True) -> List[dict]:
    """L items of {'events{k}': sparse signed voxel grids, 'image': grey in [0,1],
    'depth_*': targets in [0,1] with a 10x10 NaN patch}."""
    g = torch.Generator().manual_seed(seed)
    seq = []
    for _ in range(L):
        item = {}
        for k in range(K):
            item[f'events{k}'] = torch.randn(B, bins_events, H, W, generator=g) * \
                (torch.rand(B, bins_events, H, W, generator=g) < 0.1).float()
        item['image'] = torch.rand(B, bins_rgb, H, W, generator=g)
        if with_targets:
            for key in [f'events{k}' for k in range(K)] + ['image']:
                t = torch.rand(B, 1, H, W, generator=g)
                t[:, :, 3:13, 5:15] = float('nan')
                item['depth_' + key] = t
        seq.append(item)
    return seq


def synth_events(n: int, width: int, height: int, seed: int, hot: bool = False) -> np.ndarray:
    """[n,4] float64 rows [t, x, y, p]; t sorted in [0, 0.05); 'hot' puts 90% of
    events on 1% of the pixels to expose atomic contention."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(0.0, 0.05, n))
    if hot:
        npix = max(1, (width * height) // 100)
        hot_pix = rng.integers(0, width * height, npix)
        pix = np.where(rng.uniform(size=n) < 0.9, hot_pix[rng.integers(0, npix, n)],
                       rng.integers(0, width * height, n))
        x, y = pix % width, pix // width
    else:
        x = rng.integers(0, width, n)
        y = rng.integers(0, height, n)
    p = rng.integers(0, 2, n)
    return np.stack([t, x.astype(np.float64), y.astype(np.float64), p.astype(np.float64)], 1)
