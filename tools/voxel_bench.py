#!/usr/bin/env python
"""BASELINE config 5: events_to_voxel_grid microbench, 10k-10M events -> 5x256x512 grid.

    python tools/voxel_bench.py [--json out.json]

Per size and distribution (uniform / 90% of events on 1% of pixels): CUDA-event time of
ramnet_voxel_grid (events resident in HBM, rotated through buffers > L2 for the large sizes),
algorithmic bytes = 32 B/event + 4*5*H*W B, fraction of the measured HBM peak, and the CPU
baselines: the reference's numpy algorithm (oracle port, single thread by construction) and the
plain-C port.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_ramnet_b200 import ops  # noqa: E402
from rpg_ramnet_b200.utils.synthetic import synth_events  # noqa: E402

W, H, B = 512, 256, 5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    peaks_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    hbm = json.load(open(peaks_path))['hbm_gbs'] if os.path.exists(peaks_path) else 6650.0
    dev = torch.device('cuda', 0)
    import __graft_entry__ as ge
    from oracle import ramnet_oracle as O
    clib = ctypes.CDLL(ge.build_oracle())
    clib.voxel_oracle.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_void_p]
    rows = []
    print(f'{"events":>9s} {"dist":>8s} {"us":>9s} {"Mev/s":>10s} {"GB/s":>8s} {"frac_hbm":>8s} {"numpy Mev/s":>12s} {"C Mev/s":>9s}')
    for n in (10_000, 100_000, 1_000_000, 10_000_000):
        for hot in (False, True):
            ev = synth_events(n, W, H, seed=7, hot=hot)
            nbuf = 1 if n < 4_000_000 else 2
            bufs = [torch.from_numpy(ev).to(dev) for _ in range(nbuf)]
            if n >= 4_000_000:      # 320 MB per buffer: two of them exceed the 126 MB L2
                pass
            for i in range(3):
                ops.voxel_grid_ex(bufs[i % nbuf], B, W, H)
            torch.cuda.synchronize()
            iters = 50 if n <= 1_000_000 else 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            # the host needs ~20 us per call (allocation + ctypes + memset + launch): give the GPU a head start so that
            # the timed launches queue up behind it and the events measure device time, not the Python issue rate
            torch.cuda._sleep(int(2.5e-3 * 1.9e9) if n <= 1_000_000 else 1)
            e0.record()
            for i in range(iters):
                ops.voxel_grid_ex(bufs[i % nbuf], B, W, H)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            nbytes = n * 32 + 4 * B * H * W
            gbs = nbytes / us / 1e3
            # CPU baselines on a bounded sample
            ns = min(n, 1_000_000)
            t0 = time.perf_counter()
            O.voxel_grid(ev[:ns], B, W, H)
            t_np = time.perf_counter() - t0
            out = np.empty((B, H, W), np.float32)
            evc = np.ascontiguousarray(ev[:ns])
            t0 = time.perf_counter()
            clib.voxel_oracle(evc.ctypes.data, ns, B, W, H, out.ctypes.data)
            t_c = time.perf_counter() - t0
            row = dict(events=n, dist='hot' if hot else 'uniform', us=us, mev_s=n / us, gb_s=gbs, frac_hbm=gbs / hbm,
                       numpy_mev_s=ns / t_np / 1e6, c_mev_s=ns / t_c / 1e6, cpu_sample=ns)
            rows.append(row)
            print(f'{n:9d} {row["dist"]:>8s} {us:9.1f} {n / us:10.1f} {gbs:8.1f} {gbs / hbm:8.3f} '
                  f'{row["numpy_mev_s"]:12.2f} {row["c_mev_s"]:9.1f}')
    if args.json:
        json.dump(dict(hbm_peak_gbs=hbm, grid=[B, H, W], rows=rows), open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
