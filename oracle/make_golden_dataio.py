#!/usr/bin/env python
"""Generates tests/golden/dataio.npz from the UNMODIFIED reference (container only):
    python oracle/make_golden_dataio.py
* voxel normalisation: SynchronizedFramesEventsRawDataset.normalize_voxelgrid (data_loader/dataset_asynchronous.py:300-308,
  numpy) and EventPreprocessor.__call__ (utils/event_tensor_utils.py:34-68, torch twin)
* metrics: every function of model/metric.py:8-54 the shipped configs list
* label transform: the numpy expressions of data_loader/dataset.py:296-305 (inline code upstream -> restated here)
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import dataio_oracle as D, ref_import  # noqa: E402


def main():
    ref_import.load()
    for name, attrs in (('skimage', {}), ('skimage.measure', {'compare_ssim': None}), ('skimage.io', {}),
                        ('skimage.transform', {})):
        ref_import._stub(name, **attrs)
    import model.metric as ref_metric
    from data_loader.dataset_asynchronous import SynchronizedFramesEventsRawDataset as RefDS
    import contextlib
    import utils.event_tensor_utils as ref_etu
    ref_etu.CudaTimer = lambda *_a, **_k: contextlib.nullcontext()    # timing scaffold needs a GPU; not arithmetic
    EventPreprocessor = ref_etu.EventPreprocessor
    cases = D.synth_cases(0)
    out = {}
    opts = types.SimpleNamespace(no_normalize=False, hot_pixels_file=None, flip=False)
    pre = EventPreprocessor(opts)
    for k in ('vox_sparse', 'vox_zero', 'vox_const'):
        out[k + '_numpy'] = RefDS.normalize_voxelgrid(None, cases[k].copy())
        if k != 'vox_const':      # the torch twin divides by a zero stddev there (NaNs): not a case to pin
            out[k + '_torch'] = pre(torch.from_numpy(cases[k].copy())[None]).numpy()[0]
    for clip, reg in ((80.0, 3.70378), (1000.0, 6.2044)):
        frame = np.clip(cases['depth'], 0.0, clip) / clip
        with np.errstate(divide='ignore', invalid='ignore'):
            frame = (1.0 + np.log(frame) / reg).clip(0, 1.0)
        out[f'label_{int(clip)}'] = frame.astype(np.float32)
    p, t = cases['metric_pred'], cases['metric_target']
    for name in ('abs_rel_diff', 'squ_rel_diff', 'rms_linear', 'scale_invariant_error', 'mean_error', 'median_error', 'mse'):
        out['metric_' + name] = np.float64(getattr(ref_metric, name)(p.copy(), t.copy()))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'dataio.npz'), **out)
    print({k: (v.shape if getattr(v, 'shape', ()) else float(v)) for k, v in out.items()})


if __name__ == '__main__':
    main()
