"""CPU: the data-IO / metric oracle (oracle/dataio_oracle.py) against outputs of the reference's own functions
(tests/golden/dataio.npz, made by oracle/make_golden_dataio.py)."""
import os

import numpy as np

from oracle import dataio_oracle as D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_voxel_normalisation_matches_reference():
    g = np.load(os.path.join(GOLDEN, 'dataio.npz'))
    cases = D.synth_cases(0)
    for k in ('vox_sparse', 'vox_zero', 'vox_const'):
        np.testing.assert_array_equal(D.normalize_voxel_grid(cases[k]), g[k + '_numpy'])       # same numpy ops: bit-exact
    # the reference's torch twin (one-pass variance in float32) agrees with its numpy twin to rounding
    np.testing.assert_allclose(D.normalize_voxel_grid(cases['vox_sparse']), g['vox_sparse_torch'], rtol=2e-5, atol=2e-6)
    np.testing.assert_array_equal(g['vox_zero_torch'], 0.0)
    out = D.normalize_voxel_grid(cases['vox_sparse'])
    nz = out[out != 0]
    assert abs(float(nz.mean())) < 1e-5 and abs(float(nz.std()) - 1.0) < 1e-5
    np.testing.assert_array_equal(out == 0, cases['vox_sparse'] == 0)


def test_label_transform_and_metrics_match_reference():
    g = np.load(os.path.join(GOLDEN, 'dataio.npz'))
    cases = D.synth_cases(0)
    for clip, reg in ((80.0, 3.70378), (1000.0, 6.2044)):
        lab = D.depth_to_log_label(cases['depth'], clip, reg)
        np.testing.assert_array_equal(lab, g[f'label_{int(clip)}'])
        assert np.array_equal(np.isnan(lab), np.isnan(cases['depth']))
        assert float(np.nanmin(lab)) >= 0.0 and float(np.nanmax(lab)) <= 1.0
    p, t = cases['metric_pred'], cases['metric_target']
    for name, fn in D.METRICS.items():
        np.testing.assert_allclose(fn(p, t), float(g['metric_' + name]), rtol=1e-6, err_msg=name)
