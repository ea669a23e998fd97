"""CPU: numpy and C restatements of events_to_voxel_grid vs the reference's outputs."""
import ctypes
import os

import numpy as np
import pytest

from oracle import ramnet_oracle as O
from helpers import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    g = np.load(os.path.join(GOLDEN, 'voxel.npz'))
    names = sorted({k.split('/')[0] for k in g.files})
    return g, names


def _c_oracle():
    import __graft_entry__ as ge
    path = ge.build_oracle()
    lib = ctypes.CDLL(path)
    lib.voxel_oracle.restype = ctypes.c_int
    lib.voxel_oracle.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                 ctypes.c_void_p]
    return lib


def test_numpy_oracle_bit_exact_vs_reference():
    g, names = _cases()
    assert len(names) >= 12
    for n in names:
        b, w, h = (int(v) for v in g[n + '/shape'])
        ev = g[n + '/events']
        keep = ev.copy()
        out = O.voxel_grid(ev, b, w, h)
        assert np.array_equal(ev, keep), 'oracle must not mutate its input'
        assert out.dtype == np.float32 and out.shape == (b, h, w)
        assert np.array_equal(out, g[n + '/grid']), n      # same sequential fp32 accumulation order


def test_c_oracle_bit_exact_vs_reference():
    lib = _c_oracle()
    g, names = _cases()
    for n in names:
        b, w, h = (int(v) for v in g[n + '/shape'])
        ev = np.ascontiguousarray(g[n + '/events'], np.float64)
        out = np.empty((b, h, w), np.float32)
        lib.voxel_oracle(ev.ctypes.data, ev.shape[0], b, w, h, out.ctypes.data)
        assert np.array_equal(out, g[n + '/grid']), n


def test_empty_and_votes():
    assert not O.voxel_grid(np.zeros((0, 4)), 5, 8, 4).any()
    ev = O.synth_events(1000, 32, 24, 99)
    il, vl, ir, vr = O.voxel_grid_votes(ev, 5, 32, 24)
    grid = np.zeros(5 * 24 * 32, np.float32)
    np.add.at(grid, il[il >= 0], vl[il >= 0])
    np.add.at(grid, ir[ir >= 0], vr[ir >= 0])
    assert np.array_equal(grid.reshape(5, 24, 32), O.voxel_grid(ev, 5, 32, 24))
    # last timestamp: ti = B-1, dt = 0 -> left vote in the last bin, right vote dropped
    assert il[-1] // (32 * 24) == 4 and ir[-1] == -1
