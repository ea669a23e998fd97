"""CPU: invariants of the launch planners (forward halo kernel, tap-packed weight gradient) over a sweep of layer shapes,
through the host-only C entry ramnet_plan_describe (no CUDA call).  A configuration that violates one of these would
fail at launch (shared memory / TMEM) or silently mis-tile on the GPU."""
import ctypes
import itertools

import pytest

from rpg_ramnet_b200 import _lib
from rpg_ramnet_b200._lib import ConvDesc

SMEM_LIMIT = 227 * 1024          # dynamic shared memory per CTA on sm_100
EPI_BIAS_RELU, MMA_TF32 = 1, 1


def describe(N, H, W, C0, C1, Cout, k, stride, sm_count=148, epilogue=EPI_BIAS_RELU):
    lib = _lib.load()
    d = ConvDesc(N, H, W, C0, C1, Cout, k, stride, epilogue, MMA_TF32, 0, 0)
    buf = ctypes.create_string_buffer(2048)
    n = lib.ramnet_plan_describe(ctypes.byref(d), sm_count, buf, 2048)
    assert n > 0
    out = {}
    for kv in buf.value.decode().split():
        key, val = kv.split('=')
        out.setdefault(key + ('_w' if key in out else ''), int(val))
    return out


def split_sections(N, H, W, C0, C1, Cout, k, stride, **kw):
    """-> (halo dict or None, wgrad dict or None); the line is 'halo=.. <halo keys> wgrad=.. <wgrad keys>'."""
    lib = _lib.load()
    d = ConvDesc(N, H, W, C0, C1, Cout, k, stride, kw.get('epilogue', EPI_BIAS_RELU), MMA_TF32, 0, 0)
    buf = ctypes.create_string_buffer(2048)
    assert lib.ramnet_plan_describe(ctypes.byref(d), kw.get('sm_count', 148), buf, 2048) > 0
    text = buf.value.decode()
    h_txt, w_txt = text.split('wgrad=')
    halo = dict((a, int(b)) for a, b in (kv.split('=') for kv in h_txt.split()))
    wg = dict((a, int(b)) for a, b in (kv.split('=') for kv in ('wgrad=' + w_txt).split()))
    return (halo if halo['halo'] else None), (wg if wg['wgrad'] else None)


SHAPES = [(N, H, W, C0, C1, Cout, k, s)
          for (N, (H, W), (C0, C1), Cout, k, s) in itertools.product(
              (1, 4), ((256, 512), (128, 256), (64, 88), (32, 64), (16, 24)),
              ((32, 0), (64, 0), (64, 64), (128, 128), (256, 0), (256, 256), (160, 0)),
              (32, 64, 128, 256, 512, 160), (3, 5), (1, 2))]


def test_bench_layers_have_plans():
    """Every conv of the shipped model at the bench resolution is covered by both planners."""
    layers = [(256, 512, 32, 0, 64, 5, 2), (128, 256, 64, 0, 128, 5, 2), (64, 128, 128, 0, 256, 5, 2),
              (128, 256, 64, 64, 128, 3, 1), (128, 256, 64, 64, 64, 3, 1), (64, 128, 128, 128, 256, 3, 1),
              (64, 128, 128, 128, 128, 3, 1), (32, 64, 256, 256, 512, 3, 1), (32, 64, 256, 256, 256, 3, 1),
              (32, 64, 256, 0, 256, 3, 1), (64, 128, 256, 0, 128, 5, 1), (128, 256, 128, 0, 64, 5, 1),
              (256, 512, 64, 0, 32, 5, 1)]
    for (H, W, C0, C1, Cout, k, s) in layers:
        halo, wg = split_sections(4, H, W, C0, C1, Cout, k, s)
        assert halo is not None and wg is not None, (H, W, C0, C1, Cout, k, s)
        assert wg['problems'] == (4 if s == 2 else 1)


@pytest.mark.parametrize('sm_count', [148, 132])
def test_planner_invariants(sm_count):
    n_halo = n_wg = 0
    for shape in SHAPES:
        N, H, W, C0, C1, Cout, k, s = shape
        halo, wg = split_sections(*shape, sm_count=sm_count)
        if halo is not None:
            n_halo += 1
            tiles = halo['ptx'] * halo['pty']
            assert 1 <= tiles <= 4 and halo['ptx'] in (1, 2, 4), shape
            assert Cout % halo['bn'] == 0 and halo['bn'] % 16 == 0 and halo['bn'] <= 256, shape
            assert halo['tmem_cols'] <= 512 and halo['nbuf'] in (1, 2), shape
            assert halo['smem'] <= SMEM_LIMIT, (shape, halo['smem'])
            assert halo['a_stages'] >= 1 and halo['b_stages'] >= 2, shape
            assert halo['taps'] % halo['tpg'] == 0 and halo['tpg'] >= 1, shape
            assert halo['items'] >= 1, shape
            assert halo['nplanes'] == (4 if s == 2 else 1), shape
            if halo['pair']:
                assert halo['bn'] >= 32 and (halo['bn'] // 2) % 8 == 0, shape          # each CTA stages bn/2 weight rows
            # every output pixel is covered: patches * patch area >= output area
            Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
            per_item = tiles * 128 * (2 if halo['pair'] else 1)
            slices = Cout // halo['bn']
            assert halo['items'] * per_item >= N * Ho * Wo * slices, shape
        if wg is not None:
            n_wg += 1
            assert wg['ncols'] <= 512 and wg['ncols'] % 32 == 0, shape                 # TMEM accumulator columns
            assert wg['stages'] >= 2 and wg['smem'] <= SMEM_LIMIT, (shape, wg['smem'])
            assert wg['splits'] >= 1 and wg['groups'] >= 1, shape
            assert wg['splits'] * wg['tiles_per_cta'] >= wg['total_tiles'], shape      # the pixel splits cover every K tile
            assert (wg['splits'] - 1) * wg['tiles_per_cta'] < wg['total_tiles'], shape  # and none of them is empty
            assert wg['workspace'] == wg['problems'] * wg['splits'] * wg['groups'] * 128 * wg['ncols'] * 4, shape
            assert wg['problems'] == (4 if s == 2 else 1), shape
            # N box (unified over the problems of a batch): at least this problem's halo, at most the filter's
            assert wg['hxw'] <= 8 + k - 1 and wg['tr'] + wg['rg'] - 1 <= wg['hyw'] <= wg['tr'] + k - 1, shape
    assert n_halo > 300 and n_wg > 300        # the sweep really exercises the planners


def test_unsupported_shapes_fall_through():
    halo, wg = split_sections(1, 63, 65, 32, 0, 32, 5, 2)       # odd input with stride 2: per-tap kernel / FFMA wgrad
    assert halo is None and wg is None
    halo, wg = split_sections(1, 32, 32, 48, 0, 32, 3, 1)       # C0 not a multiple of 32
    assert halo is None and wg is None
