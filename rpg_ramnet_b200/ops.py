"""Tensor-level wrappers over the C ABI (pointers + sizes in, nothing allocated by the library).

Activations are fp32 "pixel-major" tensors: logical shape [N, C, H, W] with channels_last
strides, i.e. the memory is NHWC as include/ramnet_b200.h specifies.  PyTorch is used for device
memory and streams only.
"""
import contextlib
import ctypes
import os
from typing import Optional

import torch

from . import _lib
from ._lib import (ConvDesc, EPI_BIAS, EPI_BIAS_ADD, EPI_BIAS_RELU, EPI_BIAS_RELU_ADD, EPI_BIAS_RELU_PRED, EPI_BIAS_RES_RELU, EPI_GRU_OUT,  # noqa
                   EPI_GRU_RU, EPI_LSTM, FLAG_DYNAMIC, FLAG_HPACK, FLAG_ROUND_TF32, FLAG_S2SEG, FLAG_SM_TIME, FLAG_UPCONV, MMA_FP32, MMA_TF32, check)


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _h(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.RamnetError('rpg_ramnet_b200 ops need CUDA tensors (sm_100a); there is no CPU path')
    return _lib.handle(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _p(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def empty_nhwc(N, C, H, W, device) -> torch.Tensor:
    """fp32 activation: logical [N,C,H,W], NHWC memory."""
    return torch.empty((N, H, W, C), dtype=torch.float32, device=device).permute(0, 3, 1, 2)


def zeros_nhwc(N, C, H, W, device) -> torch.Tensor:
    return torch.zeros((N, H, W, C), dtype=torch.float32, device=device).permute(0, 3, 1, 2)


def _is_nhwc(t: torch.Tensor) -> bool:
    """NHWC memory under a logical [N,C,H,W] shape; strides of size-1 dims are irrelevant."""
    N, C, H, W = t.shape
    want = (H * W * C, 1, W * C, C)
    return all(sz == 1 or st == w for sz, st, w in zip(t.shape, t.stride(), want))


def as_nhwc(t: torch.Tensor) -> torch.Tensor:
    """Accept a foreign [N,C,H,W] tensor (e.g. a state handed back by the caller) and make sure its
    memory is NHWC fp32.  No copy when it already is."""
    if t.dtype != torch.float32:
        t = t.float()
    N, C, H, W = t.shape
    if _is_nhwc(t):
        return t
    out = empty_nhwc(N, C, H, W, t.device)
    if t.is_contiguous():
        lib = _lib.load()
        check(lib.ramnet_nchw_to_nhwc(_h(t), _p(t), _p(out), N, C, H, W, 0, _stream(t)))
    else:
        out.copy_(t)
    return out


def to_nchw_contiguous(t: torch.Tensor) -> torch.Tensor:
    N, C, H, W = t.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=t.device)
    check(_lib.load().ramnet_nhwc_to_nchw(_h(t), _p(t), _p(out), N, C, H, W, _stream(t)))
    return out


def _check_nhwc(t: torch.Tensor, name: str):
    if t.dim() != 4 or t.dtype != torch.float32 or not _is_nhwc(t):
        raise _lib.RamnetError(f'{name}: expected fp32 NHWC-strided [N,C,H,W] tensor, got {t.dtype} '
                               f'shape {tuple(t.shape)} strides {t.stride()}')


def voxel_grid(events: torch.Tensor, num_bins: int, width: int, height: int,
               oob_count: Optional[torch.Tensor] = None) -> torch.Tensor:
    """events: [n,4] float64 CUDA rows [t,x,y,p] -> [num_bins, height, width] fp32."""
    if events.dtype != torch.float64 or events.dim() != 2 or events.shape[1] != 4 or not events.is_contiguous():
        raise _lib.RamnetError('voxel_grid: events must be a contiguous [n,4] float64 tensor')
    grid = torch.empty((num_bins, height, width), dtype=torch.float32, device=events.device)
    check(_lib.load().ramnet_voxel_grid(_h(events), _p(events), events.shape[0], num_bins, width, height, _p(grid),
                                        _p(oob_count), _stream(events)))
    return grid


def voxel_grid_ex(events: torch.Tensor, num_bins: int, width: int, height: int, normalize: bool = False,
                  want_stats: bool = False):
    """events -> [num_bins, height, width] fp32 through ramnet_voxel_grid_ex: packed accumulation for large event
    counts, optional fused statistics of the non-zero voxels / in-place normalisation (event_dataset.py:144-151).
    Returns grid or (grid, stats[3] float64: sum, sum of squares, count of non-zero voxels)."""
    if events.dtype != torch.float64 or events.dim() != 2 or events.shape[1] != 4 or not events.is_contiguous():
        raise _lib.RamnetError('voxel_grid: events must be a contiguous [n,4] float64 tensor')
    dev = events.device
    lib = _lib.load()
    grid = torch.empty((num_bins, height, width), dtype=torch.float32, device=dev)
    nws = lib.ramnet_voxel_grid_workspace_bytes(num_bins, width, height)
    ws = _workspace(dev, nws) if nws and events.shape[0] >= 1000000 else None
    stats = torch.empty(3, dtype=torch.float64, device=dev) if (normalize or want_stats) else None
    check(lib.ramnet_voxel_grid_ex(_h(events), _p(events), events.shape[0], num_bins, width, height, _p(grid), None, _p(ws),
                                   nws if ws is not None else 0, _p(stats), 1 if normalize else 0, _stream(events)))
    return (grid, stats) if want_stats else grid


def voxel_votes(events: torch.Tensor, num_bins: int, width: int, height: int):
    n = events.shape[0]
    dev = events.device
    il = torch.empty(n, dtype=torch.int64, device=dev)
    ir = torch.empty(n, dtype=torch.int64, device=dev)
    vl = torch.empty(n, dtype=torch.float32, device=dev)
    vr = torch.empty(n, dtype=torch.float32, device=dev)
    check(_lib.load().ramnet_voxel_votes(_h(events), _p(events), n, num_bins, width, height, _p(il), _p(vl), _p(ir),
                                         _p(vr), _stream(events)))
    return il, vl, ir, vr


def head_conv(x_nchw: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], round_tf32: bool) -> torch.Tensor:
    x = x_nchw.contiguous()
    N, Cin, H, W = x.shape
    Cout = w.shape[0]
    y = empty_nhwc(N, Cout, H, W, x.device)
    with _Prof('head_conv', 2.0 * N * H * W * Cout * Cin * 25, x.device):
        check(_lib.load().ramnet_head_conv(_h(x), _p(x), _p(w), _p(b), _p(y), N, Cin, H, W, Cout,
                                           FLAG_ROUND_TF32 if round_tf32 else 0, _stream(x)))
    return y


def voxel_normalize_(grid: torch.Tensor) -> torch.Tensor:
    """In place: (x - mean) / stddev over the non-zero voxels (event_dataset.py:144-151)."""
    if not (grid.is_cuda and grid.dtype == torch.float32 and grid.is_contiguous()):
        raise _lib.RamnetError('voxel_normalize_: contiguous float32 CUDA tensor required')
    batch = grid.shape[0] if grid.dim() == 4 else 1        # [B, bins, H, W]: every sample on its own statistics
    stats = torch.empty(3 * batch, dtype=torch.float64, device=grid.device)
    check(_lib.load().ramnet_voxel_normalize(_h(grid), _p(grid), grid.numel() // batch, batch, _p(stats), _stream(grid)))
    return grid


def depth_to_label(depth: torch.Tensor, clip_distance: float, reg_factor: float) -> torch.Tensor:
    """Metric depth -> normalised log depth in [0, 1] (dataset.py:296-305); NaN pixels stay NaN."""
    d = depth.contiguous().float()
    if not d.is_cuda:
        raise _lib.RamnetError('depth_to_label: CUDA tensor required')
    out = torch.empty_like(d)
    check(_lib.load().ramnet_depth_to_label(_h(d), _p(d), _p(out), d.numel(), float(clip_distance), float(reg_factor),
                                            _stream(d)))
    return out


def depth_metric_sums(pred: torch.Tensor, target: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """[N, 8] float64 masked error sums of a [N, 1, H, W] prediction / target pair (see ramnet_depth_metrics)."""
    p, t = pred.detach().contiguous().float(), target.detach().contiguous().float()
    if not (p.is_cuda and t.is_cuda) or p.shape != t.shape or p.dim() != 4 or p.shape[1] != 1:
        raise _lib.RamnetError('depth_metric_sums: two [N, 1, H, W] CUDA tensors required')
    N, hw = p.shape[0], p.shape[2] * p.shape[3]
    out = torch.empty((N, 8), dtype=torch.float64, device=p.device)
    check(_lib.load().ramnet_depth_metrics(_h(p), _p(p), _p(t), N, hw, float(eps), _p(out), _stream(p)))
    return out


def head_tc_ok(Cin: int, Cout: int) -> bool:
    """The tensor-core head path covers 5*Cin <= 32 (Cin <= 6: every shipped configuration) and Cout % 32 == 0."""
    return 5 * Cin <= 32 and Cout % 32 == 0 and os.environ.get('RAMNET_HEAD_TC', '1') != '0'


def pack_weights_head(w_oihw: torch.Tensor) -> torch.Tensor:
    w = w_oihw.detach().contiguous().float()
    Cout, Cin = w.shape[0], w.shape[1]
    out = torch.empty(5 * Cout * 32, dtype=torch.float32, device=w.device)
    check(_lib.load().ramnet_pack_weights_head(_h(w), _p(w), _p(out), Cout, Cin, _stream(w)))
    return out


def head_im2row(x_nchw: torch.Tensor) -> torch.Tensor:
    """[N, Cin, H, W] -> NHWC [N, 32, H, W]: channel dx*Cin + ci holds x[ci] shifted by dx - 2 (TF32-rounded)."""
    x = x_nchw.contiguous()
    N, Cin, H, W = x.shape
    xe = empty_nhwc(N, 32, H, W, x.device)
    with _Prof('head_conv', 0.0, x.device):
        check(_lib.load().ramnet_head_im2row(_h(x), _p(x), _p(xe), N, Cin, H, W, _stream(x)))
    return xe


def head_conv_tc(xe: torch.Tensor, w_packed: torch.Tensor, b: Optional[torch.Tensor], Cin: int, Cout: int,
                 round_tf32: bool) -> torch.Tensor:
    _check_nhwc(xe, 'head_conv_tc xe')
    N, _, H, W = xe.shape
    y = empty_nhwc(N, Cout, H, W, xe.device)
    with _Prof('head_conv', 2.0 * N * H * W * Cout * Cin * 25, xe.device):
        check(_lib.load().ramnet_head_conv_tc(_h(xe), _p(xe), _p(w_packed), _p(b), _p(y), N, H, W, Cout,
                                              FLAG_ROUND_TF32 if round_tf32 else 0, _stream(xe)))
    return y


def hpack_eligible(Cout: int, ksize: int, stride: int, mma_kind: int) -> bool:
    """Layers the horizontal-tap-packed forward (RAMNET_FLAG_HPACK) is meant for: few output channels at stride 1, where a
    128 x 32 x 8 MMA per tap is bound by the A-operand shared-memory reads.  Default on since round 2 (validated on
    hardware: tests/test_gpu_experimental.py, profiles/r02_first_call_experimental_paths.txt); RAMNET_HPACK=0 disables."""
    if os.environ.get('RAMNET_HPACK', '1') == '0' or mma_kind != MMA_TF32 or stride != 1 or ksize not in (3, 5):
        return False
    # Cout > 32 runs as Cout / 32 column slices (the A operand is re-read once per slice): RAMNET_HPACK_MAXC bounds it
    return Cout % 32 == 0 and Cout <= int(os.environ.get('RAMNET_HPACK_MAXC', '32'))


def pack_weights_hpack(w_oihw: torch.Tensor) -> torch.Tensor:
    w = w_oihw.detach().contiguous().float()
    Cout, Cin, k, _ = w.shape
    out = torch.empty(w.numel(), dtype=torch.float32, device=w.device)
    check(_lib.load().ramnet_pack_weights_hpack(_h(w), _p(w), _p(out), Cout, Cin, k, _stream(w)))
    out._ramnet_hpack = True          # conv_fwd sets RAMNET_FLAG_HPACK for weights packed this way
    return out


# Planning hints OR-ed into every conv launch descriptor (engine.GraphRunner sets FLAG_SM_TIME while it captures passes
# that will overlap on two streams).
_PLAN_FLAGS = [FLAG_DYNAMIC if os.environ.get('RAMNET_FORCE_DYNAMIC') == '1' else 0]      # RAMNET_FORCE_DYNAMIC: A/B aid (every TF32 conv launch)


@contextlib.contextmanager
def plan_flags(flags: int):
    old = _PLAN_FLAGS[0]
    _PLAN_FLAGS[0] = old | flags
    try:
        yield
    finally:
        _PLAN_FLAGS[0] = old


def s2seg_eligible(Cin: int, Cout: int, ksize: int, stride: int, mma_kind: int) -> bool:
    """5x5 stride-2 single-source layers (the encoders' strided convs) run as four parity-plane K segments
    (RAMNET_FLAG_S2SEG): one plane per pipeline stage instead of all four in one.  RAMNET_S2SEG=0 disables."""
    if os.environ.get('RAMNET_S2SEG', '1') == '0':
        return False
    return mma_kind == MMA_TF32 and stride == 2 and ksize == 5 and Cin % 32 == 0 and Cout % 16 == 0


def pack_weights_s2seg(w_oihw: torch.Tensor) -> torch.Tensor:
    w = w_oihw.detach().contiguous().float()
    Cout, Cin, k, _ = w.shape
    assert k == 5
    out = torch.empty(27 * Cout * Cin, dtype=torch.float32, device=w.device)
    check(_lib.load().ramnet_pack_weights_s2seg(_h(w), _p(w), _p(out), Cout, Cin, _stream(w)))
    out._ramnet_s2seg = True          # conv_fwd sets RAMNET_FLAG_S2SEG for weights packed this way
    return out


def upconv_eligible(Cin: int, Cout: int, ksize: int, mma_kind: int) -> bool:
    """Layers ramnet_conv_fwd can run in up-conv mode (bilinear x2 + 5x5 conv in one launch on the low-resolution input,
    RAMNET_FLAG_UPCONV): TF32, 5x5, Cin % 32 == 0, 4 * Cout GEMM columns <= 512.  RAMNET_UPCONV=0 disables;
    RAMNET_UPCONV_MAXC bounds Cout (default 64: the two narrow decoders, where N = 4 * Cout turns 32 / 64-column MMAs
    into 128 / 256-column ones; the 128-channel decoder gains nothing from wider MMAs and pays the border segments)."""
    if os.environ.get('RAMNET_UPCONV', '1') == '0' or mma_kind != MMA_TF32 or ksize != 5:
        return False
    return Cin % 32 == 0 and Cout % 16 == 0 and Cout <= int(os.environ.get('RAMNET_UPCONV_MAXC', '64'))


def pack_weights_upconv(w_oihw: torch.Tensor) -> torch.Tensor:
    w = w_oihw.detach().contiguous().float()
    Cout, Cin, k, _ = w.shape
    if k != 5:
        raise _lib.RamnetError('pack_weights_upconv: 5x5 filters only')
    lib = _lib.load()
    out = torch.empty(int(lib.ramnet_upconv_packed_floats(Cout, Cin)), dtype=torch.float32, device=w.device)
    check(lib.ramnet_pack_weights_upconv(_h(w), _p(w), _p(out), Cout, Cin, _stream(w)))
    out._ramnet_upconv = True
    return out


def conv_up_fwd(x: torch.Tensor, w_packed_up: torch.Tensor, bias: Optional[torch.Tensor], Cout: int, epilogue: int,
                aux0=None, aux1=None, round_tf32: bool = False, out1=None):
    """UpsampleConvLayer.forward (submodules.py:87-97) without the 4x tensor: y = epi(conv5x5(bilinear_up2x(x)) + b) from
    the LOW-resolution x [N, Cin, H, W] -> [N, Cout, 2H, 2W] (EPI_BIAS_RELU_PRED: depth [N, 1, 2H, 2W]).
    EPI_BIAS_RELU_ADD adds aux0 ([N, Cout, 2H, 2W], the next decoder's skip state) after the ReLU."""
    _check_nhwc(x, 'conv_up_fwd x')
    N, Cin, H, W = x.shape
    dev = x.device
    if epilogue == EPI_BIAS_RELU_PRED:
        y0 = torch.empty((N, 1, 2 * H, 2 * W), dtype=torch.float32, device=dev)
    else:
        y0 = empty_nhwc(N, Cout, 2 * H, 2 * W, dev)
        if aux0 is not None:
            _check_nhwc(aux0, 'conv_up_fwd aux0')
            if tuple(aux0.shape) != (N, Cout, 2 * H, 2 * W):
                raise _lib.RamnetError(f'conv_up_fwd: aux0 shape {tuple(aux0.shape)} != {(N, Cout, 2 * H, 2 * W)}')
    flags = FLAG_UPCONV | (FLAG_ROUND_TF32 if round_tf32 else 0) | _PLAN_FLAGS[0]
    slot = _sched_slot(dev) if (flags & FLAG_DYNAMIC) else None
    if slot is None:
        flags &= ~FLAG_DYNAMIC
    d = ConvDesc(N, H, W, Cin, 0, Cout, 5, 1, epilogue, MMA_TF32, flags, 0)
    with _Prof('conv', 2.0 * N * (2 * H) * (2 * W) * Cout * Cin * 25, dev,      # algorithmic FLOPs of the reference graph
               tag=PROFILE is not None and f'upconv {H}x{W} {Cin}->{Cout} e{epilogue}'):
        check(_lib.load().ramnet_conv_fwd(_h(x), ctypes.byref(d), _p(x), None, _p(w_packed_up), _p(bias), _p(aux0), _p(aux1),
                                          _p(y0), _p(out1), None, _p(slot), 8 if slot is not None else 0, _stream(x)))
    return y0


def pack_weights(w_oihw: torch.Tensor, mma_kind: int, lstm_interleave: bool = False) -> torch.Tensor:
    w = w_oihw.detach().contiguous().float()
    Cout, Cin, k, _ = w.shape
    out = torch.empty(w.numel(), dtype=torch.float32, device=w.device)
    check(_lib.load().ramnet_pack_weights(_h(w), _p(w), _p(out), Cout, Cin, k, mma_kind, int(lstm_interleave),
                                          _stream(w)))
    return out


# bench.py hook: when a list, every wrapped launch appends (kind, algorithmic_flops, start_event, end_event)
PROFILE = None


class _Prof:
    def __init__(self, kind, flops, dev, tag=None):
        self.kind, self.flops, self.dev, self.tag = kind, flops, dev, tag

    def __enter__(self):
        if PROFILE is not None:
            self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.dev))

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.b.record(torch.cuda.current_stream(self.dev))
            PROFILE.append((self.kind, self.flops, self.a, self.b) if self.tag is None else
                           (self.kind, self.flops, self.a, self.b, self.tag))


_workspaces = {}      # (device, stream) -> [buffers, newest last]
_sched_pools = {}     # device -> [zeroed int32 pool, next free pair]
SCHED_SLOT_OVERRIDE = None


def _sched_slot(device):
    """8 zeroed bytes for ONE launch site of a RAMNET_FLAG_DYNAMIC convolution (item counter + finished-worker count;
    the kernel resets them when it ends).  Slots are never reused: a captured CUDA graph bakes the address in and
    replays it for the life of the process.  None when the pool is exhausted (the launch then runs with the static
    assignment)."""
    if SCHED_SLOT_OVERRIDE is not None:        # tests: one slot for many launches exercises the kernel's self-reset
        return SCHED_SLOT_OVERRIDE
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    pool = _sched_pools.get(key)
    if pool is None:
        pool = _sched_pools[key] = [torch.zeros(1 << 16, dtype=torch.int32, device=device), 0]
    if 2 * pool[1] + 2 > pool[0].numel():
        return None
    slot = pool[0][2 * pool[1]:2 * pool[1] + 2]
    pool[1] += 1
    return slot


def _workspace(device, nbytes: int):
    """Scratch for the kernels that need one (wgrad partial tiles), one growable buffer per (device, stream): kernels
    on one stream are ordered, so they can share it; different streams never do.  A buffer that is outgrown is KEPT
    alive for the life of the process — a captured CUDA graph may have baked its address in, and handing it back to
    the caching allocator would let graph replays scribble over other tensors (growth is geometric, so the retired
    buffers sum to less than the live one)."""
    if nbytes == 0:
        return None
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    bufs = _workspaces.setdefault(key, [])
    if not bufs or bufs[-1].numel() < nbytes:
        grow = max(nbytes, 2 * bufs[-1].numel() if bufs else 0)
        bufs.append(torch.empty(grow, dtype=torch.uint8, device=device))
    return bufs[-1]


def conv_fwd(x0: torch.Tensor, x1: Optional[torch.Tensor], w_packed: torch.Tensor, bias: Optional[torch.Tensor],
             Cout: int, ksize: int, stride: int, epilogue: int, mma_kind: int, aux0=None, aux1=None,
             round_tf32: bool = False, out0=None, out1=None, stash=None):
    """Implicit-GEMM convolution with fused epilogue.  Returns y0 or (y0, y1).
    `out0` / `out1`: optional preallocated NHWC outputs (persistent state buffers of the CUDA-graph runner)."""
    _check_nhwc(x0, 'conv_fwd x0')
    N, C0, H, W = x0.shape
    C1 = 0
    if x1 is not None:
        _check_nhwc(x1, 'conv_fwd x1')
        C1 = x1.shape[1]
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    dev = x0.device
    y1 = None
    def _out(given, C):
        if given is None:
            return empty_nhwc(N, C, Ho, Wo, dev)
        _check_nhwc(given, 'conv_fwd out')
        if tuple(given.shape) != (N, C, Ho, Wo):
            raise _lib.RamnetError(f'conv_fwd: out shape {tuple(given.shape)} != {(N, C, Ho, Wo)}')
        return given

    if epilogue == EPI_GRU_RU:
        y0, y1 = _out(out0, Cout // 2), _out(out1, Cout // 2)
    elif epilogue == EPI_LSTM:
        y0, y1 = _out(out0, Cout // 4), _out(out1, Cout // 4)
    elif epilogue == EPI_BIAS_RELU_PRED:      # depth [N,1,H,W] (+ logits when out1 is given)
        y0 = torch.empty((N, 1, Ho, Wo), dtype=torch.float32, device=dev) if out0 is None else out0
        y1 = out1
    else:
        y0 = _out(out0, Cout)
    if epilogue != EPI_BIAS_RELU_PRED:
        for a, nm in ((aux0, 'aux0'), (aux1, 'aux1')):
            if a is not None:
                _check_nhwc(a, 'conv_fwd ' + nm)
    flags = (FLAG_ROUND_TF32 if round_tf32 else 0) | (FLAG_HPACK if getattr(w_packed, '_ramnet_hpack', False) else 0)
    if getattr(w_packed, '_ramnet_s2seg', False):
        flags |= FLAG_S2SEG
    flags |= _PLAN_FLAGS[0]
    ws, nws = None, 0
    if (flags & FLAG_DYNAMIC) and mma_kind == MMA_TF32:
        ws = _sched_slot(dev)
        nws = 8 if ws is not None else 0
    if ws is None:
        flags &= ~FLAG_DYNAMIC
    d = ConvDesc(N, H, W, C0, C1, Cout, ksize, stride, epilogue, mma_kind, flags, 0)
    lib = _lib.load()
    if ws is None:
        nws = lib.ramnet_conv_workspace_bytes(ctypes.byref(d))
        ws = _workspace(dev, nws)
    with _Prof('conv', 2.0 * N * Ho * Wo * Cout * (C0 + C1) * ksize * ksize, dev,
               tag=PROFILE is not None and f'conv {H}x{W} {C0}+{C1}->{Cout} k{ksize} s{stride} e{epilogue}'):
        check(lib.ramnet_conv_fwd(_h(x0), ctypes.byref(d), _p(x0), _p(x1), _p(w_packed), _p(bias), _p(aux0),
                                  _p(aux1), _p(y0), _p(y1), _p(stash), _p(ws), nws, _stream(x0)))
    return (y0, y1) if (y1 is not None and epilogue != EPI_BIAS_RELU_PRED) else y0


def upsample2x_add(x: torch.Tensor, skip: Optional[torch.Tensor], round_tf32: bool) -> torch.Tensor:
    _check_nhwc(x, 'upsample2x_add x')
    if skip is not None:
        _check_nhwc(skip, 'upsample2x_add skip')
        if skip.shape != x.shape:
            raise _lib.RamnetError(f'upsample2x_add: skip shape {tuple(skip.shape)} != x shape {tuple(x.shape)}')
    N, C, H, W = x.shape
    y = empty_nhwc(N, C, 2 * H, 2 * W, x.device)
    with _Prof('upsample2x_add', 0.0, x.device):
        check(_lib.load().ramnet_upsample2x_add(_h(x), _p(x), _p(skip), _p(y), N, H, W, C,
                                                FLAG_ROUND_TF32 if round_tf32 else 0, _stream(x)))
    return y


def _pred_weights(w, C, concat):
    """-> (w [C], w_skip [C] or None): skip_type 'concat' splits the [1, 2C, 1, 1] weight into the x and the skip half."""
    wv = w.detach().reshape(-1).contiguous().float()
    if not concat:
        return wv, None
    if wv.numel() != 2 * C:
        raise _lib.RamnetError(f'pred (concat): weight has {wv.numel()} input channels, expected {2 * C}')
    return wv[:C].contiguous(), wv[C:].contiguous()


def pred_sigmoid(x: torch.Tensor, skip: Optional[torch.Tensor], w: torch.Tensor, b: Optional[torch.Tensor],
                 want_logits: bool = False, concat: bool = False):
    """1x1 conv to one channel + sigmoid. Returns depth [N,1,H,W] (and logits).  concat: the conv runs over
    cat([x, skip]) (unet.py:11-12) without materialising it."""
    _check_nhwc(x, 'pred_sigmoid x')
    if skip is not None:
        _check_nhwc(skip, 'pred_sigmoid skip')
    N, C, H, W = x.shape
    depth = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
    logits = torch.empty_like(depth) if want_logits else None
    wv, ws = _pred_weights(w, C, concat)
    with _Prof('pred_sigmoid', 2.0 * N * H * W * C, x.device):
        check(_lib.load().ramnet_pred_sigmoid(_h(x), _p(x), _p(skip), _p(wv), _p(ws), _p(b), _p(logits), _p(depth),
                                              N * H * W, C, _stream(x)))
    return (depth, logits) if want_logits else depth


def si_loss_stats(pred: torch.Tensor, target: torch.Tensor, out: Optional[torch.Tensor] = None,
                  log_space: bool = False) -> torch.Tensor:
    """(sum d, sum d^2, n) in float64; `out`: optional [3] slice of a larger statistics buffer (batched exchange)."""
    pred, target = pred.contiguous(), target.contiguous()
    stats = torch.empty(3, dtype=torch.float64, device=pred.device) if out is None else out
    check(_lib.load().ramnet_si_loss_stats(_h(pred), _p(pred), _p(target), pred.numel(), _p(stats),
                                           _lib.LOSS_LOG_SPACE if log_space else 0, _stream(pred)))
    return stats


def si_loss_value(stats: torch.Tensor, weight: float, n_lambda: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty((), dtype=torch.float32, device=stats.device)
    check(_lib.load().ramnet_si_loss_value(_h(stats), _p(stats), weight, n_lambda, _p(out), _stream(stats)))
    return out


def si_loss_grad(pred, target, stats, weight: float, n_lambda: float, scale: float = 1.0,
                 scale_dev: Optional[torch.Tensor] = None, log_space: bool = False) -> torch.Tensor:
    """`scale_dev`: optional float32 device scalar (autograd's grad_output) multiplied in by the kernel."""
    pred, target = pred.contiguous(), target.contiguous()
    grad = torch.empty_like(pred)
    if scale_dev is not None and (scale_dev.dtype != torch.float32 or not scale_dev.is_cuda):
        scale_dev = scale_dev.to(device=pred.device, dtype=torch.float32)
    check(_lib.load().ramnet_si_loss_grad(_h(pred), _p(pred), _p(target), pred.numel(), _p(stats), weight, n_lambda,
                                          scale, _p(scale_dev), _lib.LOSS_LOG_SPACE if log_space else 0, _p(grad),
                                          _stream(pred)))
    return grad


def adam_step(p, g, m, v, step: int, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """In-place fused Adam on flat fp32 buffers."""
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.RamnetError('adam_step: flat contiguous fp32 buffers required')
    check(_lib.load().ramnet_adam_step(_h(p), _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps,
                                       weight_decay, step, _stream(p)))


# ------------------------------------------------------------------------------------------------
# backward building blocks (include/ramnet_b200.h, "a-13")
# ------------------------------------------------------------------------------------------------
def pack_weights_dgrad(w_oihw: torch.Tensor, mma_kind: int, ci_begin: int, ci_count: int) -> torch.Tensor:
    w = w_oihw.detach().contiguous().float()
    Cout, Cin, k, _ = w.shape
    out = torch.empty(Cout * ci_count * k * k, dtype=torch.float32, device=w.device)
    check(_lib.load().ramnet_pack_weights_dgrad(_h(w), _p(w), _p(out), Cout, Cin, k, mma_kind, ci_begin, ci_count,
                                                _stream(w)))
    return out


def pack_weights_dgrad_s2(w_oihw: torch.Tensor, ci_begin: int, ci_count: int) -> torch.Tensor:
    """Sub-filters of the four input parities of a stride-2 conv, for conv_dgrad_s2 (TF32 path)."""
    w = w_oihw.detach().contiguous().float()
    Cout, Cin, k, _ = w.shape
    out = torch.empty(Cout * ci_count * k * k, dtype=torch.float32, device=w.device)
    check(_lib.load().ramnet_pack_weights_dgrad_s2(_h(w), _p(w), _p(out), Cout, Cin, k, ci_begin, ci_count, _stream(w)))
    return out


def conv_dgrad_s2(dz, w_packed_s2, ci_count, ksize, H, W):
    """Data gradient [N, ci_count, H, W] of a stride-2 conv from dz [N, Cout, H/2, W/2]: four sub-pixel convolutions."""
    _check_nhwc(dz, 'conv_dgrad_s2 dz')
    N, Cout, Ho, Wo = dz.shape
    dx = empty_nhwc(N, ci_count, H, W, dz.device)
    with _Prof('conv', 2.0 * N * Ho * Wo * Cout * ci_count * ksize * ksize, dz.device,
               tag=PROFILE is not None and f'dgrad_s2 {H}x{W} {Cout}->{ci_count} k{ksize}'):
        check(_lib.load().ramnet_conv_dgrad_s2(_h(dz), _p(dz), _p(w_packed_s2), _p(dx), N, H, W, Cout, ci_count, ksize, 0,
                                               _stream(dz)))
    return dx


def conv_wgrad(dz, x0, x1, Cout, ksize, stride, dw, db, mma_kind=MMA_FP32):
    """dw [Cout, C0+C1, k, k] += , db [Cout] += ."""
    _check_nhwc(dz, 'conv_wgrad dz')
    _check_nhwc(x0, 'conv_wgrad x0')
    N, C0, H, W = x0.shape
    C1 = 0 if x1 is None else x1.shape[1]
    d = ConvDesc(N, H, W, C0, C1, Cout, ksize, stride, 0, mma_kind, 0, 0)
    with _Prof('wgrad', 2.0 * dz.shape[0] * dz.shape[2] * dz.shape[3] * Cout * (C0 + C1) * ksize * ksize, x0.device,
               tag=PROFILE is not None and f'wgrad {H}x{W} {C0}+{C1}->{Cout} k{ksize} s{stride}'):
        lib = _lib.load()
        nws = lib.ramnet_conv_wgrad_workspace_bytes(_h(x0), ctypes.byref(d))
        ws = _workspace(x0.device, nws)
        check(lib.ramnet_conv_wgrad(_h(x0), ctypes.byref(d), _p(dz), _p(x0), _p(x1), _p(dw), _p(db), _p(ws), nws,
                                    _lib.WGRAD_FULL, _stream(x0)))


class WgradAccumulator:
    """Weight gradient of ONE layer accumulated over the passes of a training step (BPTT calls the layer's backward once
    per pass): `add()` runs only the tensor-core kernel, whose epilogue writes (first call) or adds (later calls) its
    partial tiles into a workspace this object owns; `finalize(dw)` runs the split sum + scatter ONCE and accumulates
    into dw [Cout, C0+C1, k, k].  Built by `wgrad_accumulator()`, which returns None for shapes the tap-packed TF32
    kernel does not cover (callers then use conv_wgrad per call)."""

    def __init__(self, desc, nws, device, head=None):
        self.desc, self.nws, self.head = desc, nws, head
        self.ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=device)
        self.dirty = False
        self.flops = 0.0

    def add(self, dz, x0, x1=None):
        lib, d = _lib.load(), self.desc
        mode = _lib.WGRAD_PARTIAL_ADD if self.dirty else _lib.WGRAD_PARTIAL_FIRST
        with _Prof('wgrad', self.flops, x0.device, tag=PROFILE is not None and f'wgrad-acc {d.H}x{d.W} {d.C0}+{d.C1}->{d.Cout}'):
            if self.head is not None:
                N, Cin, H, W, Cout = self.head
                check(lib.ramnet_head_conv_wgrad_tc(_h(x0), _p(x0), _p(dz), None, None, N, Cin, H, W, Cout, _p(self.ws),
                                                    self.nws, mode, _stream(x0)))
            else:
                check(lib.ramnet_conv_wgrad(_h(x0), ctypes.byref(d), _p(dz), _p(x0), _p(x1), None, None, _p(self.ws),
                                            self.nws, mode, _stream(x0)))
        self.dirty = True

    def finalize(self, dw):
        if not self.dirty:
            return
        lib, d = _lib.load(), self.desc
        with _Prof('wgrad_finalize', 0.0, dw.device):
            if self.head is not None:
                N, Cin, H, W, Cout = self.head
                check(lib.ramnet_head_conv_wgrad_tc(_h(dw), None, None, _p(dw), None, N, Cin, H, W, Cout, _p(self.ws),
                                                    self.nws, _lib.WGRAD_FINALIZE, _stream(dw)))
            else:
                check(lib.ramnet_conv_wgrad(_h(dw), ctypes.byref(d), None, None, None, _p(dw), None, _p(self.ws), self.nws,
                                            _lib.WGRAD_FINALIZE, _stream(dw)))
        self.dirty = False


def wgrad_accumulator(x0, x1, dz, Cout, ksize, stride, mma_kind):
    """WgradAccumulator for the layer whose backward sees (dz, x0, x1) of these shapes, or None when unsupported."""
    if mma_kind != MMA_TF32 or os.environ.get('RAMNET_WGRAD_DEFER', '1') == '0':
        return None
    N, C0, H, W = x0.shape
    C1 = 0 if x1 is None else x1.shape[1]
    d = ConvDesc(N, H, W, C0, C1, Cout, ksize, stride, 0, mma_kind, 0, 0)
    lib = _lib.load()
    nws = lib.ramnet_conv_wgrad_workspace_bytes(_h(x0), ctypes.byref(d))
    if nws == 0:
        return None
    acc = WgradAccumulator(d, nws, x0.device)
    acc.flops = 2.0 * dz.shape[0] * dz.shape[2] * dz.shape[3] * Cout * (C0 + C1) * ksize * ksize
    # probe once: the deferred modes exist for the tap-packed kernel only
    rc = lib.ramnet_conv_wgrad(_h(x0), ctypes.byref(d), _p(dz), _p(x0), _p(x1), None, None, _p(acc.ws), nws,
                               _lib.WGRAD_PARTIAL_FIRST, _stream(x0))
    if rc == _lib.RAMNET_EUNSUPPORTED:
        return None
    check(rc)
    acc.dirty = True
    return acc


def head_wgrad_accumulator(xe, dz, Cin):
    if os.environ.get('RAMNET_WGRAD_DEFER', '1') == '0':
        return None
    N, _, H, W = xe.shape
    Cout = dz.shape[1]
    lib = _lib.load()
    nws = lib.ramnet_head_conv_wgrad_tc_workspace_bytes(_h(xe), N, Cin, H, W, Cout)
    if nws == 0:
        return None
    acc = WgradAccumulator(ConvDesc(N, H, W, 32, 0, Cout, 5, 1, 0, MMA_TF32, 0, 0), nws, xe.device, head=(N, Cin, H, W, Cout))
    acc.flops = 2.0 * N * H * W * Cout * Cin * 25
    rc = lib.ramnet_head_conv_wgrad_tc(_h(xe), _p(xe), _p(dz), None, None, N, Cin, H, W, Cout, _p(acc.ws), nws,
                                       _lib.WGRAD_PARTIAL_FIRST, _stream(xe))
    if rc == _lib.RAMNET_EUNSUPPORTED:
        return None
    check(rc)
    acc.dirty = True
    return acc


def head_conv_wgrad(x_nchw, dz, dw, db):
    x = x_nchw.contiguous()
    N, Cin, H, W = x.shape
    _check_nhwc(dz, 'head_conv_wgrad dz')
    with _Prof('head_wgrad', 2.0 * N * H * W * dz.shape[1] * Cin * 25, x.device):
        check(_lib.load().ramnet_head_conv_wgrad(_h(x), _p(x), _p(dz), _p(dw), _p(db), N, Cin, H, W, dz.shape[1], _stream(x)))


def head_conv_wgrad_tc(xe, dz, dw, db, Cin):
    """Head conv weight gradient from the unrolled input (head_im2row): dw [Cout, Cin, 5, 5] +=, db [Cout] +=."""
    _check_nhwc(xe, 'head_conv_wgrad_tc xe')
    _check_nhwc(dz, 'head_conv_wgrad_tc dz')
    N, _, H, W = xe.shape
    Cout = dz.shape[1]
    with _Prof('head_wgrad', 2.0 * N * H * W * Cout * Cin * 25, xe.device):
        lib = _lib.load()
        nws = lib.ramnet_head_conv_wgrad_tc_workspace_bytes(_h(xe), N, Cin, H, W, Cout)
        ws = _workspace(xe.device, nws)
        check(lib.ramnet_head_conv_wgrad_tc(_h(xe), _p(xe), _p(dz), _p(dw), _p(db), N, Cin, H, W, Cout, _p(ws), nws,
                                            _lib.WGRAD_FULL, _stream(xe)))


def zero_insert2x(x, Hout, Wout, skip=None, want_sum=False):
    """y[n, 2h, 2w] = x[n, h, w] (+ skip), zeros elsewhere; `want_sum`: also return the dense x + skip."""
    _check_nhwc(x, 'zero_insert2x x')
    N, C, H, W = x.shape
    y = empty_nhwc(N, C, Hout, Wout, x.device)
    if skip is not None:
        _check_nhwc(skip, 'zero_insert2x skip')
    xs = empty_nhwc(N, C, H, W, x.device) if want_sum else None
    check(_lib.load().ramnet_zero_insert2x(_h(x), _p(x), _p(skip), _p(y), _p(xs), N, H, W, C, Hout, Wout, _stream(x)))
    return (y, xs) if want_sum else y


_colsum_scratch = {}


def _colsum_ws(device):
    """Scratch of the atomic-free bias-gradient fold (ticket counter + per-block partial sums), one per (device, stream),
    zeroed once: the kernels return the counter to 0.  Opt-in (RAMNET_COLSUM_LASTBLOCK=1): measured SLOWER than the
    per-block atomics on B200 (relu_bwd 40 vs 27 us, training step 66.3 vs 62.9 ms, profiles/r02_colsum_ab.txt) because
    the single last block's pass over blocks x C partial sums is a serial tail; kept for its fixed summation order."""
    if os.environ.get('RAMNET_COLSUM_LASTBLOCK', '0') != '1':
        return None
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _colsum_scratch.get(key)
    if ws is None:
        ws = _colsum_scratch[key] = torch.zeros(_lib.load().ramnet_colsum_scratch_bytes() // 4, dtype=torch.float32,
                                                device=device)
    return ws


def relu_bwd(dy, y, round_tf32=False, db=None):
    """dz = dy * (y > 0); db [C] (optional) += column sums of dz (the layer's bias gradient, fused)."""
    _check_nhwc(dy, 'relu_bwd dy')
    _check_nhwc(y, 'relu_bwd y')
    dz = empty_nhwc(*y.shape, y.device)
    check(_lib.load().ramnet_relu_bwd(_h(y), _p(dy), _p(y), _p(dz), y.numel(), y.shape[1], _p(db),
                                      _p(_colsum_ws(y.device) if db is not None else None),
                                      FLAG_ROUND_TF32 if round_tf32 else 0, _stream(y)))
    return dz


def colsum_fusable(C: int) -> bool:
    """The pointwise adjoints can fold the bias gradient in when 256 % (C/4) == 0 (every shipped layer)."""
    return C % 4 == 0 and C // 4 <= 256 and 256 % (C // 4) == 0


def gru_out_bwd(dhn, h, u, o, round_tf32=False, db_o=None, db_ru=None):
    N, C, H, W = h.shape
    dzo, dh = empty_nhwc(N, C, H, W, h.device), empty_nhwc(N, C, H, W, h.device)
    dzru = empty_nhwc(N, 2 * C, H, W, h.device)
    ws = _colsum_ws(h.device) if (db_o is not None or db_ru is not None) else None
    check(_lib.load().ramnet_gru_out_bwd(_h(h), _p(dhn), _p(h), _p(u), _p(o), _p(dzo), _p(dzru), _p(dh), _p(db_o), _p(db_ru),
                                         _p(ws), N * H * W, C, FLAG_ROUND_TF32 if round_tf32 else 0, _stream(h)))
    return dzo, dzru, dh


def gru_ru_bwd(drh, h, r, dzru, dh, round_tf32=False, db_ru=None):
    N, C, H, W = h.shape
    check(_lib.load().ramnet_gru_ru_bwd(_h(h), _p(drh), _p(h), _p(r), _p(dzru), _p(dh), _p(db_ru),
                                        _p(_colsum_ws(h.device) if db_ru is not None else None), N * H * W, C,
                                        FLAG_ROUND_TF32 if round_tf32 else 0, _stream(h)))


def pred_bwd(ddepth, depth, x, w, skip=None, concat=False):
    """-> (dx, dskip, dw, db): dskip is dx itself for the summed skip, its own tensor for the concatenated one (dw then [2C])."""
    _check_nhwc(x, 'pred_bwd x')
    N, C, H, W = x.shape
    dx = empty_nhwc(N, C, H, W, x.device)
    wv, ws = _pred_weights(w, C, concat)
    dskip = empty_nhwc(N, C, H, W, x.device) if concat else None
    dw = torch.zeros(2 * C if concat else C, dtype=torch.float32, device=x.device)
    db = torch.zeros(1, dtype=torch.float32, device=x.device)
    check(_lib.load().ramnet_pred_bwd(_h(x), _p(ddepth.contiguous()), _p(None if depth is None else depth.contiguous()), _p(x),
                                      _p(skip), _p(wv), _p(ws), _p(dx), _p(dskip), _p(dw), _p(db), N * H * W, C, _stream(x)))
    return dx, (dskip if concat else (dx if skip is not None else None)), dw, db


NORM_ACT = {None: 0, 'none': 0, 'relu': _lib.NORM_RELU, 'sigmoid': _lib.NORM_SIGMOID}


def norm_fwd(z, kind: str, act, gamma=None, beta=None, res=None, running_mean=None, running_var=None, momentum=0.1,
             eps=1e-5, batch_stats=True, round_tf32=False):
    """act(norm(z) (+ res)) for an NHWC conv output z (ramnet_norm_fwd).  kind 'BN': statistics per channel over the
    batch; 'IN': per (sample, channel).  batch_stats=False normalises with the running statistics (eval-mode norm
    that gradients flow through) and updates nothing; otherwise running_mean / running_var (when given) are updated
    in place exactly as nn.BatchNorm2d / F.instance_norm do in train mode.  Returns (y, stats) -- stats [G, C, 2]
    (mean, invstd) is what norm_bwd needs."""
    _check_nhwc(z, 'norm_fwd z')
    if res is not None:
        _check_nhwc(res, 'norm_fwd res')
    N, C, H, W = z.shape
    flags = NORM_ACT[act] | (_lib.NORM_INSTANCE if kind == 'IN' else 0) | (0 if batch_stats else _lib.NORM_RUNNING) | \
        (_lib.NORM_ROUND_TF32 if round_tf32 else 0)
    G = N if (kind == 'IN' and batch_stats) else 1
    sums = torch.empty(_lib.load().ramnet_norm_scratch_bytes(N, C) // 8, dtype=torch.float64, device=z.device)
    stats = torch.empty((G, C, 2), dtype=torch.float32, device=z.device)
    y = empty_nhwc(N, C, H, W, z.device)
    for t in (gamma, beta, running_mean, running_var):
        if t is not None and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == C):
            raise _lib.RamnetError('norm_fwd: per-channel parameters must be contiguous fp32 CUDA tensors of C elements')
    check(_lib.load().ramnet_norm_fwd(_h(z), _p(z), _p(res), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
                                      float(momentum), float(eps), N, H * W, C, flags, _p(sums), _p(stats), _p(y), _stream(z)))
    return y, stats


def norm_bwd(dy, y, z, stats, kind: str, act, gamma=None, batch_stats=True, round_tf32=False, want_dres=False,
             dgamma=None, dbeta=None):
    """Adjoint of norm_fwd: returns (dz, dres); dgamma / dbeta ([C], optional) are accumulated into."""
    _check_nhwc(dy, 'norm_bwd dy')
    N, C, H, W = z.shape
    flags = NORM_ACT[act] | (_lib.NORM_INSTANCE if kind == 'IN' else 0) | (0 if batch_stats else _lib.NORM_RUNNING) | \
        (_lib.NORM_ROUND_TF32 if round_tf32 else 0)
    G = stats.shape[0]
    sums = torch.empty(_lib.load().ramnet_norm_scratch_bytes(N, C) // 8, dtype=torch.float64, device=z.device)
    coef = torch.empty((G, C, 2), dtype=torch.float32, device=z.device)
    dz = empty_nhwc(N, C, H, W, z.device)
    dres = empty_nhwc(N, C, H, W, z.device) if want_dres else None
    check(_lib.load().ramnet_norm_bwd(_h(z), _p(dy), _p(y if NORM_ACT[act] else None), _p(z), _p(stats), _p(gamma), N, H * W, C,
                                      flags, _p(sums), _p(coef), _p(dz), _p(dres), _p(dgamma), _p(dbeta), _stream(z)))
    return dz, dres


def pred_logits(x, skip, w, b, concat=False):
    """1x1 conv to one channel, no activation (a norm layer follows): [N,1,H,W]."""
    _check_nhwc(x, 'pred_logits x')
    N, C, H, W = x.shape
    logits = torch.empty((N, 1, H, W), dtype=torch.float32, device=x.device)
    wv, ws = _pred_weights(w, C, concat)
    check(_lib.load().ramnet_pred_sigmoid(_h(x), _p(x), _p(skip), _p(wv), _p(ws), _p(b), _p(logits), None, N * H * W, C, _stream(x)))
    return logits


def pred_logits_bwd(dlogits, x, w, skip=None, concat=False):
    """Adjoint of pred_logits (ramnet_pred_bwd with depth = NULL)."""
    return pred_bwd(dlogits, None, x, w, skip, concat)


def upsample2x_bwd(dy):
    _check_nhwc(dy, 'upsample2x_bwd dy')
    N, C, H2, W2 = dy.shape
    dx = empty_nhwc(N, C, H2 // 2, W2 // 2, dy.device)
    check(_lib.load().ramnet_upsample2x_bwd(_h(dy), _p(dy), _p(dx), N, H2 // 2, W2 // 2, C, _stream(dy)))
    return dx


def lstm_bwd(dh, dc, gates, c_prev, c_new, round_tf32=False, db=None):
    N, C, H, W = c_prev.shape
    dz = empty_nhwc(N, 4 * C, H, W, c_prev.device)
    dc_prev = empty_nhwc(N, C, H, W, c_prev.device)
    check(_lib.load().ramnet_lstm_bwd(_h(c_prev), _p(dh), _p(dc), _p(gates), _p(c_prev), _p(c_new), _p(dz), _p(dc_prev),
                                      _p(db), N * H * W, C, FLAG_ROUND_TF32 if round_tf32 else 0, _stream(c_prev)))
    return dz, dc_prev


def msg_loss_stats(pred, target, start_scale=1, scales=4, want_signs=False):
    """-> stats [2 * scales] float64 (sum |g|, count per scale); with want_signs also the int8 sign pairs per pooled
    pixel that msg_loss_grad consumes (the only thing the backward pass needs from the forward pass)."""
    pred, target = pred.contiguous(), target.contiguous()
    N, C, H, W = pred.shape
    if C != 1:
        raise _lib.RamnetError('multi_scale_grad_loss: single-channel depth maps [N,1,H,W] expected')
    lib = _lib.load()
    P = lib.ramnet_msg_pooled_count(N, H, W, start_scale, scales)
    if P <= 0:
        raise _lib.RamnetError(f'multi_scale_grad_loss: {H}x{W} is not divisible by start_scale * 2^(scales-1)')
    stats = torch.empty(2 * scales, dtype=torch.float64, device=pred.device)
    ws = torch.empty(lib.ramnet_msg_workspace_bytes(N, H, W, start_scale, scales), dtype=torch.uint8, device=pred.device)
    signs = torch.empty((P, 2), dtype=torch.int8, device=pred.device) if want_signs else None
    check(lib.ramnet_msg_loss_stats(_h(pred), _p(pred), _p(target), N, H, W, start_scale, scales, _p(stats), _p(ws),
                                    _p(signs), _stream(pred)))
    return (stats, signs) if want_signs else stats


def msg_loss_value(stats, N, scales=4):
    out = torch.empty((), dtype=torch.float32, device=stats.device)
    check(_lib.load().ramnet_msg_loss_value(_h(stats), _p(stats), N, scales, _p(out), _stream(stats)))
    return out


def msg_loss_grad(signs, shape, stats, start_scale=1, scales=4, scale=1.0, n_batch=0, scale_dev=None):
    """d loss / d pred [N,1,H,W] from the sign pairs of msg_loss_stats(want_signs=True) and the (possibly all-reduced) stats."""
    N, C, H, W = shape
    dev = signs.device
    grad = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
    ws = torch.empty(signs.shape[0], dtype=torch.float32, device=dev)
    if scale_dev is not None and (scale_dev.dtype != torch.float32 or not scale_dev.is_cuda):
        scale_dev = scale_dev.to(device=dev, dtype=torch.float32)
    check(_lib.load().ramnet_msg_loss_grad(_h(signs), _p(signs), N, H, W, start_scale, scales, _p(stats), int(n_batch), scale,
                                           _p(scale_dev), _p(ws), _p(grad), _stream(signs)))
    return grad


def msg_sobel_preview(pred, target, pool: int) -> torch.Tensor:
    """[N,1,H/pool,W/pool] Sobel magnitude of AvgPool2d(pool)(pred - target) (MultiScaleGradient preview=True)."""
    pred, target = pred.detach().contiguous().float(), target.detach().contiguous().float()
    N, C, H, W = pred.shape
    if C != 1:
        raise _lib.RamnetError('msg_sobel_preview: single-channel depth maps [N,1,H,W] expected')
    out = torch.empty((N, 1, H // pool, W // pool), dtype=torch.float32, device=pred.device)
    check(_lib.load().ramnet_msg_sobel_preview(_h(pred), _p(pred), _p(target), N, H, W, int(pool), _p(out), _stream(pred)))
    return out


def adam_step_dev(p, g, m, v, step_counter, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, increment=True):
    """Graph-capturable fused Adam: `step_counter` is an int32 CUDA tensor holding the number of steps taken so far
    (incremented by this call when `increment`; later slices of the same step pass False)."""
    check(_lib.load().ramnet_adam_step_dev(_h(p), _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps,
                                           weight_decay, _p(step_counter), int(bool(increment)), _stream(p)))


def tf32_pipe_rate(device_index: int = 0):
    """(TFLOP/s, ms) of back-to-back tcgen05.mma kind::tf32 128x256x8 on every SM, measured now (synchronises)."""
    tf, ms = ctypes.c_double(), ctypes.c_double()
    check(_lib.load().ramnet_tf32_pipe_rate(_lib.handle(device_index), ctypes.byref(tf), ctypes.byref(ms)))
    return tf.value, ms.value
