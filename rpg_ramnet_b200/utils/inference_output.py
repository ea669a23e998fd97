"""Device side of the inference driver's output stage (mirrors RAM_Net/test.py:259-360,365-379).

`test.py` pulls every fp32 depth map to the host and derives an 8-bit grey PNG, a colour-mapped PNG and (with
--calculate_scale) a metric-space scale factor from it with numpy / matplotlib / OpenCV.  `depth_outputs` produces the
same payloads with one reduction + one streaming kernel (ramnet_depth_output): uint8 grey [N, H, W], uint8 BGR
[N, H, W, 3] and the two sums of the scale factor, so the host only has to `cv2.imwrite` bytes it received.

    lut = colormap_lut(color_mapper_overall)            # once: the reference's ScalarMappable sampled at i / 255
    grey, bgr, scale = depth_outputs(pred, lut=lut, target=gt, reg_factor=reg, clip_distance=clip)
    cv2.imwrite(path, grey[0].cpu().numpy());  cv2.imwrite(path2, bgr[0].cpu().numpy())
"""
import ctypes

import numpy as np
import torch

from .. import _lib, ops


def colormap_lut(color_mapper=None):
    """256 x 3 float32 RGB table: the caller's matplotlib ScalarMappable (test.py:204-205) sampled at i / 255.  Without
    matplotlib a perceptually ordered fallback ramp is returned (documented, only used when no mapper is given)."""
    if color_mapper is not None:
        x = np.arange(256, dtype=np.float64) / 255.0
        return np.ascontiguousarray(np.asarray(color_mapper.to_rgba(x))[:, :3], dtype=np.float32)
    t = np.linspace(0.0, 1.0, 256)
    return np.stack([t ** 0.5, t ** 1.5, 0.25 + 0.5 * np.sin(np.pi * t) ** 2], 1).astype(np.float32)


def depth_outputs(depth, lut=None, target=None, reg_factor=None, clip_distance=None, want_grey=True):
    """depth (and target): [N, 1, H, W] float32 CUDA.  Returns (grey uint8 [N,H,W] | None, bgr uint8 [N,H,W,3] | None,
    scale float64 [N] | None) — scale[n] = sum(p t) / sum(p p) in metric space (test.py:376)."""
    d = depth.detach().contiguous().float()
    if not d.is_cuda or d.dim() != 4 or d.shape[1] != 1:
        raise _lib.RamnetError('depth_outputs: a [N, 1, H, W] CUDA tensor is required')
    N, _, H, W = d.shape
    dev = d.device
    grey = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if want_grey else None
    bgr, lut_d = None, None
    if lut is not None:
        lut_d = torch.as_tensor(np.asarray(lut, dtype=np.float32)).contiguous().to(dev)
        if tuple(lut_d.shape) != (256, 3):
            raise _lib.RamnetError('depth_outputs: lut must be 256 x 3 (RGB in [0, 1])')
        bgr = torch.empty((N, H, W, 3), dtype=torch.uint8, device=dev)
    sums, t = None, None
    if target is not None and reg_factor is not None and clip_distance is not None:
        t = target.detach().contiguous().float().to(dev)
        if t.shape != d.shape:
            raise _lib.RamnetError('depth_outputs: target shape differs from depth')
        sums = torch.empty((N, 2), dtype=torch.float64, device=dev)
    scratch = torch.empty(N * 4, dtype=torch.int32, device=dev)
    p = ops._p
    ops.check(_lib.load().ramnet_depth_output(ops._h(d), p(d), p(t), N, H * W, p(lut_d), p(grey), p(bgr), p(sums),
                                              float(reg_factor or 0.0), float(clip_distance or 0.0), p(scratch),
                                              ops._stream(d)))
    scale = None if sums is None else sums[:, 0] / sums[:, 1]
    return grey, bgr, scale
