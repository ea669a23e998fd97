"""events_to_voxel_grid on the device (mirrors RAM_Net/utils/event_tensor_utils.py:71-117 and its
torch twin :120-187; live twin data_loader/dataset_asynchronous.py:253-298).

Same name and argument order `(events, num_bins, width, height)`.  `events` may be a [N,4] numpy
array or torch tensor (rows [timestamp, x, y, polarity]); unlike the reference the input is NOT
modified in place.  Returns a [num_bins, height, width] float32 CUDA tensor.
"""
import numpy as np
import torch

from .. import ops


def events_to_voxel_grid(events, num_bins, width, height, device=None, normalize=False):
    """`normalize=True` (extension): also applies the loaders' normalisation (mean / stddev of the non-zero voxels,
    data_loader/event_dataset.py:144-151) in the same call — the statistics are gathered while the grid is written."""
    assert events.shape[1] == 4
    assert num_bins > 0
    assert width > 0
    assert height > 0
    if isinstance(events, np.ndarray):
        ev = torch.from_numpy(np.ascontiguousarray(events, dtype=np.float64))
    else:
        ev = events
    if device is None:
        device = ev.device if ev.is_cuda else torch.device('cuda', torch.cuda.current_device())
    ev = ev.to(device=device, dtype=torch.float64, non_blocking=True).contiguous()
    return ops.voxel_grid_ex(ev, int(num_bins), int(width), int(height), normalize=bool(normalize))


def events_to_voxel_grid_pytorch(events, num_bins, width, height, device):
    """Same entry point name as the reference's torch variant (:120)."""
    return events_to_voxel_grid(events, num_bins, width, height, device=device)


def normalize_voxel_grid(voxel_grid):
    """Device twin of the loaders' normalisation (data_loader/event_dataset.py:144-151,
    dataset_asynchronous.py:300-308 `normalize_voxelgrid`, utils/event_tensor_utils.py:52-66): the mean and
    stddev of the NON-ZERO voxels become (0, 1); zeros stay zero; nothing happens when there are no events or
    the stddev is 0.  In place on a float32 CUDA tensor (as the reference mutates its array), returns it.
    A 4-D [B, bins, H, W] tensor is a batch of grids, each normalised on its own statistics in one launch pair."""
    return ops.voxel_normalize_(voxel_grid)


def depth_to_log_label(depth, clip_distance, reg_factor):
    """Device twin of the label transform in data_loader/dataset.py:296-305: clip to `clip_distance`,
    normalise, `1 + log(d) / reg_factor`, clip to [0, 1]; NaN (no ground truth) is preserved for the
    loss's mask (model/loss.py:7-8)."""
    return ops.depth_to_label(depth, clip_distance, reg_factor)
