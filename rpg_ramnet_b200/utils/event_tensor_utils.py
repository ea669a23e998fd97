"""events_to_voxel_grid on the device (mirrors RAM_Net/utils/event_tensor_utils.py:71-117 and its
torch twin :120-187; live twin data_loader/dataset_asynchronous.py:253-298).

Same name and argument order `(events, num_bins, width, height)`.  `events` may be a [N,4] numpy
array or torch tensor (rows [timestamp, x, y, polarity]); unlike the reference the input is NOT
modified in place.  Returns a [num_bins, height, width] float32 CUDA tensor.
"""
import numpy as np
import torch

from .. import ops


def events_to_voxel_grid(events, num_bins, width, height, device=None):
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
    return ops.voxel_grid(ev, int(num_bins), int(width), int(height))


def events_to_voxel_grid_pytorch(events, num_bins, width, height, device):
    """Same entry point name as the reference's torch variant (:120)."""
    return events_to_voxel_grid(events, num_bins, width, height, device=device)
