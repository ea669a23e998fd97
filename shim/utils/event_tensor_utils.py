"""Overlay for RAM_Net/utils/event_tensor_utils.py:71-187 (same names and argument order)."""
from rpg_ramnet_b200.utils.event_tensor_utils import (depth_to_log_label, events_to_voxel_grid,  # noqa: F401
                                                      events_to_voxel_grid_pytorch, normalize_voxel_grid)

__all__ = ['events_to_voxel_grid', 'events_to_voxel_grid_pytorch', 'normalize_voxel_grid', 'depth_to_log_label']
