"""rpg_ramnet_b200 — B200-native (sm_100a) RAM-Net hot path behind the reference's module surface.

    from rpg_ramnet_b200 import ERGB2DepthRecurrent, ERGB2Depth, events_to_voxel_grid, scale_invariant_loss

The arithmetic lives in libramnet_sm100a.so (hand-written CUDA, C ABI in include/ramnet_b200.h);
this package is the Python host side that mirrors the reference's operator interface.
"""
from ._lib import RamnetError, launch_count  # noqa: F401
from .model import ERGB2Depth, ERGB2DepthRecurrent, StateNetPhasedRecurrent, UNet  # noqa: F401
from .model.loss import MultiScaleGradient, multi_scale_grad_loss, scale_invariant_loss  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .utils.event_tensor_utils import events_to_voxel_grid, events_to_voxel_grid_pytorch  # noqa: F401

__version__ = '0.1.0'
