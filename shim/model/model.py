"""Overlay for RAM_Net/model/model.py: the names `from model.model import *` must provide (train.py:6,204; test.py:6,170;
utils/loading_utils.py:17) now resolve to the B200 path.  Same class names, config keys, forward() contract."""
from rpg_ramnet_b200.base import BaseModel                                             # noqa: F401
from rpg_ramnet_b200.model.model import BaseERGB2Depth, ERGB2Depth, ERGB2DepthRecurrent  # noqa: F401
from rpg_ramnet_b200.model.statenet import StateNetPhasedRecurrent                     # noqa: F401
from rpg_ramnet_b200.model.unet import UNet                                            # noqa: F401

__all__ = ['BaseModel', 'BaseERGB2Depth', 'ERGB2Depth', 'ERGB2DepthRecurrent', 'StateNetPhasedRecurrent', 'UNet']
