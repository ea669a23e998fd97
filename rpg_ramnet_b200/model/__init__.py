from .model import ERGB2Depth, ERGB2DepthRecurrent  # noqa: F401
from .statenet import StateNetPhasedRecurrent  # noqa: F401
from .unet import UNet  # noqa: F401
