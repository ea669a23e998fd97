"""Import the UNMODIFIED reference Python modules from /root/reference/RAM_Net.

TEST INFRASTRUCTURE ONLY, and only usable in the build container:
/root/reference does not exist on the GPU box, so nothing in the `-m gpu`
tests, smoke() or bench.py imports this module.  It is used by
``oracle/make_golden.py`` to generate ``tests/golden/*.npz`` and by the
container-only CPU tests that compare the oracle with the live reference.

Stubs: the reference imports kornia (model/loss.py:3), skimage
(model/metric.py:2) and matplotlib (trainer/lstm_trainer.py) which are not
installed; none of them is on the path we exercise (scale_invariant_loss,
model forward).  numpy>=1.24 removed ``np.int`` which
event_tensor_utils.py:97-102 uses.
"""
import os
import sys
import types

REF_ROOT = '/root/reference/RAM_Net'


def available() -> bool:
    return os.path.isdir(REF_ROOT)


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference's model classes, loss and voxeliser."""
    if not available():
        raise RuntimeError('reference not present at ' + REF_ROOT)
    import numpy as np
    if not hasattr(np, 'int'):
        np.int = int  # noqa: removed alias used by the reference
    try:
        import kornia  # noqa: F401
    except Exception:
        _stub('kornia')
        _stub('kornia.filters')
        _stub('kornia.filters.sobel', spatial_gradient=None, sobel=None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # make sure `model`, `base`, `utils` resolve to the reference, not to anything else
    for k in [k for k in sys.modules if k == 'model' or k.startswith('model.')]:
        if not getattr(sys.modules[k], '__file__', '').startswith(REF_ROOT):
            del sys.modules[k]
    import model.model as ref_model
    import model.loss as ref_loss
    import utils.event_tensor_utils as ref_vox
    ns = types.SimpleNamespace()
    ns.ERGB2DepthRecurrent = ref_model.ERGB2DepthRecurrent
    ns.ERGB2Depth = ref_model.ERGB2Depth
    ns.scale_invariant_loss = ref_loss.scale_invariant_loss
    ns.events_to_voxel_grid = ref_vox.events_to_voxel_grid
    return ns


def build_model(ns, arch: str, config: dict, seed: int = 0):
    """Construct like train.py:203-204 (manual_seed(0) then eval(arch)(config['model']))
    and retarget the device object to CPU (the ctor only builds it, model.py:77)."""
    import contextlib
    import io
    import torch
    cfg = dict(config)
    cfg.setdefault('gpu', 0)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = getattr(ns, arch)(cfg)
    m.gpu = torch.device('cpu')
    m.eval()
    return m
