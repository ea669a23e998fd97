"""Parameter containers with the reference's attribute names.

These modules own the `nn.Parameter`s under exactly the names, shapes and construction order of
RAM_Net/model/submodules.py, so `torch.manual_seed(0)` initialisation, `state_dict()` keys,
checkpoints, `torch.optim` and TensorBoard histograms behave as with the reference.  They hold
NO forward arithmetic: the engine (rpg_ramnet_b200/engine.py) reads their parameters, packs them
for the CUDA kernels and runs the fused graph.  Each class cites the reference block it mirrors.
"""
import torch.nn as nn
from torch.nn import init


def _norm_layer(norm, channels, momentum=0.1, in_track=True):
    if norm == 'BN':
        return nn.BatchNorm2d(channels, momentum=momentum)
    if norm == 'IN':
        return nn.InstanceNorm2d(channels, track_running_stats=in_track)
    return None


class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f'{type(self).__name__} is a parameter container; run the owning model, '
                           'which executes the fused CUDA graph')


class ConvLayer(_Container):
    """submodules.py:8-35 — conv2d (+ norm_layer); bias dropped iff norm == 'BN' (:13)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, activation='relu', norm=None,
                 BN_momentum=0.1):
        super().__init__()
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=(norm != 'BN'))
        self.activation = activation
        self.norm = norm
        nl = _norm_layer(norm, out_channels, BN_momentum)
        if nl is not None:
            self.norm_layer = nl


class UpsampleConvLayer(_Container):
    """submodules.py:69-97 — bilinear x2 then conv2d."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, activation='relu', norm=None):
        super().__init__()
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=(norm != 'BN'))
        self.activation = activation
        self.norm = norm
        nl = _norm_layer(norm, out_channels)
        if nl is not None:
            self.norm_layer = nl


class TransposedConvLayer(_Container):
    """submodules.py:38-66 — stride-2 ConvTranspose2d with output_padding 1."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, activation='relu', norm=None):
        super().__init__()
        self.transposed_conv2d = nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=2,
                                                    padding=padding, output_padding=1, bias=(norm != 'BN'))
        self.activation = activation
        self.norm = norm
        nl = _norm_layer(norm, out_channels)
        if nl is not None:
            self.norm_layer = nl


class ConvLSTM(_Container):
    """submodules.py:303-358 — one Gates conv, (input+hidden) -> 4*hidden."""

    def __init__(self, input_size, hidden_size, kernel_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.Gates = nn.Conv2d(input_size + hidden_size, 4 * hidden_size, kernel_size, padding=kernel_size // 2)


class ConvGRU(_Container):
    """submodules.py:414-454 — three gate convs, orthogonal weights, zero biases (:429-434)."""

    def __init__(self, input_size, hidden_size, kernel_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        pad = kernel_size // 2
        self.reset_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=pad)
        self.update_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=pad)
        self.out_gate = nn.Conv2d(input_size + hidden_size, hidden_size, kernel_size, padding=pad)
        for g in (self.reset_gate, self.update_gate, self.out_gate):
            init.orthogonal_(g.weight)
        for g in (self.reset_gate, self.update_gate, self.out_gate):
            init.constant_(g.bias, 0.)


def _recurrent_block(kind, channels):
    assert kind in ('convlstm', 'convgru')
    return (ConvLSTM if kind == 'convlstm' else ConvGRU)(input_size=channels, hidden_size=channels, kernel_size=3)


class RecurrentConvLayer(_Container):
    """submodules.py:100-120 — only a recurrent block (its conv is commented out upstream)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, recurrent_block_type='convlstm',
                 activation='relu', norm=None, BN_momentum=0.1):
        super().__init__()
        self.recurrent_block_type = recurrent_block_type
        self.recurrent_block = _recurrent_block(recurrent_block_type, out_channels)


class Recurrent2ConvLayer(_Container):
    """submodules.py:122-142 — ConvLayer followed by a recurrent block."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=0, recurrent_block_type='convlstm',
                 activation='relu', norm=None, BN_momentum=0.1):
        super().__init__()
        self.recurrent_block_type = recurrent_block_type
        self.conv = ConvLayer(in_channels, out_channels, kernel_size, stride, padding, activation, norm,
                              BN_momentum=BN_momentum)
        self.recurrent_block = _recurrent_block(recurrent_block_type, out_channels)


class ResidualBlock(_Container):
    """submodules.py:182-215 — conv1, (bn1, bn2), conv2."""

    def __init__(self, in_channels, out_channels, stride=1, downsample=None, norm=None, BN_momentum=0.1):
        super().__init__()
        bias = norm != 'BN'
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=stride, padding=1, bias=bias)
        self.norm = norm
        if norm == 'BN':
            self.bn1 = nn.BatchNorm2d(out_channels, momentum=BN_momentum)
            self.bn2 = nn.BatchNorm2d(out_channels, momentum=BN_momentum)
        elif norm == 'IN':
            self.bn1 = nn.InstanceNorm2d(out_channels)
            self.bn2 = nn.InstanceNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1, bias=bias)
        self.downsample = downsample
