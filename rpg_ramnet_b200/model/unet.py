"""UNet — the non-recurrent baseline graph (mirrors RAM_Net/model/unet.py:23-131)."""
import torch.nn as nn

from .. import engine as E
from .. import ops
from .._lib import RamnetError
from .submodules import ConvLayer, ResidualBlock, TransposedConvLayer, UpsampleConvLayer


class BaseUNet(nn.Module):
    """unet.py:23-84."""

    def __init__(self, num_input_channels, num_output_channels=1, skip_type='sum', activation='sigmoid',
                 num_encoders=4, base_num_channels=32, num_residual_blocks=2, norm=None, use_upsample_conv=True,
                 kernel_size=5):
        super().__init__()
        if skip_type not in ('sum', 'concat', 'no_skip', None):
            raise KeyError('Could not identify skip_type, please add "skip_type": "sum", "concat" or "no_skip" '
                           'to config["model"]')
        assert num_input_channels > 0 and num_output_channels > 0
        self.num_input_channels, self.num_output_channels = num_input_channels, num_output_channels
        self.skip_type, self.activation_name, self.norm, self.kernel_size = skip_type, activation, norm, kernel_size
        self.use_upsample_conv = use_upsample_conv
        print('Using UpsampleConvLayer (slow, but no checkerboard artefacts)' if use_upsample_conv else
              'Using TransposedConvLayer (fast, with checkerboard artefacts)')
        self.UpsampleLayer = UpsampleConvLayer if use_upsample_conv else TransposedConvLayer
        self.num_encoders, self.base_num_channels = num_encoders, base_num_channels
        self.num_residual_blocks = num_residual_blocks
        self.max_num_channels = base_num_channels * pow(2, num_encoders)
        self.encoder_input_sizes = [base_num_channels * pow(2, i) for i in range(num_encoders)]
        self.encoder_output_sizes = [base_num_channels * pow(2, i + 1) for i in range(num_encoders)]

    def build_resblocks(self):
        self.resblocks = nn.ModuleList(
            ResidualBlock(self.max_num_channels, self.max_num_channels, norm=self.norm)
            for _ in range(self.num_residual_blocks))

    def build_decoders(self):
        self.decoders = nn.ModuleList()
        for c in reversed(self.encoder_output_sizes):
            self.decoders.append(self.UpsampleLayer(c if self.skip_type == 'sum' else 2 * c, c // 2,
                                                    kernel_size=5, padding=2, norm=self.norm))

    def build_prediction_layer(self):
        self.pred = ConvLayer(self.base_num_channels if self.skip_type == 'sum' else 2 * self.base_num_channels,
                              self.num_output_channels, 1, activation=None, norm=self.norm)


class UNet(BaseUNet):
    """unet.py:87-131: head, encoders, resblocks, decoders with a skip on EVERY level, pred(x + head)."""

    def __init__(self, num_input_channels, num_output_channels=1, skip_type='sum', activation='sigmoid',
                 num_encoders=4, base_num_channels=32, num_residual_blocks=2, norm=None, use_upsample_conv=True,
                 mma_kind=None):
        super().__init__(num_input_channels, num_output_channels, skip_type, activation, num_encoders,
                         base_num_channels, num_residual_blocks, norm, use_upsample_conv)
        self.head = ConvLayer(num_input_channels, base_num_channels, kernel_size=5, stride=1, padding=2)
        self.encoders = nn.ModuleList(
            ConvLayer(cin, cout, kernel_size=5, stride=2, padding=2, norm=norm)
            for cin, cout in zip(self.encoder_input_sizes, self.encoder_output_sizes))
        self.build_resblocks()
        self.build_decoders()
        self.build_prediction_layer()
        self._mma_kind_name = mma_kind
        self._wcache = E.WeightCache()

    def forward(self, x, return_logits=False):
        if self.skip_type not in ('sum', 'concat') or self.activation_name != 'sigmoid' or self.num_output_channels != 1:
            # skip_type='no_skip' is ill-formed upstream (decoders are built for 2C channels, unet.py:78, and fed C)
            raise RamnetError("UNet: skip_type 'sum' or 'concat', sigmoid, 1 output channel are implemented")
        concat = self.skip_type == 'concat'
        if concat and not self.use_upsample_conv:
            raise RamnetError("UNet: skip_type='concat' with TransposedConvLayer decoders is not implemented")
        kind = E.resolve_mma_kind(self._mma_kind_name)
        tf32 = kind == ops.MMA_TF32
        cache, n = self._wcache, self.num_encoders
        if x.dim() != 4 or x.shape[2] % (1 << n) or x.shape[3] % (1 << n):
            raise RamnetError(f'input {tuple(x.shape)}: H and W must be divisible by {1 << n}')
        x = E.head_layer(cache, 'head', self.head.conv2d, x, tf32)
        head, blocks = x, []
        for i, enc in enumerate(self.encoders):
            x = E.conv_layer(cache, f'enc{i}', enc.conv2d, kind, x, ops.EPI_BIAS_RELU,
                             norm_mod=getattr(enc, 'norm_layer', None), norm_kind=enc.norm, training=self.training,
                             round_out=True)
            blocks.append(x)
        for i, rb in enumerate(self.resblocks):
            y = E.conv_layer(cache, f'res{i}/1', rb.conv1, kind, x, ops.EPI_BIAS_RELU, norm_mod=getattr(rb, 'bn1', None),
                             norm_kind=rb.norm, training=self.training, round_out=True)
            x = E.conv_layer(cache, f'res{i}/2', rb.conv2, kind, y, ops.EPI_BIAS_RES_RELU, res=x,
                             norm_mod=getattr(rb, 'bn2', None), norm_kind=rb.norm, training=self.training, round_out=True)
        for i, dec in enumerate(self.decoders):
            if not self.use_upsample_conv:        # TransposedConvLayer decoders (unet.py:48-51), skip sum fused in
                x = E.transposed_conv_layer(cache, f'dec{i}', dec.transposed_conv2d, kind, x, blocks[n - i - 1],
                                            getattr(dec, 'norm_layer', None), dec.norm, self.training)
                continue
            if concat:
                # cat([x, skip]) -> bilinear x2 -> conv == conv over the virtual concat [up(x) | up(skip)]: the kernel's two
                # K segments (the ConvGRU's [x | h] mechanism), no concatenated tensor (unet.py:11-12,126-127)
                up, up1 = E.upsample_add(x, None, tf32), E.upsample_add(blocks[n - i - 1], None, tf32)
            else:
                up, up1 = E.upsample_add(x, blocks[n - i - 1], tf32), None
            x = E.conv_layer(cache, f'dec{i}', dec.conv2d, kind, up, ops.EPI_BIAS_RELU, x1=up1,
                             norm_mod=getattr(dec, 'norm_layer', None), norm_kind=dec.norm, training=self.training)
        pr = self.pred
        return E.pred_layer(x, pr.conv2d, getattr(pr, 'norm_layer', None), pr.norm, self.training, return_logits, skip=head,
                            concat=concat)
