"""ERGB2Depth / ERGB2DepthRecurrent — the drop-in boundary (mirrors RAM_Net/model/model.py:12-219).

Same constructor (`config` dict with the reference's keys), same
`forward(item, prev_super_states, prev_states_lstm) -> (predictions, super_states, states_lstm)`
contract, same schedule (K event passes then one image pass, a decoder pass after each,
model.py:161-217), so `trainer/lstm_trainer.py:270-272` and `test.py:230-232` can call it
unchanged.  One extra, optional config key: `mma_kind` ('tf32' default | 'fp32').
"""
import torch

from ..base import BaseModel
from .. import ops
from .._lib import RamnetError
from .statenet import StateNetPhasedRecurrent
from .unet import UNet


class BaseERGB2Depth(BaseModel):
    """Config parsing of model.py:12-77, key for key (defaults included)."""

    def __init__(self, config):
        super().__init__(config)
        assert 'num_bins_rgb' in config
        assert 'num_bins_events' in config
        self.num_bins_rgb = int(config['num_bins_rgb'])
        self.num_bins_events = int(config['num_bins_events'])
        self.skip_type = str(config.get('skip_type', 'sum'))
        self.state_combination = str(config.get('state_combination', 'sum'))
        self.num_encoders = int(config.get('num_encoders', 4))
        self.base_num_channels = int(config.get('base_num_channels', 32))
        self.num_residual_blocks = int(config.get('num_residual_blocks', 2))
        self.recurrent_block_type = str(config.get('recurrent_block_type', 'convlstm'))
        self.norm = str(config['norm']) if 'norm' in config else None
        self.use_upsample_conv = bool(config.get('use_upsample_conv', True))
        self.every_x_rgb_frame = config.get('every_x_rgb_frame', 1)
        self.baseline = config.get('baseline', False)
        self.loss_composition = config.get('loss_composition', False)
        self.kernel_size = int(config.get('kernel_size', 5))
        self.mma_kind = config.get('mma_kind', None)
        # optional: replay each pass as one CUDA graph (inference; see engine.GraphRunner for the aliasing contract)
        self.cuda_graphs = bool(config.get('cuda_graphs', False))
        # cuda_graphs: device-resident inputs are not being written by work still queued on the caller's stream, so the
        # runner need not order its input copies after that stream (engine.GraphRunner)
        self.inputs_static = bool(config.get('inputs_static', False))
        self.gpu = torch.device('cuda:' + str(config['gpu']))


class ERGB2Depth(BaseERGB2Depth):
    """model.py:79-111: non-recurrent UNet on item['image']."""

    def __init__(self, config):
        super().__init__(config)
        self.unet = UNet(num_input_channels=self.num_bins_rgb, num_output_channels=1, skip_type=self.skip_type,
                         activation='sigmoid', num_encoders=self.num_encoders,
                         base_num_channels=self.base_num_channels, num_residual_blocks=self.num_residual_blocks,
                         norm=self.norm, use_upsample_conv=self.use_upsample_conv, mma_kind=self.mma_kind)

    def forward(self, item, prev_super_states, prev_states_lstm):
        x = item['image'].to(self.gpu, non_blocking=True)
        return {'image': self.unet(x)}, {'image': None}, prev_states_lstm


class ERGB2DepthRecurrent(BaseERGB2Depth):
    """model.py:114-219."""

    def __init__(self, config):
        super().__init__(config)
        self.statenetphasedrecurrent = StateNetPhasedRecurrent(
            num_input_channels_rgb=self.num_bins_rgb, num_input_channels_events=self.num_bins_events,
            num_output_channels=1, skip_type=self.skip_type, state_combination=self.state_combination,
            activation='sigmoid', num_encoders=self.num_encoders, base_num_channels=self.base_num_channels,
            num_residual_blocks=self.num_residual_blocks, norm=self.norm, use_upsample_conv=self.use_upsample_conv,
            recurrent_block_type=self.recurrent_block_type, baseline=self.baseline, mma_kind=self.mma_kind)
        self.max_num_channels = self.base_num_channels * pow(2, self.num_encoders)
        self._runners = {}

    def _run_pass(self, which, x, prev_super_states, last, nxt=None):
        """forward_events / forward_images + forward_decoder, eagerly or as CUDA-graph replays.  `nxt`: the (which, x)
        of the pass that follows within the same item, so that the runner can start it under this pass's decoder."""
        net = self.statenetphasedrecurrent
        if self._graphs_active():
            runner = self._runner(x)
            runner.inputs_static = self.inputs_static
            s, pred = runner.run(which, x, prev_super_states, nxt)
            return s, {'encoders': [None] * net.num_encoders, 'state_comb': list(s)}, pred
        static = self.inputs_static
        if not x.is_cuda:
            fs = net._front_stream_on(torch.device('cuda', self.gpu) if isinstance(self.gpu, int) else torch.device(self.gpu))
            if fs is not None:          # host input: copied on the stream that consumes it, no tie to the caller's stream
                with torch.cuda.stream(fs):
                    x = x.to(self.gpu, non_blocking=True)
                static = True
        x = x.to(self.gpu, non_blocking=True)
        if prev_super_states is None:
            prev_super_states = self._zero_states(x.shape[0], x.shape[2], x.shape[3])
        return net._pass(which, x, prev_super_states, last, inputs_static=static)

    def _graphs_active(self):
        return self.cuda_graphs and self.statenetphasedrecurrent.graph_capable() and not self.training

    def _runner(self, x):
        B, _, H, W = x.shape
        key = (B, H, W)
        runner = self._runners.get(key)
        if runner is None:
            from ..engine import GraphRunner
            runner = self._runners[key] = GraphRunner(self.statenetphasedrecurrent, B, H, W, self.gpu)
        return runner

    def _stage_ahead(self, plan):
        """CUDA-graph mode: start the host-to-device copies of the passes still to run (engine.GraphRunner.stage)."""
        if self._graphs_active():
            for which, x in plan:
                self._runner(x).stage(which, x)

    def _zero_states(self, B, H, W):
        """model.py:146-159, allocated on the device directly (the reference builds them on the host
        and copies)."""
        pair = (not bool(self.baseline)) and self.state_combination == 'convlstm'
        states = []
        for i in range(self.num_encoders):
            h, w = int(H / pow(2, i + 1)), int(W / pow(2, i + 1))
            c = int(self.base_num_channels * pow(2, i + 1))
            if pair:
                states.append([ops.zeros_nhwc(B, c, h, w, self.gpu), ops.zeros_nhwc(B, c, h, w, self.gpu)])
            else:
                states.append(ops.zeros_nhwc(B, c, h, w, self.gpu))
        return states

    def forward(self, item, prev_super_states, prev_states_lstm):
        net = self.statenetphasedrecurrent
        predictions, super_states, states_lstm = {}, {}, {}
        if prev_super_states is None and not self._graphs_active():
            B, _, H, W = item['image'].shape
            prev_super_states = self._zero_states(B, H, W)
        bl, K = self.baseline, self.every_x_rgb_frame
        events_through_image_encoder = bl == 'ergb0' or (bl == 'e' and self.loss_composition == 'image')
        last = None
        if (not bool(bl)) or events_through_image_encoder:
            if events_through_image_encoder:      # model.py:165-168
                n_event_passes, last = K - 1, prev_states_lstm['image']
            else:                                  # model.py:170-172
                n_event_passes, last = K, prev_states_lstm['events{}'.format(K - 1)]
            which = 'images' if (bl == 'ergb0' or bl == 'e') else 'events'   # baselines have no event encoder
            plan = [(which, item['events{}'.format(k)]) for k in range(n_event_passes)] + [('images', item['image'])]
            for k in range(n_event_passes):
                key = 'events{}'.format(k)
                self._stage_ahead(plan[k:])
                s, l, predictions[key] = self._run_pass(which, item[key], prev_super_states, last, nxt=plan[k + 1])
                super_states[key], states_lstm[key] = s, l
                prev_super_states, last = s, l
        if (not bool(bl)) or bl == 'rgb' or (bl == 'e' and self.loss_composition != 'image'):
            last = prev_states_lstm['image']       # model.py:203-208
        s, l, predictions['image'] = self._run_pass('images', item['image'], prev_super_states, last)
        super_states['image'], states_lstm['image'] = s, l
        return predictions, super_states, states_lstm
