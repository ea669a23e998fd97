"""StateNetPhasedRecurrent — the RAM-Net graph (mirrors RAM_Net/model/statenet.py:120-315).

Same constructor signature, child-module names/order (=> identical seeded init and state_dict keys)
and the same three entry points `forward_events`, `forward_images`, `forward_decoder` with the
same argument and return structure.  The arithmetic is the fused CUDA graph of
rpg_ramnet_b200/engine.py; activations and states are fp32 NHWC-strided [N,C,H,W] tensors.
"""
import os

import torch
import torch.nn as nn

from .. import engine as E
from .. import ops
from .._lib import RamnetError
from .submodules import (ConvLayer, Recurrent2ConvLayer, RecurrentConvLayer, ResidualBlock, TransposedConvLayer,
                         UpsampleConvLayer)


class BaseStateNet(nn.Module):
    """statenet.py:39-117 (argument validation + the shared tail builders)."""

    def __init__(self, num_input_channels_rgb, num_input_channels_events, num_output_channels=1, skip_type='sum',
                 state_combination='sum', activation='sigmoid', num_encoders=4, base_num_channels=32,
                 num_residual_blocks=2, norm=None, use_upsample_conv=True, recurrent_block_type='convlstm',
                 baseline=False):
        super().__init__()
        if skip_type not in ('sum', 'concat', 'no_skip', None):
            raise KeyError('Could not identify skip_type, please add "skip_type": "sum", "concat" or "no_skip" '
                           'to config["model"]')
        if state_combination not in ('sum', 'conv', 'convlstm', 'convgru'):
            raise KeyError('Could not identify state_combination, please add "state_combination": "sum", "conv", '
                           '"convlstm" or "convgru" to config["model"]')
        assert num_input_channels_rgb > 0 or num_input_channels_events > 0
        assert num_output_channels > 0
        self.num_input_channels_rgb = num_input_channels_rgb
        self.num_input_channels_events = num_input_channels_events
        self.num_output_channels = num_output_channels
        self.skip_type = skip_type
        self.state_combination = state_combination
        self.recurrent_block_type = recurrent_block_type
        self.activation_name = activation
        self.norm = norm
        self.baseline = baseline
        self.use_upsample_conv = use_upsample_conv
        print('Using UpsampleConvLayer (slow, but no checkerboard artefacts)' if use_upsample_conv else
              'Using TransposedConvLayer (fast, with checkerboard artefacts)')
        self.UpsampleLayer = UpsampleConvLayer if use_upsample_conv else TransposedConvLayer
        self.num_encoders = num_encoders
        self.base_num_channels = base_num_channels
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


def _flat_tensors(obj):
    """Every tensor inside a (nested) state container: lists / tuples / dicts of tensors or None."""
    if obj is None:
        return []
    if torch.is_tensor(obj):
        return [obj]
    if isinstance(obj, dict):
        obj = list(obj.values())
    out = []
    for o in obj:
        out.extend(_flat_tensors(o))
    return out


class StateNetPhasedRecurrent(BaseStateNet):
    def __init__(self, num_input_channels_rgb, num_input_channels_events, num_output_channels=1, skip_type='sum',
                 state_combination='sum', activation='sigmoid', num_encoders=4, base_num_channels=32,
                 num_residual_blocks=2, norm=None, use_upsample_conv=True, recurrent_block_type='convlstm',
                 baseline=False, mma_kind=None):
        super().__init__(num_input_channels_rgb, num_input_channels_events, num_output_channels, skip_type,
                         state_combination, activation, num_encoders, base_num_channels, num_residual_blocks, norm,
                         use_upsample_conv, recurrent_block_type, baseline)
        has_events = not bool(baseline)
        stateful = state_combination in ('convlstm', 'conv', 'convgru')
        # registration order == reference (statenet.py:139-155): head_rgb, encoders_rgb, head_events,
        # encoders_events, state_combination_events, state_combination_images
        self.head_rgb = ConvLayer(num_input_channels_rgb, base_num_channels, kernel_size=5, stride=1, padding=2)
        self.encoders_rgb = nn.ModuleList()
        if has_events:
            self.head_events = ConvLayer(num_input_channels_events, base_num_channels, kernel_size=5, stride=1,
                                         padding=2)
            self.encoders_events = nn.ModuleList()
        if stateful:
            if has_events:
                self.state_combination_events = nn.ModuleList()
            self.state_combination_images = nn.ModuleList()
        else:
            self.state_combination_events, self.state_combination_images = [], []

        def encoder(cin, cout):
            if recurrent_block_type == 'convlstm':
                return Recurrent2ConvLayer(cin, cout, kernel_size=5, stride=2, padding=2, norm=norm,
                                           recurrent_block_type=recurrent_block_type)
            if recurrent_block_type == 'conv':
                return ConvLayer(cin, cout, kernel_size=5, stride=2, padding=2, norm=norm)
            return None

        def combiner(cout):
            if state_combination in ('convlstm', 'convgru'):
                return RecurrentConvLayer(cout, cout, kernel_size=5, stride=1, padding=2, norm=norm,
                                          recurrent_block_type=state_combination)
            if state_combination == 'conv':
                return ConvLayer(cout * 2, cout, kernel_size=5, stride=1, padding=2, norm=norm)
            return None

        # per level: encoder_rgb, encoder_events, comb_events, comb_images (statenet.py:157-198)
        for cin, cout in zip(self.encoder_input_sizes, self.encoder_output_sizes):
            e = encoder(cin, cout)
            if e is not None:
                self.encoders_rgb.append(e)
            if has_events:
                e = encoder(cin, cout)
                if e is not None:
                    self.encoders_events.append(e)
            if has_events:
                self.state_combination_events.append(combiner(cout))
            self.state_combination_images.append(combiner(cout))

        self.build_resblocks()
        self.build_decoders()
        self.build_prediction_layer()
        self._mma_kind_name = mma_kind
        self._wcache = E.WeightCache()
        self._front_streams, self._front_tok = {}, None

    # ---- configuration the CUDA graph supports ------------------------------------------------
    def _kind(self):
        return E.resolve_mma_kind(self._mma_kind_name)

    def _check_supported(self):
        if self.state_combination in ('sum', 'conv'):
            # statenet.py:231 / :272 tuple-unpack the single tensor state_sum/state_conv return:
            # ill-formed upstream (works only by accident for batch size 2)
            raise RamnetError(f"state_combination={self.state_combination!r} is ill-formed in the reference "
                              "(statenet.py:231 unpacks one tensor into two); use 'convgru' or 'convlstm'")
        if self.recurrent_block_type not in ('conv', 'convlstm'):
            raise RamnetError(f'recurrent_block_type={self.recurrent_block_type!r} is not supported')
        if self.skip_type != 'sum':
            # skip_type='concat' crashes in the reference StateNet too (decoder 0 gets no skip, :302-303)
            raise RamnetError("only skip_type='sum' is well-formed for StateNetPhasedRecurrent")
        if self.activation_name != 'sigmoid':
            raise RamnetError('only the sigmoid output activation is implemented')
        if self.num_output_channels != 1:
            raise RamnetError('only num_output_channels=1 is implemented')

    # ---- encoders -----------------------------------------------------------------------------
    def _pass(self, which, x, prev_super_state, prev_states_lstm, out_states=None, return_logits=False, inputs_static=False):
        """One full pass: encoder of one modality + state update + decoder.  `out_states` (optional) are
        preallocated state buffers the new super states are written into (CUDA-graph runner).

        Eager path (training, cuda_graphs off), round 2: the FRONT (head, encoders, state update) runs on a second stream
        and the decoder on the caller's, so the front of the next pass -- which needs this pass's states, not its
        decoder -- runs under this pass's decoder; autograd replays every node on its forward stream, so the backward
        pass overlaps the same way.  Everything returned is ordered on the caller's stream.  RAMNET_FRONT_STREAM=0
        disables.  The front stream waits for the caller's stream only when it has to: new weights, states it did not
        produce itself, or a device input that is not declared static."""
        fs = self._front_stream(x)
        if fs is None or out_states is not None:
            s, l = self._encode(which, x, prev_super_state, prev_states_lstm, out_states)
            return s, l, self.forward_decoder(s, return_logits)
        cur = torch.cuda.current_stream(x.device)
        tok = (E._WEIGHT_EPOCH, self.training) + tuple(p._version for p in self.parameters())
        flat_prev = _flat_tensors(prev_super_state) + _flat_tensors(prev_states_lstm)
        ours = all(getattr(t, '_ramnet_front', None) is fs for t in flat_prev)
        if tok != self._front_tok or not ours or not inputs_static:
            fs.wait_stream(cur)
            self._front_tok = tok
        with torch.cuda.stream(fs):
            s, l = self._encode(which, x, prev_super_state, prev_states_lstm, None)
            done = torch.cuda.Event()
            done.record(fs)
        x.record_stream(fs)
        for t in flat_prev:
            t.record_stream(fs)
        for t in _flat_tensors(s) + _flat_tensors(l):     # read by the decoder / the caller on `cur`, by the next front on `fs`
            t.record_stream(cur)
            try:
                t._ramnet_front = fs
            except AttributeError:
                pass
        cur.wait_event(done)
        return s, l, self.forward_decoder(s, return_logits)

    def _front_stream(self, x):
        return self._front_stream_on(x.device) if x.is_cuda else None

    def _front_stream_on(self, device):
        if device.type != 'cuda' or os.environ.get('RAMNET_FRONT_STREAM', '1') == '0':
            return None
        device = torch.device('cuda', torch.cuda.current_device() if device.index is None else device.index)
        fs = self._front_streams.get(device)
        if fs is None:
            fs = self._front_streams[device] = torch.cuda.Stream(device=device)
        return fs

    def graph_capable(self):
        """CUDA-graph replay covers the configurations whose only recurrent state is the super state."""
        return self.recurrent_block_type == 'conv' and not bool(self.baseline) and \
            self.state_combination in ('convgru', 'convlstm')

    def _encode(self, which, x, prev_super_state, prev_states_lstm, out_states=None):
        self._check_supported()
        kind = self._kind()
        tf32 = kind == ops.MMA_TF32
        cache, n = self._wcache, self.num_encoders
        head = self.head_events if which == 'events' else self.head_rgb
        encoders = self.encoders_events if which == 'events' else self.encoders_rgb
        combs = self.state_combination_events if which == 'events' else self.state_combination_images
        baseline_path = bool(self.baseline) and which == 'images'
        if x.dim() != 4 or x.shape[2] % (1 << n) or x.shape[3] % (1 << n):
            raise RamnetError(f'input {tuple(x.shape)}: H and W must be divisible by 2**num_encoders = {1 << n} '
                              '(the reference fails on such shapes too, SURVEY Appendix A)')
        x = E.head_layer(cache, which + '/head', head.conv2d, x, tf32)
        if prev_states_lstm is None:
            prev_states_lstm = {'encoders': [None] * n, 'state_comb': [None] * n}
        super_states, states_lstm = [], {'encoders': [], 'state_comb': []}
        for i in range(n):
            enc = encoders[i]
            if self.recurrent_block_type == 'conv':
                x = E.conv_layer(cache, f'{which}/enc{i}', enc.conv2d, kind, x, ops.EPI_BIAS_RELU,
                                 norm_mod=getattr(enc, 'norm_layer', None), norm_kind=enc.norm, training=self.training,
                                 round_out=True)
                enc_state = None
            else:
                x = E.conv_layer(cache, f'{which}/enc{i}', enc.conv.conv2d, kind, x, ops.EPI_BIAS_RELU,
                                 norm_mod=getattr(enc.conv, 'norm_layer', None), norm_kind=enc.conv.norm,
                                 training=self.training, round_out=True)
                enc_state = E.lstm_layer(cache, f'{which}/enc{i}/lstm', enc.recurrent_block, kind, x,
                                         prev_states_lstm['encoders'][i])
                x = enc_state[0]
            blk = combs[i].recurrent_block
            if self.state_combination == 'convlstm' and not baseline_path:
                st = E.lstm_layer(cache, f'{which}/comb{i}', blk, kind, x, prev_super_state[i],   # state = prev super state
                                  out_state=None if out_states is None else out_states[i])
                super_state = comb_state = st
            elif self.state_combination == 'convgru':
                hprev = None if prev_super_state[i] is None else ops.as_nhwc(prev_super_state[i])
                hnew = E.gru_layer(cache, f'{which}/comb{i}', blk, kind, x, hprev,
                                   out_h=None if out_states is None else out_states[i])
                super_state = comb_state = hnew
                if baseline_path:
                    x = hnew
            else:  # baseline + convlstm: private LSTM state, recurrent output feeds the next encoder
                st = E.lstm_layer(cache, f'{which}/comb{i}', blk, kind, x, prev_states_lstm['state_comb'][i])
                x, super_state, comb_state = st[0], st[0], st
            super_states.append(super_state)
            states_lstm['encoders'].append(enc_state)
            states_lstm['state_comb'].append(comb_state)
        return super_states, states_lstm

    def forward_events(self, x, prev_super_state, prev_states_lstm, times=None):
        """statenet.py:204-239."""
        if bool(self.baseline):
            raise RamnetError('baseline models have no event encoder (statenet.py:143-147)')
        return self._encode('events', x, prev_super_state, prev_states_lstm)

    def forward_images(self, x, prev_super_state, prev_states_lstm, times=None):
        """statenet.py:241-288."""
        return self._encode('images', x, prev_super_state, prev_states_lstm)

    # ---- decoder ------------------------------------------------------------------------------
    def forward_decoder(self, super_states, return_logits=False):
        """statenet.py:290-315: resblocks(S[-1]) -> dec0(x) -> dec_i(x + S[n-i-1]) -> pred -> sigmoid."""
        self._check_supported()
        kind = self._kind()
        cache, n = self._wcache, self.num_encoders
        tup = (not bool(self.baseline)) and self.state_combination == 'convlstm'
        pick = (lambda s: s[0]) if tup else (lambda s: s)
        x = ops.as_nhwc(pick(super_states[-1]))
        for i, rb in enumerate(self.resblocks):
            y = E.conv_layer(cache, f'res{i}/1', rb.conv1, kind, x, ops.EPI_BIAS_RELU, norm_mod=getattr(rb, 'bn1', None),
                             norm_kind=rb.norm, training=self.training, round_out=True)
            x = E.conv_layer(cache, f'res{i}/2', rb.conv2, kind, y, ops.EPI_BIAS_RES_RELU, res=x,
                             norm_mod=getattr(rb, 'bn2', None), norm_kind=rb.norm, training=self.training, round_out=True)
        pr = self.pred
        tf32 = kind == ops.MMA_TF32
        nd = len(self.decoders)

        def skip_of(i):
            return None if i == 0 or i >= nd else ops.as_nhwc(pick(super_states[n - i - 1]))

        def grad_free(dec):
            return not E.needs_grad(x, dec.conv2d.weight, dec.conv2d.bias, pr.conv2d.weight, pr.conv2d.bias)

        # Up-conv decoders (inference, TF32): bilinear x2 + 5x5 conv in ONE launch on the low-resolution tensor
        # (ops.conv_up_fwd); the skip sum a decoder needs is formed by the epilogue of the layer before it
        # (EPI_BIAS_RELU_ADD), so neither the 4x tensor nor the sum is ever written on its own.
        # live norms (train-mode statistics) sit between a conv and its activation: those layers run unfused
        live_norm = any(E.norm_is_live(getattr(m, 'norm_layer', None), m.norm) for m in list(self.decoders) + [pr])
        up_mode = [self.use_upsample_conv and tf32 and grad_free(dec) and not live_norm and
                   ops.upconv_eligible(dec.conv2d.in_channels, dec.conv2d.out_channels, dec.conv2d.kernel_size[0], kind)
                   for dec in self.decoders]
        skip_added = False          # x already holds x + skip of the decoder about to run
        for i, dec in enumerate(self.decoders):
            skip = None if skip_added else skip_of(i)
            skip_added = False
            if not self.use_upsample_conv:        # TransposedConvLayer decoder (statenet.py:81-82)
                x = E.transposed_conv_layer(cache, f'dec{i}', dec.transposed_conv2d, kind, x, skip,
                                            getattr(dec, 'norm_layer', None), dec.norm, self.training)
                continue
            last = i == nd - 1
            next_up = (not last) and up_mode[i + 1]
            nm, nk = getattr(dec, 'norm_layer', None), dec.norm
            fuse = last and tf32 and dec.conv2d.out_channels % 32 == 0 and dec.conv2d.out_channels <= 256 and \
                grad_free(dec) and not live_norm
            if fuse:
                pw, pb = pr.conv2d.weight.detach().float(), None if pr.conv2d.bias is None else pr.conv2d.bias.detach().float()
                pw, pb = E._fold_norm(pw, pb, getattr(pr, 'norm_layer', None), pr.norm, self.training)
                pbias = pb if pb is not None else torch.zeros(1, dtype=torch.float32, device=x.device)
                pw, pbias = pw.reshape(-1).contiguous(), pbias.reshape(-1).contiguous()
            if up_mode[i]:
                if skip is not None:              # unreachable: the layer before an up-conv decoder always forms its skip sum
                    raise RamnetError('internal: up-conv decoder reached without its skip sum')
                p = E.pack_upconv(cache, f'dec{i}', dec.conv2d, nm, nk, self.training)
                if fuse:
                    N, _, Hh, Ww = x.shape
                    logits = torch.empty((N, 1, 2 * Hh, 2 * Ww), dtype=torch.float32, device=x.device) if return_logits else None
                    depth = ops.conv_up_fwd(x, p.w, p.b, p.Cout, ops.EPI_BIAS_RELU_PRED, aux0=pw, aux1=pbias, out1=logits)
                    return (depth, logits) if return_logits else depth
                if next_up:
                    x = ops.conv_up_fwd(x, p.w, p.b, p.Cout, ops.EPI_BIAS_RELU_ADD, aux0=skip_of(i + 1), round_tf32=True)
                    skip_added = True
                else:
                    x = ops.conv_up_fwd(x, p.w, p.b, p.Cout, ops.EPI_BIAS_RELU)
                continue
            up = E.upsample_add(x, skip, tf32)
            if fuse:
                # last decoder + pred + sigmoid in one kernel: the 32-channel full-resolution tensor is never written
                p = E.pack_conv(cache, f'dec{i}', dec.conv2d, kind, nm, nk, self.training, hpack_ok=True)
                N, _, Hh, Ww = up.shape
                logits = torch.empty((N, 1, Hh, Ww), dtype=torch.float32, device=up.device) if return_logits else None
                depth = ops.conv_fwd(up, None, p.w, p.b, p.Cout, p.ksize, p.stride, ops.EPI_BIAS_RELU_PRED, kind,
                                     aux0=pw, aux1=pbias, out1=logits)
                return (depth, logits) if return_logits else depth
            if next_up:
                x = E.conv_layer(cache, f'dec{i}', dec.conv2d, kind, up, ops.EPI_BIAS_RELU_ADD, res=skip_of(i + 1),
                                 norm_mod=nm, norm_kind=nk, training=self.training, round_out=True)
                skip_added = True
            else:
                x = E.conv_layer(cache, f'dec{i}', dec.conv2d, kind, up, ops.EPI_BIAS_RELU,
                                 norm_mod=nm, norm_kind=nk, training=self.training)
        return E.pred_layer(x, pr.conv2d, getattr(pr, 'norm_layer', None), pr.norm, self.training, return_logits)
