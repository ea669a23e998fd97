#!/usr/bin/env python
"""Target for compute-sanitizer (racecheck / memcheck / synccheck): BASELINE config 1 (128x128, batch 1) forward, two
timesteps, plus one backward + fused Adam step, in the default TF32 mode — every tcgen05 / TMA / mbarrier kernel of the
path runs at least once.  RAMNET_PAIR=2 forces the cta_group::2 pair mode wherever a configuration fits, RAMNET_PAIR=0
turns it off (tools/r2_sanitize.sh runs both).  Checks the depth maps against the oracle so that a sanitizer run that
"passes" on garbage is caught."""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rpg_ramnet_b200 as R  # noqa: E402
from oracle import ramnet_oracle as O  # noqa: E402

HW = int(os.environ.get('SANITIZE_HW', '128'))
cfg = dict(num_bins_rgb=1, num_bins_events=5, skip_type='sum', recurrent_block_type='conv', state_combination='convgru',
           num_encoders=3, base_num_channels=32, num_residual_blocks=2, use_upsample_conv=True, norm='none',
           every_x_rgb_frame=1, gpu=0)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = R.ERGB2DepthRecurrent(cfg)
model.eval().to('cuda:0')
seq = O.synth_sequence(1, HW, HW, 2, 1, seed=1, with_targets=True)
sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
worst = 0.0
ps, pl, os_, ol = None, {'events0': None, 'image': None}, None, {'events0': None, 'image': None}
with torch.no_grad():
    for item in seq:
        preds, supers, lstm = model(item, ps, pl)
        o_preds, o_supers, o_lstm = O.ergb2depth_recurrent(sd, cfg, item, os_, ol)
        for k in o_preds:
            worst = max(worst, float(((preds[k].cpu() - o_preds[k]).abs() / o_preds[k].abs()).max()))
        ps, pl, os_, ol = supers['image'], lstm, o_supers['image'], o_lstm
assert worst <= 1e-3, worst
# one training step (backward kernels: dgrad, tap-packed wgrad, gate adjoints, loss, Adam)
model.train()
opt = R.FusedAdam(model.parameters(), lr=3e-4)
opt.zero_grad()
ps, pl, terms = None, {'events0': None, 'image': None}, []
for item in seq[:1]:
    preds, supers, lstm = model(item, ps, pl)
    terms += [R.scale_invariant_loss(preds[k], item['depth_' + k].to('cuda:0')) for k in preds]
    terms += [0.25 * R.multi_scale_grad_loss(preds[k], item['depth_' + k].to('cuda:0')) for k in preds]
loss = sum(terms)
loss.backward()
opt.step()
ev = O.synth_events(20000, 64, 48, seed=2)
grid = R.events_to_voxel_grid(ev, 5, 64, 48)           # fused zero-fill + votes (cooperative launch)
# live norm layers (train-mode BatchNorm statistics, forward + backward), TransposedConvLayer decoders, ConvLSTM state
ncfg = dict(cfg, norm='BN', use_upsample_conv=False, state_combination='convlstm')
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    nmodel = R.ERGB2DepthRecurrent(ncfg)
nmodel.train().to('cuda:0')
nseq = O.synth_sequence(2, 64, 64, 1, 1, seed=3, with_targets=True)
npreds, _, _ = nmodel(nseq[0], None, {'events0': None, 'image': None})
nloss = sum(R.scale_invariant_loss(npreds[k], nseq[0]['depth_' + k].to('cuda:0')) for k in npreds)
nloss.backward()
# UNet baseline with the concatenated skip (two-source decoder convs, split-weight pred) and InstanceNorm
ucfg = dict(cfg, num_bins_rgb=6, skip_type='concat', norm='IN')
with contextlib.redirect_stdout(io.StringIO()):
    umodel = R.ERGB2Depth(ucfg)
umodel.train().to('cuda:0')
upred, _, _ = umodel({'image': torch.rand(1, 6, 64, 64)}, None, None)
upred['image'].mean().backward()
torch.cuda.synchronize()
print(f'sanitize target ok: fwd max rel err {worst:.2e}, loss {loss.item():.6f}, norm-model loss {nloss.item():.6f}, launches {R.launch_count(0)}, '
      f'RAMNET_PAIR={os.environ.get("RAMNET_PAIR", "default")}')
