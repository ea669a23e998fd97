"""Fused Adam over one flat fp32 buffer — a `torch.optim.Optimizer`, so it drops into the reference trainer where
`base/base_trainer.py:36-37` builds `getattr(optim, config['optimizer_type'])(model.parameters(), **config['optimizer'])`:
`param_groups` (lr schedulers, `base_trainer.py:38-43,119-122`), `state` (resume moves `optimizer.state.values()` to the
GPU, `:171-175`), `state_dict()` / `load_state_dict()` in torch.optim.Adam's own format (checkpoints interchange with
the reference's `optim.Adam`), `zero_grad()`, `step()`.

Parameters are re-homed into one contiguous buffer (each nn.Parameter becomes a view of it, so state_dict /
checkpoints / the CUDA graph see the same tensors), gradients accumulate into views of a second flat buffer, and
`step()` is ONE ramnet_adam_step launch per parameter group instead of ~10 kernels x 68 tensors.  Under data
parallelism the flat gradient buffer is summed across ranks first — in `n_buckets` slices on a side stream, each slice's
Adam launch waiting only for its own slice, so the tail of the all-reduce overlaps the optimiser.
"""
import torch

from . import engine, ops
from .distributed import BucketedAllReduce, dp_world_size


def _pad4(n):
    return ((n + 3) // 4) * 4


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 process_group=None, capturable=False, n_buckets=4):
        if amsgrad:
            raise ValueError('FusedAdam: amsgrad is not implemented (the reference configs do not use it)')
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False)
        super().__init__(params, defaults)
        for g in self.param_groups:
            g['params'] = [p for p in g['params'] if p.requires_grad]
        self.params = [p for g in self.param_groups for p in g['params']]
        if not self.params:
            raise ValueError('FusedAdam: no trainable parameters')
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise ops._lib.RamnetError('FusedAdam needs CUDA parameters (move the model to the GPU first)')
        self.process_group = process_group
        # 16-byte aligned segments so every parameter view keeps vector alignment
        total = sum(_pad4(p.numel()) for p in self.params)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_count = 0
        # capturable=True keeps the step number on the device so that step() can live inside a captured CUDA graph
        # (one counter per parameter group: every group's launch increments its own)
        self.capturable = capturable
        self.step_dev = torch.zeros(len(self.param_groups), dtype=torch.int32, device=dev)
        self._offsets, self._group_span = {}, []
        off = 0
        with torch.no_grad():
            for g in self.param_groups:
                g0 = off
                for p in g['params']:
                    n = p.numel()
                    view = self.flat_p[off:off + n].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
                    p.grad = self.flat_g[off:off + n].view(p.shape)
                    self._offsets[p] = off
                    self.state[p] = {'step': torch.zeros((), dtype=torch.float32),
                                     'exp_avg': self.exp_avg[off:off + n].view(p.shape),
                                     'exp_avg_sq': self.exp_avg_sq[off:off + n].view(p.shape)}
                    off += _pad4(n)
                self._group_span.append((g0, off))
        self._reducer = BucketedAllReduce(self.flat_g, n_buckets, process_group)
        engine.bump_weight_epoch()

    # ---- gradients ---------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=False):
        """Zeroes the flat gradient buffer (one memset); `set_to_none` is ignored: .grad stays a view of it."""
        self.flat_g.zero_()
        for p in self.params:            # re-attach the views if something replaced .grad
            off = self._offsets[p]
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)

    def _fold_foreign_grads(self):
        for p in self.params:            # autograd may have swapped in a fresh .grad tensor: fold it back
            g, off = p.grad, self._offsets[p]
            if g is not None and g.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                self.flat_g[off:off + p.numel()].view(p.shape).copy_(g)
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)

    # ---- step ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._fold_foreign_grads()
        self.step_count += 1
        single = len(self.param_groups) == 1
        if single and dp_world_size(self.process_group) > 1:
            # bucketed: slice k's Adam launch waits for slice k's all-reduce only
            g = self.param_groups[0]
            for lo, hi, ready in self._reducer.launch():
                ready()
                self._adam(g, 0, lo, hi, bump=(lo == 0))
        else:
            for lo, hi, ready in self._reducer.launch():
                ready()
            for gi, (g, (lo, hi)) in enumerate(zip(self.param_groups, self._group_span)):
                if hi > lo:
                    self._adam(g, gi, lo, hi, bump=True)
        engine.bump_weight_epoch()       # packed weights / captured graphs are stale now
        return loss

    def _adam(self, g, gi, lo, hi, bump):
        """One launch over flat[lo:hi].  Capturable mode: the device step counter of group gi is incremented by the
        launch that has bump=True and only read by the others of the same step."""
        kw = dict(lr=float(g['lr']), beta1=g['betas'][0], beta2=g['betas'][1], eps=g['eps'], weight_decay=g['weight_decay'])
        p, gr, m, v = self.flat_p[lo:hi], self.flat_g[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi]
        if self.capturable:
            ops.adam_step_dev(p, gr, m, v, self.step_dev[gi:gi + 1], increment=bump, **kw)
        else:
            ops.adam_step(p, gr, m, v, self.step_count, **kw)

    # ---- checkpoints (torch.optim.Adam format) -----------------------------------------------------------------
    def _current_step(self):
        return int(self.step_dev.max().item()) if self.capturable else self.step_count

    def state_dict(self):
        step = self._current_step()      # capturable: graph replays never run the Python step(), the device knows
        self.step_count = step
        for p in self.params:
            self.state[p]['step'] = torch.tensor(float(step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        if 'state' not in state_dict and 'exp_avg' in state_dict:      # round-1 flat format
            self.step_count = int(state_dict['step'])
            self.exp_avg.copy_(state_dict['exp_avg'])
            self.exp_avg_sq.copy_(state_dict['exp_avg_sq'])
        else:
            super().load_state_dict(state_dict)      # casts / copies the per-parameter state next to each parameter
            step = 0
            for p in self.params:
                st, off, n = self.state.get(p, {}), self._offsets[p], p.numel()
                if 'exp_avg' in st:
                    self.exp_avg[off:off + n].view(p.shape).copy_(st['exp_avg'])
                    self.exp_avg_sq[off:off + n].view(p.shape).copy_(st['exp_avg_sq'])
                    step = max(step, int(float(st.get('step', 0))))
                self.state[p] = {'step': torch.tensor(float(step)),
                                 'exp_avg': self.exp_avg[off:off + n].view(p.shape),
                                 'exp_avg_sq': self.exp_avg_sq[off:off + n].view(p.shape)}
            self.step_count = step
        self.step_dev.fill_(self.step_count)         # bias correction continues from the saved step after a resume
