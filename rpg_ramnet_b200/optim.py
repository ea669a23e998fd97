"""Fused Adam over one flat fp32 buffer (mirrors torch.optim.Adam as the reference builds it at
base/base_trainer.py:36-37: `getattr(optim, config['optimizer_type'])(params, **config['optimizer'])`).

Parameters are re-homed into one contiguous buffer (each nn.Parameter becomes a view of it, so
state_dict / checkpoints / the CUDA graph see the same tensors), gradients accumulate into views of a
second flat buffer, and `step()` is ONE ramnet_adam_step launch instead of ~10 kernels x 68 tensors.
Under data parallelism `step()` first sums the flat gradient buffer across ranks (one all-reduce of
59.5 MB for the shipped block).
"""
import torch

from . import engine, ops
from .distributed import all_reduce_flat_grads


class FusedAdam:
    def __init__(self, params, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 capturable=False):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError('FusedAdam: no trainable parameters')
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise ops._lib.RamnetError('FusedAdam needs CUDA parameters (move the model to the GPU first)')
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.process_group = process_group
        # 16-byte aligned segments so every parameter view keeps vector alignment
        sizes = [((p.numel() + 3) // 4) * 4 for p in self.params]
        total = sum(sizes)
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_count = 0
        # capturable=True keeps the step number on the device so that step() can live inside a captured CUDA graph
        self.capturable = capturable
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        off = 0
        with torch.no_grad():
            for p, sz in zip(self.params, sizes):
                view = self.flat_p[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
                off += sz
        engine.bump_weight_epoch()

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        off = 0
        for p in self.params:            # re-attach the views if something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
            off += ((p.numel() + 3) // 4) * 4

    def step(self):
        off = 0
        for p in self.params:            # autograd may have swapped in a fresh .grad tensor: fold it back
            g = p.grad
            if g is not None and g.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                self.flat_g[off:off + p.numel()].view(p.shape).copy_(g)
                p.grad = self.flat_g[off:off + p.numel()].view(p.shape)
            off += ((p.numel() + 3) // 4) * 4
        all_reduce_flat_grads(self.flat_g, self.process_group)
        self.step_count += 1
        if self.capturable:
            ops.adam_step_dev(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.step_dev, lr=self.lr,
                              beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay)
        else:
            ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.step_count, lr=self.lr,
                          beta1=self.betas[0], beta2=self.betas[1], eps=self.eps, weight_decay=self.weight_decay)
        engine.bump_weight_epoch()       # packed weights / captured graphs are stale now

    def state_dict(self):
        return {'step': self.step_count, 'exp_avg': self.exp_avg, 'exp_avg_sq': self.exp_avg_sq,
                'lr': self.lr, 'betas': self.betas, 'eps': self.eps, 'weight_decay': self.weight_decay}

    def load_state_dict(self, sd):
        self.step_count = int(sd['step'])
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])
