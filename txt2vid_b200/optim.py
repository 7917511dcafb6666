"""Fused multi-tensor Adam on the sm_100a kernel (t2v_adam_step); drop-in for the
torch.optim.Adam the reference builds at txt2vid/train/gan.py:93-94 (eps 1e-8, no weight decay)."""
import torch

from . import kernels as K
from . import ops


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = 1.0
        self.dyn = None          # device {lr/(1-b1^t), 1/sqrt(1-b2^t)} in CUDA-graph mode (trainer.GraphedTrainStep)

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            ps, gs, ms, vs = [], [], [], []
            step = None
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['step'] += 1
                step = st['step']
                g = p.grad
                if g.stride() != p.stride():
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
                    p.grad = g
                if st['exp_avg'].stride() != p.stride():      # parameter was re-homed after state creation
                    for k in ('exp_avg', 'exp_avg_sq'):
                        st[k] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(st[k])
                ps.append(p)
                gs.append(g)
                ms.append(st['exp_avg'])
                vs.append(st['exp_avg_sq'])
            if ps:
                b1, b2 = group['betas']
                K.adam_step(ps, gs, ms, vs, group['lr'], b1, b2, group['eps'], step, self.grad_scale, self.dyn)
        ops.bump_weight_epoch()
