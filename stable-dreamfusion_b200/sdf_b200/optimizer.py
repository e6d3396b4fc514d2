"""Adan with the same constructor / state-dict layout as the reference's optimizer.py:23-101 (Adan(params, lr, betas, eps,
weight_decay, max_grad_norm, no_prox, foreach)), executed by the fused kernels of csrc/adan.cu.

State per parameter: exp_avg, exp_avg_sq, exp_avg_diff, neg_pre_grad (reference names, optimizer.py:160-168); per group: step.
`loss_scale` folds the GradScaler unscale + inf-check + skip into the same kernels (no host synchronisation).
"""
import torch
from torch.optim.optimizer import Optimizer

from . import _lib


class Adan(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0, no_prox=False, foreach=False):
        if not 0.0 <= max_grad_norm:
            raise ValueError('Invalid Max grad norm: {}'.format(max_grad_norm))
        if not 0.0 <= lr:
            raise ValueError('Invalid learning rate: {}'.format(lr))
        if not 0.0 <= eps:
            raise ValueError('Invalid epsilon value: {}'.format(eps))
        for i in range(3):
            if not 0.0 <= betas[i] < 1.0:
                raise ValueError('Invalid beta parameter at index {}: {}'.format(i, betas[i]))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm, no_prox=no_prox, foreach=foreach)
        super().__init__(params, defaults)
        self._acc = None
        self.loss_scale = 1.0
        self.half_mirrors = {}          # id(param) -> fp16 tensor kept in sync by the step kernel

    @torch.no_grad()
    def restart_opt(self):
        for group in self.param_groups:
            group['step'] = 0
            for p in group['params']:
                if p.requires_grad:
                    state = self.state[p]
                    state['exp_avg'] = torch.zeros_like(p)
                    state['exp_avg_sq'] = torch.zeros_like(p)
                    state['exp_avg_diff'] = torch.zeros_like(p)

    @torch.no_grad()
    def step(self, closure=None, zero_grad=False):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plist = [(g, p) for g in self.param_groups for p in g['params'] if p.grad is not None]
        if not plist:
            return loss
        dev = plist[0][1].device
        if self._acc is None or self._acc.device != dev:
            self._acc = torch.zeros(2, device=dev, dtype=torch.float32)
        st = _lib.stream()
        inv_scale = 1.0 / float(self.loss_scale)
        _lib.call('sdf_adan_begin', _lib.ptr(self._acc), st)
        for _, p in plist:
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous() or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('fused Adan needs contiguous fp32 parameters and gradients')
            _lib.call('sdf_adan_grad_norm', _lib.ptr(g), g.numel(), inv_scale, _lib.ptr(self._acc), st)
        for group in self.param_groups:
            if not any(p.grad is not None for p in group['params']):
                continue
            group['step'] = group.get('step', 0) + 1
            b1, b2, b3 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state['exp_avg'] = torch.zeros_like(p)
                    state['exp_avg_sq'] = torch.zeros_like(p)
                    state['exp_avg_diff'] = torch.zeros_like(p)
                if 'neg_pre_grad' not in state:
                    state['neg_pre_grad'] = torch.zeros_like(p)
                mirror = self.half_mirrors.get(id(p))
                _lib.call('sdf_adan_step', _lib.ptr(p), _lib.ptr(p.grad), _lib.ptr(state['exp_avg']), _lib.ptr(state['exp_avg_diff']),
                          _lib.ptr(state['exp_avg_sq']), _lib.ptr(state['neg_pre_grad']), p.numel(), b1, b2, b3, int(group['step']), float(group['lr']),
                          float(group['weight_decay']), float(group['eps']), float(self.defaults['max_grad_norm']), int(bool(group['no_prox'])),
                          inv_scale, _lib.ptr(self._acc), _lib.ptr(mirror), int(bool(zero_grad)), st)
        return loss
