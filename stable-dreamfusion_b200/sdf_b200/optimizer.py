"""Adan with the same constructor / state-dict layout as the reference's optimizer.py:23-101 (Adan(params, lr, betas, eps,
weight_decay, max_grad_norm, no_prox, foreach)), executed by the fused kernels of csrc/adan.cu.

State per parameter: exp_avg, exp_avg_sq, exp_avg_diff, neg_pre_grad (reference names, optimizer.py:160-168); per group: step.
`loss_scale` folds the GradScaler unscale + inf-check + skip into the same kernels (no host synchronisation): the EXECUTED step
count lives on the device (`_steps_dev`, one int32 per group) and only advances when the gradients are finite, so a skipped step
leaves bias corrections and the first-step initialisation of neg_pre_grad alone — GradScaler.step() simply not calling
optimizer.step() (nerf/utils.py:1066).  group['step'] on the host counts calls; `sync_steps()` (used by state_dict) replaces it with
the device truth.  Optional extras riding in the same pass: an fp16 mirror of a parameter (`half_mirrors`), gradient zeroing
(also on skipped steps), and the torch_ema shadow update (`ema_attach` / step(ema=True), nerf/utils.py:282-283,1090-1091).
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
        self._steps_dev = None
        self.loss_scale = 1.0
        self.half_mirrors = {}          # id(param) -> fp16 tensor kept in sync by the step kernel
        self.ema_decay = None
        self.ema_num_updates = 0
        self.ema_shadow = {}            # id(param) -> fp32 shadow (torch_ema.ExponentialMovingAverage.shadow_params)

    # ------------------------------------------------------------------ device-side step counters
    def _device_state(self, dev):
        if self._acc is None or self._acc.device != dev:
            self._acc = torch.zeros(2, device=dev, dtype=torch.float32)
            self._steps_dev = torch.tensor([int(g.get('step', 0)) for g in self.param_groups], device=dev, dtype=torch.int32)

    def sync_steps(self):
        """group['step'] <- executed steps (one device read; not on the training path)"""
        if self._steps_dev is not None:
            for g, s in zip(self.param_groups, self._steps_dev.tolist()):
                g['step'] = int(s)

    def state_dict(self):
        self.sync_steps()
        sd = super().state_dict()
        for st in sd['state'].values():           # a tensor that never took a step has no neg_pre_grad in the reference's state either
            t = st.get('neg_pre_grad')
            if t is not None and t.numel() > 0 and bool(torch.isnan(t.reshape(-1)[0])):
                del st['neg_pre_grad']
        return sd

    def load_state_dict(self, sd):
        super().load_state_dict(sd)
        self._acc = None                 # counters are re-read from group['step'] at the next step

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
        if self._steps_dev is not None:
            self._steps_dev.zero_()

    # ------------------------------------------------------------------ EMA (torch_ema semantics)
    @torch.no_grad()
    def ema_attach(self, decay):
        self.ema_decay = float(decay)
        self.ema_num_updates = 0
        self.ema_shadow = {id(p): p.detach().clone() for g in self.param_groups for p in g['params']}

    def _ema_factor(self):
        """one-minus-decay of the NEXT update: decay = min(decay, (1 + n) / (10 + n)) with n counted from 1"""
        self.ema_num_updates += 1
        n = self.ema_num_updates
        return 1.0 - min(self.ema_decay, (1 + n) / (10 + n))

    @torch.no_grad()
    def ema_update(self):
        """stand-alone shadow update (what the reference does once per epoch, nerf/utils.py:1090-1091)"""
        omd = self._ema_factor()
        st = _lib.stream()
        for g in self.param_groups:
            for p in g['params']:
                _lib.call('sdf_ema_update', _lib.ptr(self.ema_shadow[id(p)]), _lib.ptr(p), p.numel(), omd, st)

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, zero_grad=False, ema=False):
        """ema=True folds this step's torch_ema update into the parameter pass (skipped, like the step, on non-finite gradients)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        plist = [(g, p) for g in self.param_groups for p in g['params'] if p.grad is not None]
        if not plist:
            return loss
        dev = plist[0][1].device
        self._device_state(dev)
        st = _lib.stream()
        inv_scale = 1.0 / float(self.loss_scale)
        _lib.call('sdf_adan_begin', _lib.ptr(self._acc), st)
        for _, p in plist:
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous() or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('fused Adan needs contiguous fp32 parameters and gradients')
            _lib.call('sdf_adan_grad_norm', _lib.ptr(g), g.numel(), inv_scale, _lib.ptr(self._acc), st)
        _lib.call('sdf_adan_advance', _lib.ptr(self._acc), _lib.ptr(self._steps_dev), len(self.param_groups), st)
        omd = self._ema_factor() if (ema and self.ema_decay is not None) else 0.0
        for gi, group in enumerate(self.param_groups):
            group['step'] = group.get('step', 0) + 1           # host view: calls (optimizer.py:191-194 increments every group)
            b1, b2, b3 = group['betas']
            step_dev = self._steps_dev.data_ptr() + 4 * gi
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state[p]
                if len(state) == 0:
                    state['exp_avg'] = torch.zeros_like(p)
                    state['exp_avg_sq'] = torch.zeros_like(p)
                    state['exp_avg_diff'] = torch.zeros_like(p)
                if 'neg_pre_grad' not in state:
                    state['neg_pre_grad'] = torch.full_like(p, float('nan'))      # 'never stepped' sentinel, see csrc/adan.cu:adan_one
                mirror = self.half_mirrors.get(id(p))
                shadow = self.ema_shadow.get(id(p)) if omd != 0.0 else None
                _lib.call('sdf_adan_step', _lib.ptr(p), _lib.ptr(p.grad), _lib.ptr(state['exp_avg']), _lib.ptr(state['exp_avg_diff']),
                          _lib.ptr(state['exp_avg_sq']), _lib.ptr(state['neg_pre_grad']), p.numel(), b1, b2, b3, int(group['step']), step_dev,
                          float(group['lr']), float(group['weight_decay']), float(group['eps']), float(self.defaults['max_grad_norm']),
                          int(bool(group['no_prox'])), inv_scale, _lib.ptr(self._acc), _lib.ptr(mirror), _lib.ptr(shadow), float(omd),
                          int(bool(zero_grad)), st)
        return loss
