"""TorchAdan — TEST / BENCH INFRASTRUCTURE ONLY, never imported by the product.

The reference's Adan (optimizer.py:23-258, foreach=False: global-norm clip, then ~20 elementwise ops per parameter tensor) restated in
plain PyTorch for the CPU reference arm of bench.py (`--impl reference`, oracle/nerf_o2.py) and pinned to the reference optimizer's own
trajectory by tests/test_adan_golden.py.  (The GPU reference arm no longer restates anything: bench.py `--impl reference-cuda` runs the
reference's unmodified Trainer / renderer / network / optimizer through oracle/ref_harness.py.)
"""
import math

import torch


class TorchAdan(torch.optim.Optimizer):
    """optimizer.py:23-258 with foreach=False: global-norm clip, then ~20 elementwise PyTorch ops per parameter tensor"""

    def __init__(self, params, lr=1e-3, betas=(0.98, 0.92, 0.99), eps=1e-8, weight_decay=0.0, max_grad_norm=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm))
        self.loss_scale = 1.0

    @torch.no_grad()
    def step(self, zero_grad=False):
        mg = self.defaults["max_grad_norm"]
        clip = torch.tensor(1.0, device=self.param_groups[0]["params"][0].device)
        if mg > 0:
            gn = torch.zeros(1, device=clip.device)
            for g in self.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        gn.add_(p.grad.pow(2).sum())
            gn = torch.sqrt(gn)
            clip = torch.clamp(mg / (gn + self.defaults["eps"]), max=1.0)
        for group in self.param_groups:
            b1, b2, b3 = group["betas"]
            group["step"] = group.get("step", 0) + 1
            bc1, bc2, bc3s = 1.0 - b1 ** group["step"], 1.0 - b2 ** group["step"], math.sqrt(1.0 - b3 ** group["step"])
            lr, wd, eps = group["lr"], group["weight_decay"], group["eps"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    st["exp_avg_diff"] = torch.zeros_like(p)
                grad = p.grad.mul(clip)
                if "neg_pre_grad" not in st or group["step"] == 1:
                    st["neg_pre_grad"] = grad.clone().mul_(-1.0)
                npg = st["neg_pre_grad"]
                npg.add_(grad)                                          # diff
                st["exp_avg"].mul_(b1).add_(grad, alpha=1 - b1)
                st["exp_avg_diff"].mul_(b2).add_(npg, alpha=1 - b2)
                npg.mul_(b2).add_(grad)                                  # update
                st["exp_avg_sq"].mul_(b3).addcmul_(npg, npg, value=1 - b3)
                denom = (st["exp_avg_sq"].sqrt() / bc3s).add_(eps)
                p.addcdiv_(st["exp_avg"], denom, value=-lr / bc1)
                p.addcdiv_(st["exp_avg_diff"], denom, value=-lr * b2 / bc2)
                p.div_(1 + lr * wd)
                npg.zero_().add_(grad, alpha=-1.0)
        if zero_grad:
            for g in self.param_groups:
                for p in g["params"]:
                    p.grad = None
