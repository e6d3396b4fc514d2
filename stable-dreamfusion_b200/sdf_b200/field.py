"""Fused radiance field of the -O backbone: autograd front-end of csrc/fused_field{,_bwd}.cu.

fused_field(...) computes what NeRFNetwork.forward (nerf/network_grid.py:104-130) computes —
sigma, shaded colour and the finite-difference normal — in one kernel, and back-propagates
into the hash table and sigma_net in one kernel.
"""
import numpy as np
import torch
from torch.autograd import Function

from . import _lib

SHADING_ID = {'albedo': 0, 'lambertian': 1, 'textureless': 2, 'normal': 3}
AUX_STRIDE = 10

# When set (by the trainer), the backward kernel scatters straight into the parameters' existing .grad buffers
# (fp32, contiguous) and returns None to autograd: no 48.8 MB zero-fill, no accumulate pass.  Off by default so the
# function is a plain differentiable op.
DIRECT_GRAD_ACCUM = False

def half_table(embeddings):
    """fp16 working copy of the fp32 hash table.  Cast once per fused call (the reference casts it once per
    encoder call, i.e. 7x per shaded step, gridencoder/grid.py:46-47); 73 MB of traffic, ~11 us on B200.
    A cache keyed on (data_ptr, _version) is NOT safe: `.data` writes do not bump the version counter."""
    return embeddings.detach().to(torch.half)


def _f32c(t):
    t = t.detach()
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


class _FusedField(Function):
    @staticmethod
    def forward(ctx, xyzs, embeddings, w1, b1, w2, b2, w3, b3, offsets, light_d, cfg):
        _lib.require_cuda(xyzs, embeddings, offsets)
        xyzs = _f32c(xyzs)
        M = xyzs.shape[0]
        dev = xyzs.device
        table = cfg.get('table_half')
        if table is None:
            table = half_table(embeddings) if embeddings.dtype == torch.float32 else embeddings.detach()
        ws = [_f32c(t) for t in (w1, b1, w2, b2, w3, b3)]
        shading = SHADING_ID[cfg['shading']]
        if light_d is None:
            light = None
            per_sample = 0
        else:
            light = _f32c(light_d).view(-1, 3)
            per_sample = 1 if light.shape[0] > 1 else 0
            if per_sample and light.shape[0] != M:
                raise RuntimeError('fused_field: light_d must be [3], [1,3] or [M,3]')
        need_grad = cfg.get('train', True)
        sig = torch.empty(M, device=dev, dtype=torch.float32)
        col = torch.empty(M, 3, device=dev, dtype=torch.float32) if cfg.get('want_color', True) else None
        nrm = torch.empty(M, 3, device=dev, dtype=torch.float32) if (shading != 0 and cfg.get('want_color', True)) else None
        aux = torch.empty(M, AUX_STRIDE, device=dev, dtype=torch.float32) if need_grad else None
        L = offsets.shape[0] - 1
        args = (_lib.ptr(xyzs), M, None, _lib.ptr(table), _lib.ptr(offsets), L, int(cfg['levels_active']), float(cfg['S']), int(cfg['H']),
                int(cfg['smoothstep']), *[_lib.ptr(t) for t in ws], float(cfg['bound']), float(cfg['blob_density']),
                float(cfg['blob_radius']), shading, _lib.ptr(light), per_sample, float(cfg['ratio']))
        feat = None
        if need_grad and M > 0:
            nbytes = _lib.query('sdf_field_feat_bytes', M, shading)
            if nbytes <= _lib.feat_stash_budget_bytes():
                feat = torch.empty(nbytes // 4, device=dev, dtype=torch.int32)
        _lib.call('sdf_field_forward', *args, _lib.ptr(sig), _lib.ptr(col), _lib.ptr(nrm), _lib.ptr(aux), _lib.ptr(feat), _lib.stream())
        if need_grad:
            ctx.feat = feat
            ctx.save_for_backward(xyzs, embeddings, table, offsets, light, aux, *ws)
            ctx.cfg = {k: v for k, v in cfg.items() if k != 'table_half'}
            ctx.per_sample = per_sample
            ps = (embeddings, w1, b1, w2, b2, w3, b3)
            ctx.direct_params = ps if all(isinstance(p, torch.nn.Parameter) and p.dtype == torch.float32 for p in ps) else None
        return sig, col, nrm

    @staticmethod
    def backward(ctx, g_sig, g_col, g_nrm):
        xyzs, embeddings, table, offsets, light, aux, w1, b1, w2, b2, w3, b3 = ctx.saved_tensors
        cfg = ctx.cfg
        M = xyzs.shape[0]
        dev = xyzs.device
        f = lambda g: None if g is None else g.float().contiguous()
        g_sig, g_col, g_nrm = f(g_sig), f(g_col), f(g_nrm)
        direct = DIRECT_GRAD_ACCUM and ctx.direct_params is not None and all(
            p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous() for p in ctx.direct_params)
        if direct:
            g_table = ctx.direct_params[0].grad
            gws = [p.grad for p in ctx.direct_params[1:]]
        else:
            g_table = torch.zeros(embeddings.shape, device=dev, dtype=torch.float32)
            gws = [torch.zeros_like(t) for t in (w1, b1, w2, b2, w3, b3)]
        shading = SHADING_ID[cfg['shading']]
        L = offsets.shape[0] - 1
        _lib.call('sdf_field_backward', _lib.ptr(xyzs), M, None, _lib.ptr(table), _lib.ptr(offsets), L, int(cfg['levels_active']),
                  float(cfg['S']), int(cfg['H']), int(cfg['smoothstep']), *[_lib.ptr(t) for t in (w1, b1, w2, b2, w3, b3)],
                  float(cfg['bound']), float(cfg['blob_density']), float(cfg['blob_radius']), shading, _lib.ptr(light), ctx.per_sample,
                  float(cfg['ratio']), _lib.ptr(aux), _lib.ptr(g_sig), _lib.ptr(g_col), _lib.ptr(g_nrm), _lib.ptr(g_table),
                  *[_lib.ptr(t) for t in gws], _lib.ptr(ctx.feat), _lib.stream())
        if direct:
            return (None,) * 11
        if embeddings.dtype != torch.float32:
            g_table = g_table.to(embeddings.dtype)
        return (None, g_table, *gws, None, None, None)


def fused_field(xyzs, embeddings, w1, b1, w2, b2, w3, b3, offsets, light_d, *, shading='albedo', ratio=1.0, bound=1.0,
                per_level_scale=2.0, base_resolution=16, smoothstep=True, levels_active=None, blob_density=5.0, blob_radius=0.2,
                want_color=True, table_half=None):
    """-> (sigma [M], color [M,3] | None, normal [M,3] | None); differentiable wrt embeddings and the six MLP tensors."""
    L = offsets.shape[0] - 1
    train = torch.is_grad_enabled() and any(t.requires_grad for t in (embeddings, w1, b1, w2, b2, w3, b3))
    cfg = dict(shading=shading, ratio=ratio, bound=bound, S=float(np.log2(per_level_scale)), H=base_resolution, smoothstep=bool(smoothstep),
               levels_active=L if levels_active is None else levels_active, blob_density=blob_density, blob_radius=blob_radius,
               want_color=want_color, train=train, table_half=table_half)
    return _FusedField.apply(xyzs, embeddings, w1, b1, w2, b2, w3, b3, offsets, light_d, cfg)
