"""Inference render (test views, 800x800 frames): the march -> field -> composite loop of nerf/renderer.py:759-794 with the loop
bookkeeping on the device.

The reference asks the host for the alive count every iteration (`rays_alive[rays_alive >= 0]`: a blocking sync, a mask kernel, a
scan and a gather) and sizes the next launches from it.  Here every launch has capacity N and reads (n_alive, n_step, M) from a
32-byte device state; compaction is one warp-aggregated scatter (csrc/raymarch.cu: k_compact_alive) whose last block advances the
state.  The host never blocks on the GPU inside a frame: it polls a pinned mirror of n_alive (written by the compaction kernel) to
stop enqueueing, and keeps at most `AHEAD` iterations queued so that a finished frame wastes only a few empty launches.
"""
import torch

from . import _lib
from .field import SHADING_ID

AHEAD = 24          # iterations the host may run ahead of the device


class _EvalWorkspace:
    def __init__(self, N, device):
        f = lambda *s: torch.empty(*s, device=device, dtype=torch.float32)
        self.N = N
        self.state = torch.zeros(8, device=device, dtype=torch.int32)
        self.host_alive = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.alive = [torch.empty(N, device=device, dtype=torch.int32) for _ in range(2)]
        self.rays_t, self.nears, self.fars = f(N), f(N), f(N)
        self.xyzs, self.dirs, self.ts = f(N, 3), f(N, 3), f(N, 2)          # n_alive * n_step <= N rows are live per iteration
        self.sigmas, self.colors, self.normals = f(N), f(N, 3), f(N, 3)
        self.bg = f(N, 3)


def render_eval(model, rays_o, rays_d, *, light_d=None, ambient_ratio=1.0, shading='albedo', bg_color=None, perturb=False, T_thresh=1e-4):
    """-> dict(image [N,3], depth [N], weights_sum [N]); no autograd."""
    opt = model.opt
    with torch.no_grad():
        rays_o = rays_o.detach().float().contiguous().view(-1, 3)
        rays_d = rays_d.detach().float().contiguous().view(-1, 3)
        N, dev = rays_o.shape[0], rays_o.device
        key = ('eval', N, dev)
        ws = model._ws.get(key)
        if ws is None:
            ws = model._ws[key] = _EvalWorkspace(N, dev)
        st = _lib.stream()
        p = _lib.ptr
        weights_sum = torch.empty(N, device=dev)
        depth = torch.empty(N, device=dev)
        image_c = torch.empty(N, 3, device=dev)
        _lib.call('sdf_near_far_from_aabb', p(rays_o), p(rays_d), p(model.aabb_infer), N, 0.2, p(ws.nears), p(ws.fars), st)
        if light_d is None:
            l = rays_o[0] + torch.randn(3, device=dev)
            light_d = l / torch.sqrt(torch.clamp((l * l).sum(), min=1e-20))
        light = light_d.detach().float().contiguous().view(-1, 3)
        if light.shape[0] != 1:
            raise RuntimeError('render_eval: one light direction per call (the reference test views pass a single light_d)')
        _lib.call('sdf_infer_begin', p(ws.state), N, int(opt.max_steps), p(ws.alive[0]), p(ws.rays_t), p(ws.nears), p(weights_sum), p(depth), p(image_c),
                  ws.host_alive.data_ptr(), st)
        sn = model.sigma_net.net
        c = model.field_cfg()
        table = model.table_half()
        sid = SHADING_ID[shading]
        wts = [p(t) for t in (sn[0].weight, sn[0].bias, sn[1].weight, sn[1].bias, sn[2].weight, sn[2].bias)]
        m_dev = ws.state.data_ptr() + 8            # &state.M
        noises = torch.rand(N, device=dev) if perturb else None
        events = []
        it, cur = 0, 0
        max_iters = int(opt.max_steps)             # n_step >= 1 per iteration, so the device-side step counter ends the loop by then
        while it < max_iters:
            _lib.call('sdf_infer_march', p(ws.state), N, p(ws.alive[cur]), p(ws.rays_t), p(rays_o), p(rays_d), float(model.bound), 0, float(opt.dt_gamma),
                      int(opt.max_steps), int(model.cascade), int(model.grid_size), p(model.density_bitfield), p(ws.fars), p(ws.xyzs), p(ws.dirs),
                      p(ws.ts), p(noises) if it == 0 else None, st)
            _lib.call('sdf_field_forward', p(ws.xyzs), N, m_dev, p(table), p(c['offsets']), c['L'], c['levels_active'], c['S'], int(c['H']),
                      int(c['smoothstep']), *wts, model.bound, c['blob_density'], c['blob_radius'], sid, p(light), 0, float(ambient_ratio),
                      p(ws.sigmas), p(ws.colors), None, None, None, st)
            _lib.call('sdf_infer_composite', p(ws.state), N, float(T_thresh), 0, p(ws.alive[cur]), p(ws.rays_t), p(ws.sigmas), p(ws.colors), p(ws.ts),
                      p(weights_sum), p(depth), p(image_c), st)
            _lib.call('sdf_infer_compact', p(ws.state), N, p(ws.alive[cur]), p(ws.alive[1 - cur]), ws.host_alive.data_ptr(), st)
            cur = 1 - cur
            it += 1
            ev = torch.cuda.Event()
            ev.record()
            events.append(ev)
            if len(events) > AHEAD:
                events.pop(0).synchronize()        # iteration it - AHEAD has finished: its alive count is in the pinned mirror
                if int(ws.host_alive[0]) == 0:
                    break
        # background + mix (nerf/renderer.py:796-808)
        image = torch.empty(N, 3, device=dev)
        bn = model.bg_net.net if model.bg_net is not None else None
        if bg_color is None and bn is None:
            bg_color = torch.ones(3, device=dev)
        if bg_color is None:
            bgw = [p(bn[0].weight), p(bn[0].bias), p(bn[1].weight), p(bn[1].bias)]
            bgc = None
        else:
            bgw = [None] * 4
            bgc = torch.as_tensor(bg_color, device=dev, dtype=torch.float32).expand(3).contiguous() if not torch.is_tensor(bg_color) else bg_color.float().contiguous()
        _lib.call('sdf_background_forward', p(rays_d), N, *bgw, p(bgc), int(model.half_round), p(image_c), p(weights_sum), None, p(image), None, 0, 3, st)
    return {'image': image, 'depth': depth, 'weights_sum': weights_sum, 'iterations': it}
