"""The training render of the -O preset as ONE differentiable op with a device-side sample count.

What the reference spreads over nerf/renderer.py:710-757 + :796-808 (run_cuda, training branch), nerf/network_grid.py:104-147 and the
regulariser lines of nerf/utils.py:686-704 — near/far, two marching passes around a blocking `.item()`, `safe_normalize`, 7 hash-grid
encodes + MLPs, compositing, the background net, the mix, the permute to NCHW and the entropy / orientation means — is one
autograd.Function here, built from capacity-sized launches:

    near/far -> march (count | device scan | write)  -> fused field (m_dev)  -> composite  -> background + mix + NCHW  -> regularisers
                        M stays in device memory (a pinned mirror is written for logging only)

Nothing between the pose upload and the optimiser step synchronises the host, every launch has static shapes (capacity =
rays x max_steps, the march's hard upper bound, so no overflow handling exists), and the backward is the same chain reversed:
background/mix -> regularisers -> composite -> fused field backward scattering straight into the parameters' .grad buffers.

Random draws happen in the reference's order (light direction randn(3) per call, then the march jitter rand(N)), so a run seeded like
the reference's sees the same numbers (tests/test_gpu_dropin_reference.py).
"""
import torch
from torch.autograd import Function

from . import _lib
from .field import AUX_STRIDE, SHADING_ID


class RenderWorkspace:
    """Capacity-sized device buffers of one render configuration (N rays): allocated once, reused by every step."""

    def __init__(self, N, max_steps, device):
        self.N, self.cap, self.device = N, N * max_steps, device
        f = lambda *s: torch.empty(*s, device=device, dtype=torch.float32)
        cap = self.cap
        self.nears, self.fars = f(N), f(N)
        self.rays = torch.empty(N, 2, device=device, dtype=torch.int32)
        self.counter = torch.zeros(1, device=device, dtype=torch.int32)
        self.host_M = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.xyzs, self.dirs, self.ts = f(cap, 3), f(cap, 3), f(cap, 2)
        self.sigmas, self.colors, self.normals, self.aux = f(cap), f(cap, 3), f(cap, 3), f(cap, AUX_STRIDE)
        self.weights = f(cap)
        # feature stash of the fused field (7 x 64 B per sample): the backward streams it back instead of re-gathering the hash grid
        nbytes = _lib.query('sdf_field_feat_bytes', cap, 1)
        self.feat = torch.empty(nbytes // 4, device=device, dtype=torch.int32) if nbytes <= _lib.feat_stash_budget_bytes() else None
        self.weights_sum, self.depth, self.image_c, self.bg = f(N), f(N), f(N, 3), f(N, 3)
        self.light = None                      # [cap, 3], allocated on first per-sample-light use
        self.reg_scratch, self.reg_out = torch.zeros(3, device=device), torch.zeros(2, device=device)
        # backward
        self.g_image_c, self.g_ws = f(N, 3), f(N)
        self.g_weights, self.g_normals = f(cap), f(cap, 3)
        self.g_sigmas, self.g_colors = f(cap), f(cap, 3)
        self.generation = 0

    def light_buffer(self):
        if self.light is None:
            self.light = torch.empty(self.cap, 3, device=self.device, dtype=torch.float32)
        return self.light


def _p(t):
    return _lib.ptr(t)


class _RenderTrain(Function):
    """inputs: the differentiable parameters (listed so autograd routes gradients), rays, and a config dict"""

    @staticmethod
    def forward(ctx, rays_o, rays_d, light, bg_color, table, w1, b1, w2, b2, w3, b3, bw1, bb1, bw2, bb2, cfg):
        ws = cfg['ws']
        ws.generation += 1
        ctx.set_materialize_grads(False)          # unused outputs arrive as None in backward instead of zero tensors
        N, cap = ws.N, ws.cap
        st = _lib.stream()
        m = cfg['march']
        rays_o = rays_o.detach().float().contiguous().view(-1, 3)
        rays_d = rays_d.detach().float().contiguous().view(-1, 3)
        assert rays_o.shape[0] == N
        _lib.call('sdf_near_far_from_aabb', _p(rays_o), _p(rays_d), _p(cfg['aabb']), N, 0.2, _p(ws.nears), _p(ws.fars), st)
        noises = torch.rand(N, dtype=torch.float32, device=rays_o.device) if cfg['perturb'] else None
        margs = (_p(rays_o), _p(rays_d), _p(cfg['bitfield']), float(m['bound']), 0, float(m['dt_gamma']), int(m['max_steps']), N,
                 int(m['cascade']), int(m['grid_size']), _p(ws.nears), _p(ws.fars), _p(noises))
        _lib.call('sdf_march_rays_train_count', *margs, _p(ws.rays), _p(ws.counter), ws.host_M.data_ptr(), st)
        _lib.call('sdf_march_rays_train_write', *margs, _p(ws.xyzs), _p(ws.dirs), _p(ws.ts), _p(ws.rays), cap, st)
        # light: [3] shared, or one direction per ray expanded to its samples
        shading = SHADING_ID[cfg['shading']]
        per_sample = 0
        light = light.detach().float().contiguous().view(-1, 3)
        if light.shape[0] > 1:
            assert light.shape[0] == N
            lbuf = ws.light_buffer()
            _lib.call('sdf_expand_ray_vec3', _p(light), _p(ws.rays), N, cap, _p(lbuf), st)
            light, per_sample = lbuf, 1
        f = cfg['field']
        ws_list = [t.detach() for t in (w1, b1, w2, b2, w3, b3)]
        fargs = (_p(ws.xyzs), cap, _p(ws.counter), _p(cfg['table_half']), _p(f['offsets']), int(f['L']), int(f['levels_active']), float(f['S']),
                 int(f['H']), int(f['smoothstep']), *[_p(t) for t in ws_list], float(m['bound']), float(f['blob_density']), float(f['blob_radius']),
                 shading, _p(light), per_sample, float(cfg['ratio']))
        need_n = shading != 0
        _lib.call('sdf_field_forward', *fargs, _p(ws.sigmas), _p(ws.colors), _p(ws.normals) if need_n else None, _p(ws.aux), _p(ws.feat), st)
        _lib.call('sdf_composite_rays_train_forward', _p(ws.sigmas), _p(ws.colors), _p(ws.ts), _p(ws.rays), cap, N, float(cfg['T_thresh']), 0,
                  _p(ws.weights), _p(ws.weights_sum), _p(ws.depth), _p(ws.image_c), st)
        B, HW, C = cfg['B'], N // cfg['B'], cfg['C']
        dev = rays_o.device
        image = torch.empty(N, 3, device=dev, dtype=torch.float32)
        pred = torch.empty(B, C, HW, device=dev, dtype=torch.float32)
        use_net = bg_color is None
        bgc = None if use_net else bg_color.detach().float().contiguous()
        bgw = [t.detach() for t in (bw1, bb1, bw2, bb2)] if use_net else [None] * 4
        _lib.call('sdf_background_forward', _p(rays_d), N, *[_p(t) for t in bgw], _p(bgc), int(cfg['half_round']), _p(ws.image_c), _p(ws.weights_sum),
                  _p(ws.bg), _p(image), _p(pred), HW, C, st)
        want_orient = need_n and cfg['lambda_orient'] > 0
        _lib.call('sdf_render_regularizers_forward', _p(ws.weights), _p(ws.normals) if want_orient else None, _p(ws.dirs), _p(ws.counter), cap,
                  _p(ws.reg_scratch), _p(ws.reg_out), st)
        reg_terms = ws.reg_out.clone()
        ctx.cfg, ctx.gen = cfg, ws.generation
        ctx.fargs, ctx.shading, ctx.want_orient, ctx.use_net = fargs, shading, want_orient, use_net
        ctx.keep = (rays_o, rays_d, light, bgc, noises)              # keeps the pointers inside fargs alive
        ctx.params = (table, w1, b1, w2, b2, w3, b3, bw1, bb1, bw2, bb2)
        depth = ws.depth.clone()
        ctx.mark_non_differentiable(depth)
        return pred.view(B, C, cfg['H'], cfg['W']), image, ws.weights_sum.clone(), reg_terms, depth

    @staticmethod
    def backward(ctx, g_pred, g_image, g_wsum, g_reg, g_depth):
        cfg = ctx.cfg
        ws = cfg['ws']
        if ws.generation != ctx.gen:
            raise RuntimeError('render_train: the workspace was reused by a later render before this backward ran')
        N, cap = ws.N, ws.cap
        st = _lib.stream()
        rays_o, rays_d, light, bgc, _ = ctx.keep
        table, w1, b1, w2, b2, w3, b3, bw1, bb1, bw2, bb2 = ctx.params
        B, HW, C = cfg['B'], N // cfg['B'], cfg['C']
        fc = lambda g: None if g is None else g.float().contiguous()
        g_pred, g_image, g_wsum, g_reg = fc(g_pred), fc(g_image), fc(g_wsum), fc(g_reg)
        if g_pred is None and g_image is None:
            g_image = torch.zeros(N, 3, device=rays_o.device)
        direct = cfg['direct_grads']

        def grad_buf(p):
            if p is None:
                return None
            if direct:
                if p.grad is None:
                    p.grad = torch.zeros_like(p, dtype=torch.float32)
                return p.grad
            return torch.zeros_like(p, dtype=torch.float32)
        gb = [grad_buf(p) if ctx.use_net else None for p in (bw1, bb1, bw2, bb2)]
        bgw = [t.detach() for t in (bw1, bb1, bw2, bb2)] if ctx.use_net else [None] * 4
        _lib.call('sdf_background_backward', _p(g_image), _p(g_pred), HW, C, _p(rays_d), N, *[_p(t) for t in bgw], _p(bgc), int(cfg['half_round']),
                  _p(ws.weights_sum), _p(ws.g_image_c), _p(ws.g_ws), *[_p(t) for t in gb], st)
        if g_wsum is not None:
            ws.g_ws.add_(g_wsum.view(-1))
        # regularisers: g_reg = (d loss / d mean-entropy, d loss / d mean-orientation), read on the device
        have_reg = g_reg is not None and (cfg['lambda_entropy'] > 0 or ctx.want_orient)
        if have_reg:
            _lib.call('sdf_render_regularizers_backward', _p(g_reg), 1.0, 1.0, _p(ws.weights), _p(ws.normals) if ctx.want_orient else None,
                      _p(ws.dirs), _p(ws.counter), cap, _p(ws.g_weights), _p(ws.g_normals) if ctx.want_orient else None, st)
        _lib.call('sdf_composite_rays_train_backward', _p(ws.g_weights) if have_reg else None, _p(ws.g_ws), None, _p(ws.g_image_c), _p(ws.sigmas),
                  _p(ws.colors), _p(ws.ts), _p(ws.rays), _p(ws.weights_sum), _p(ws.depth), _p(ws.image_c), cap, N, float(cfg['T_thresh']), 0,
                  _p(ws.g_sigmas), _p(ws.g_colors), st)
        gf = [grad_buf(p) for p in (table, w1, b1, w2, b2, w3, b3)]
        _lib.call('sdf_field_backward', *ctx.fargs, _p(ws.aux), _p(ws.g_sigmas), _p(ws.g_colors),
                  _p(ws.g_normals) if (have_reg and ctx.want_orient) else None, *[_p(t) for t in gf], _p(ws.feat), st)
        if direct:
            return (None,) * 16
        return (None, None, None, None, *gf, *gb, None)


def render_train(model, rays_o, rays_d, *, light_d=None, ambient_ratio=1.0, shading='albedo', bg_color=None, perturb=True, T_thresh=1e-4,
                 as_latent=False, B=1, H=None, W=None, direct_grads=False):
    """-> dict(pred_rgb [B,C,H,W], image [N,3], weights_sum [N], depth [N], loss_entropy, loss_orient, reg (lambda-weighted sum))

    model: sdf_b200.ngp.InstantNGP.  rays_*: [B*H*W, 3] (any leading shape).  light_d: None (the reference's rays_o + randn(3) draw),
    [3] or one row per ray.  bg_color: None -> background net, else a 3-colour.  Regulariser weights come from model.opt."""
    opt = model.opt
    rays_o = rays_o.reshape(-1, 3)
    rays_d = rays_d.reshape(-1, 3)
    N = rays_o.shape[0]
    dev = rays_o.device
    if light_d is None:
        # nerf/renderer.py:726-727: one randn(3) offset for the whole call, normalised per ray (rows differ only between views)
        off = torch.randn(3, device=dev, dtype=torch.float32)
        rows = rays_o[:: max(N // B, 1)][:B] if B > 1 else rays_o[:1]
        l = rows.float() + off
        l = l / torch.sqrt(torch.clamp((l * l).sum(-1, keepdim=True), min=1e-20))
        light_d = l if B == 1 else l.repeat_interleave(N // B, dim=0)
    ws = model.workspace(N)
    sn = model.sigma_net.net
    bn = model.bg_net.net if model.bg_net is not None else None
    if bg_color is None and bn is None:
        bg_color = torch.ones(3, device=dev)                          # nerf/renderer.py:803: no bg net -> white
    lam_e = float(opt.lambda_entropy) * model.entropy_ramp
    lam_o = float(opt.lambda_orient)
    cfg = dict(ws=ws, aabb=model.aabb_train, bitfield=model.density_bitfield, perturb=bool(perturb), shading=shading, ratio=float(ambient_ratio),
               T_thresh=float(T_thresh), B=int(B), H=int(H), W=int(W), C=4 if as_latent else 3, half_round=int(model.half_round),
               lambda_entropy=lam_e, lambda_orient=lam_o, direct_grads=bool(direct_grads),
               table_half=model.table_half(), march=dict(bound=model.bound, dt_gamma=opt.dt_gamma, max_steps=opt.max_steps, cascade=model.cascade,
                                                         grid_size=model.grid_size), field=model.field_cfg())
    none = lambda m, i, a: getattr(m[i], a) if m is not None else None
    pred, image, wsum, reg_terms, depth = _RenderTrain.apply(
        rays_o, rays_d, light_d, bg_color, model.encoder.embeddings, sn[0].weight, sn[0].bias, sn[1].weight, sn[1].bias, sn[2].weight, sn[2].bias,
        none(bn, 0, 'weight'), none(bn, 0, 'bias'), none(bn, 1, 'weight'), none(bn, 1, 'bias'), cfg)
    out = {'pred_rgb': pred, 'image': image, 'weights_sum': wsum, 'depth': depth, 'loss_entropy': reg_terms[0], 'loss_orient': reg_terms[1]}
    out['reg'] = lam_e * reg_terms[0] + lam_o * reg_terms[1]
    out['weights'] = ws.weights            # [capacity]; rows >= M are undefined (M: ws.counter on the device, ws.host_M mirror)
    return out
