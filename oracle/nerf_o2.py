"""Pure-PyTorch restatement of the reference's `-O2` (vanilla backbone, no CUDA extension) NeRF path — the CPU baseline
of BASELINE.json config C1 (32x32 render, 1 view/step).

TEST INFRASTRUCTURE ONLY (oracle, kind "port"): used by bench.py's cpu_baseline / `--impl reference` arm; never imported by the
product package.  Pinned: tests/golden/nerf_o2.npz holds outputs of the reference's own nerf/network.py + nerf/renderer.py:run on
the same weights / rays / seeds (generator tests/golden/make_golden_o2.py, max |diff| 0.0); tests/test_oracle_o2_golden.py replays it.  Restates:
  nerf/network.py:11-116   ResBlock / BasicBlock / MLP / NeRFNetwork (frequency_torch(12) -> 5-layer ResBlock MLP(64) -> 4;
                           background frequency_torch(4) -> MLP 27->32->3), normals by autograd (:181-186)
  encoding.py:5-52         FreqEncoder_torch (log-sampled sin/cos, input included)
  nerf/renderer.py:56-67   near_far_from_bound(type='sphere')      :19-53  sample_pdf
  nerf/renderer.py:560-707 run(): 64 uniform + 32 importance samples per ray, cumprod compositing, orientation loss
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def safe_normalize(x, eps=1e-20):
    return x / torch.sqrt(torch.clamp(torch.sum(x * x, -1, keepdim=True), min=eps))


class FreqEncoderTorch(nn.Module):
    def __init__(self, input_dim, n_freqs):
        super().__init__()
        self.freqs = (2 ** torch.linspace(0, n_freqs - 1, n_freqs)).tolist()
        self.output_dim = input_dim + input_dim * n_freqs * 2

    def forward(self, x):
        out = [x]
        for f in self.freqs:
            out += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(out, dim=-1)


class ResBlock(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.dense = nn.Linear(din, dout)
        self.norm = nn.LayerNorm(dout)
        self.skip = nn.Linear(din, dout, bias=False) if din != dout else None

    def forward(self, x):
        out = self.norm(self.dense(x))
        out = out + (self.skip(x) if self.skip is not None else x)
        return F.silu(out)


class BasicBlock(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.dense = nn.Linear(din, dout)

    def forward(self, x):
        return F.relu(self.dense(x))


class MLP(nn.Module):
    def __init__(self, din, dout, hidden, layers, block=BasicBlock):
        super().__init__()
        net = []
        for l in range(layers):
            if l == 0:
                net.append(BasicBlock(din, hidden))
            elif l != layers - 1:
                net.append(block(hidden, hidden))
            else:
                net.append(nn.Linear(hidden, dout))
        self.net = nn.ModuleList(net)

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


def sample_pdf(bins, weights, n_samples):
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = torch.rand(list(cdf.shape[:-1]) + [n_samples]).contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    inds_g = torch.stack([below, above], -1)
    shp = [inds_g.shape[0], inds_g.shape[1], cdf.shape[-1]]
    cdf_g = torch.gather(cdf.unsqueeze(1).expand(shp), 2, inds_g)
    bins_g = torch.gather(bins.unsqueeze(1).expand(shp), 2, inds_g)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_g[..., 0]) / denom
    return bins_g[..., 0] + t * (bins_g[..., 1] - bins_g[..., 0])


class VanillaNeRF(nn.Module):
    def __init__(self, bound=1.0, min_near=0.01, blob_density=5.0, blob_radius=0.2, num_steps=64, upsample_steps=32, lambda_orient=1e-2):
        super().__init__()
        self.bound, self.min_near, self.blob_density, self.blob_radius = bound, min_near, blob_density, blob_radius
        self.num_steps, self.upsample_steps, self.lambda_orient = num_steps, upsample_steps, lambda_orient
        self.encoder = FreqEncoderTorch(3, 12)
        self.sigma_net = MLP(self.encoder.output_dim, 4, 64, 5, block=ResBlock)
        self.encoder_bg = FreqEncoderTorch(3, 4)
        self.bg_net = MLP(self.encoder_bg.output_dim, 3, 32, 2)

    def density_blob(self, x):
        d = (x ** 2).sum(-1)
        return self.blob_density * torch.exp(-d / (2 * self.blob_radius ** 2))

    def common_forward(self, x):
        h = self.sigma_net(self.encoder(x))
        sigma = torch.exp(h[..., 0] + self.density_blob(x).detach())
        return sigma, torch.sigmoid(h[..., 1:])

    def forward(self, x, l, ratio, shading):
        if shading == 'albedo':
            sigma, color = self.common_forward(x)
            return sigma, color, None
        with torch.enable_grad():
            x.requires_grad_(True)
            sigma, albedo = self.common_forward(x)
            normal = -torch.autograd.grad(torch.sum(sigma), x, create_graph=True)[0]
        normal = torch.nan_to_num(safe_normalize(normal))
        lam = ratio + (1 - ratio) * (normal * l).sum(-1).clamp(min=0)
        if shading == 'textureless':
            color = lam.unsqueeze(-1).repeat(1, 3)
        elif shading == 'normal':
            color = (normal + 1) / 2
        else:
            color = albedo * lam.unsqueeze(-1)
        return sigma, color, normal

    def render(self, rays_o, rays_d, ambient_ratio=1.0, shading='albedo', bg_color=None, perturb=True):
        rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        N = rays_o.shape[0]
        radius = rays_o.norm(dim=-1, keepdim=True)
        nears, fars = radius - self.bound, radius + self.bound
        light_d = safe_normalize(rays_o + torch.randn(3))
        T = self.num_steps
        z = torch.linspace(0.0, 1.0, T).unsqueeze(0).expand(N, T)
        z = nears + (fars - nears) * z
        sd = (fars - nears) / T
        if perturb:
            z = z + (torch.rand(z.shape) - 0.5) * sd
        lo, hi = -self.bound, self.bound
        xyzs = (rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z.unsqueeze(-1)).clamp(lo, hi)
        sigma, _ = self.common_forward(xyzs.reshape(-1, 3))
        sigma = sigma.view(N, T)
        with torch.no_grad():
            deltas = torch.cat([z[..., 1:] - z[..., :-1], sd * torch.ones_like(z[..., :1])], dim=-1)
            alphas = 1 - torch.exp(-deltas * sigma)
            shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
            weights = alphas * torch.cumprod(shifted, dim=-1)[..., :-1]
            z_mid = z[..., :-1] + 0.5 * deltas[..., :-1]
            new_z = sample_pdf(z_mid, weights[:, 1:-1], self.upsample_steps).detach()
            new_xyzs = (rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z.unsqueeze(-1)).clamp(lo, hi)
        z = torch.cat([z, new_z], dim=1)
        z, idx = torch.sort(z, dim=1)
        xyzs = torch.cat([xyzs, new_xyzs], dim=1)
        xyzs = torch.gather(xyzs, 1, idx.unsqueeze(-1).expand_as(xyzs))
        # the reference evaluates density at the new samples, then the full network again at all of them (renderer.py:625-668)
        _ = self.common_forward(new_xyzs.reshape(-1, 3))
        dirs = safe_normalize(rays_d.view(-1, 1, 3).expand_as(xyzs))
        ld = light_d.view(-1, 1, 3).expand_as(xyzs)
        sigmas, rgbs, normals = self(xyzs.reshape(-1, 3), ld.reshape(-1, 3), ambient_ratio, shading)
        sigmas = sigmas.view(N, -1)
        deltas = torch.cat([z[..., 1:] - z[..., :-1], sd * torch.ones_like(z[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * sigmas)
        shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(shifted, dim=-1)[..., :-1]
        rgbs = rgbs.view(N, -1, 3)
        weights_sum = weights.sum(-1)
        image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2)
        if bg_color is None:
            bg_color = torch.sigmoid(self.bg_net(self.encoder_bg(rays_d)))
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        out = {'image': image, 'weights': weights, 'weights_sum': weights_sum}
        if normals is not None and self.lambda_orient > 0:
            normals = normals.view(N, -1, 3)
            out['loss_orient'] = (weights.detach() * (normals * dirs).sum(-1).clamp(min=0) ** 2).sum(-1).mean()
        return out
