"""Host-side mirror of the reference's -O backbone, nerf/network_grid.py:13-171: same module / parameter
names (encoder.embeddings, encoder.offsets, sigma_net.net.{0,1,2}.{weight,bias}, bg_net.net.{0,1}.*), same
methods (common_forward, finite_difference_normal, normal, forward, density, background, get_params).

`fused=True` (default) routes forward()/density() through the single fused kernel pair of
csrc/fused_field{,_bwd}.cu; `fused=False` runs the operator-by-operator graph exactly as the reference does
(on the drop-in gridencoder / freqencoder ops) and is what the fused path is tested against.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.amp import custom_bwd, custom_fwd

from freqencoder import FreqEncoder
from gridencoder import GridEncoder

from .field import fused_field
from .renderer import NeRFRenderer, safe_normalize


class _trunc_exp(torch.autograd.Function):
    # activation.py:5-18
    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float)
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g):
        x = ctx.saved_tensors[0]
        return g * torch.exp(x.clamp(max=15))


trunc_exp = _trunc_exp.apply


class MLP(nn.Module):
    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.net = nn.ModuleList([
            nn.Linear(dim_in if l == 0 else dim_hidden, dim_out if l == num_layers - 1 else dim_hidden, bias=bias)
            for l in range(num_layers)])

    def forward(self, x):
        for l in range(self.num_layers):
            x = self.net[l](x)
            if l != self.num_layers - 1:
                x = F.relu(x, inplace=True)
        return x


class NeRFNetwork(NeRFRenderer):
    def __init__(self, opt, num_layers=3, hidden_dim=64, num_layers_bg=2, hidden_dim_bg=32, fused=True):
        super().__init__(opt)
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                                   desired_resolution=2048 * self.bound, gridtype='hash', align_corners=False, interpolation='smoothstep')
        self.in_dim = self.encoder.output_dim
        self.sigma_net = MLP(self.in_dim, 4, hidden_dim, num_layers, bias=True)
        if opt.density_activation != 'exp':
            raise NotImplementedError('this build implements the exp density activation of the -O preset')
        self.density_activation = trunc_exp
        if self.opt.bg_radius > 0:
            self.num_layers_bg = num_layers_bg
            self.hidden_dim_bg = hidden_dim_bg
            self.encoder_bg = FreqEncoder(input_dim=3, degree=6)
            self.in_dim_bg = self.encoder_bg.output_dim
            self.bg_net = MLP(self.in_dim_bg, 3, hidden_dim_bg, num_layers_bg, bias=True)
        else:
            self.bg_net = None
        self.fused = fused and num_layers == 3 and hidden_dim == 64

    # ---- operator-by-operator graph (reference structure)
    def common_forward(self, x):
        enc = self.encoder(x, bound=self.bound, max_level=self.max_level)
        h = self.sigma_net(enc)
        sigma = self.density_activation(h[..., 0] + self.density_blob(x))
        albedo = torch.sigmoid(h[..., 1:])
        return sigma, albedo

    def finite_difference_normal(self, x, epsilon=1e-2):
        def at(dx, dy, dz):
            return self.common_forward((x + torch.tensor([[dx, dy, dz]], device=x.device)).clamp(-self.bound, self.bound))[0]
        dx_pos, dx_neg = at(epsilon, 0, 0), at(-epsilon, 0, 0)
        dy_pos, dy_neg = at(0, epsilon, 0), at(0, -epsilon, 0)
        dz_pos, dz_neg = at(0, 0, epsilon), at(0, 0, -epsilon)
        normal = torch.stack([0.5 * (dx_pos - dx_neg) / epsilon, 0.5 * (dy_pos - dy_neg) / epsilon, 0.5 * (dz_pos - dz_neg) / epsilon], dim=-1)
        return -normal

    def normal(self, x):
        if self.fused:
            _, _, n = self._fused(x, torch.zeros(3, device=x.device), 1.0, 'normal')
            return n
        normal = self.finite_difference_normal(x)
        normal = safe_normalize(normal)
        return torch.nan_to_num(normal)

    def _levels_active(self):
        import math
        L = self.encoder.num_levels
        return L if self.max_level is None else max(min(int(math.ceil(self.max_level * L)), L), 1)

    def _fused(self, x, l, ratio, shading, want_color=True):
        n = self.sigma_net.net
        return fused_field(x, self.encoder.embeddings, n[0].weight, n[0].bias, n[1].weight, n[1].bias, n[2].weight, n[2].bias,
                           self.encoder.offsets, l, shading=shading, ratio=ratio, bound=self.bound,
                           per_level_scale=self.encoder.per_level_scale, base_resolution=self.encoder.base_resolution,
                           smoothstep=self.encoder.interp_id == 1, levels_active=self._levels_active(),
                           blob_density=self.opt.blob_density, blob_radius=self.opt.blob_radius, want_color=want_color)

    def forward(self, x, d, l=None, ratio=1, shading='albedo'):
        if self.fused:
            sigma, color, normal = self._fused(x, l, ratio, shading)
            return sigma, color, normal
        sigma, albedo = self.common_forward(x)
        if shading == 'albedo':
            normal = None
            color = albedo
        else:
            normal = self.normal(x)
            lambertian = ratio + (1 - ratio) * (normal * l).sum(-1).clamp(min=0)
            if shading == 'textureless':
                color = lambertian.unsqueeze(-1).repeat(1, 3)
            elif shading == 'normal':
                color = (normal + 1) / 2
            else:
                color = albedo * lambertian.unsqueeze(-1)
        return sigma, color, normal

    def density(self, x):
        if self.fused:
            sigma, albedo, _ = self._fused(x, None, 1.0, 'albedo')
            return {'sigma': sigma, 'albedo': albedo}
        sigma, albedo = self.common_forward(x)
        return {'sigma': sigma, 'albedo': albedo}

    def background(self, d):
        h = self.encoder_bg(d)
        h = self.bg_net(h)
        return torch.sigmoid(h)

    def get_params(self, lr):
        params = [
            {'params': self.encoder.parameters(), 'lr': lr * 10},
            {'params': self.sigma_net.parameters(), 'lr': lr},
        ]
        if self.opt.bg_radius > 0:
            params.append({'params': self.bg_net.parameters(), 'lr': lr})
        return params
