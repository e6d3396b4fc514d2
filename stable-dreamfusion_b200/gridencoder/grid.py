"""Drop-in `gridencoder` package (reference: gridencoder/grid.py:25-205).

Same surface — grid_encode(inputs, embeddings, offsets, per_level_scale, base_resolution,
calc_grad_inputs=False, gridtype=0, align_corners=False, interpolation=0, max_level=None)
and GridEncoder(...) with .embeddings [n,C] / .offsets [L+1] / grad_total_variation /
grad_weight_decay — backed by libsdf_b200.so (csrc/gridenc.cu).

What changes underneath:
  * outputs are produced directly as [B, L*C] and the gradient is consumed in that layout
    (no permute / .contiguous() kernels, grid.py:64,82);
  * with an fp16 table the table gradient is accumulated in fp32 (atomicAdd on float2)
    and returned as fp32: no fp16 atomics, no cast-back of 12.2 M entries, better precision.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from sdf_b200 import _lib

_gridtype_to_id = {'hash': 0, 'tiled': 1}
_interp_to_id = {'linear': 0, 'smoothstep': 1}

def _half_table(embeddings):
    # cast per call, like the reference (grid.py:46-47); a (data_ptr, _version) cache is unsafe under `.data` writes
    return embeddings.detach().to(torch.half)


class _grid_encode(Function):
    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, max_level=None):
        _lib.require_cuda(inputs, embeddings, offsets)
        inputs = inputs.float().contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        max_level = L if max_level is None else max(min(int(math.ceil(max_level * L)), L), 1)

        # half table only under autocast and for even C, as the reference (grid.py:44-47)
        use_half = torch.is_autocast_enabled('cuda') and C % 2 == 0 and embeddings.dtype == torch.float32
        if use_half:
            table = _half_table(embeddings)
        else:
            table = embeddings.detach().contiguous()
        if table.dtype not in (torch.float32, torch.float16):
            raise RuntimeError('grid_encode: embeddings must be float32 or float16')
        dtype_id = 1 if table.dtype == torch.float16 else 0

        alloc = torch.zeros if max_level < L else torch.empty
        outputs = alloc(B, L * C, device=inputs.device, dtype=table.dtype)
        dy_dx = alloc(B, L * D * C, device=inputs.device, dtype=table.dtype) if calc_grad_inputs else None

        _lib.call('sdf_grid_encode_forward', _lib.ptr(inputs), _lib.ptr(table), _lib.ptr(offsets), _lib.ptr(outputs), B, D, C, L,
                  max_level, S, H, _lib.ptr(dy_dx), int(gridtype), int(bool(align_corners)), int(interpolation), dtype_id,
                  _lib.stream())

        ctx.save_for_backward(inputs, offsets, dy_dx)
        ctx.dims = [B, D, C, L, S, H, gridtype, interpolation, max_level, dtype_id]
        ctx.align_corners = align_corners
        ctx.emb_shape = tuple(embeddings.shape)
        ctx.emb_dtype = embeddings.dtype
        return outputs

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, grad):
        inputs, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, interpolation, max_level, dtype_id = ctx.dims
        grad = grad.contiguous()
        want = torch.float16 if dtype_id == 1 else torch.float32
        if grad.dtype != want:
            grad = grad.to(want)
        # fp32 accumulation regardless of the table dtype
        grad_embeddings = torch.zeros(ctx.emb_shape, dtype=torch.float32, device=grad.device)
        grad_inputs = torch.zeros(B, D, dtype=want, device=grad.device) if dy_dx is not None else None
        _lib.call('sdf_grid_encode_backward', _lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(offsets), _lib.ptr(grad_embeddings), B, D, C, L,
                  max_level, S, H, _lib.ptr(dy_dx), _lib.ptr(grad_inputs), int(gridtype), int(bool(ctx.align_corners)),
                  int(interpolation), dtype_id, 0, _lib.stream())
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        if ctx.emb_dtype != torch.float32:
            grad_embeddings = grad_embeddings.to(ctx.emb_dtype)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash', align_corners=False, interpolation='linear'):
        super().__init__()
        # the finest resolution desired at the last level, if provided, overrides per_level_scale
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))

        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners

        # level offsets: host float64 arithmetic exactly as the reference (grid.py:124-134)
        offsets = []
        offset = 0
        self.max_params = 2 ** log2_hashmap_size
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            params_in_level = min(self.max_params, (resolution) ** input_dim)
            params_in_level = int(np.ceil(params_in_level / 8) * 8)
            offsets.append(offset)
            offset += params_in_level
        offsets.append(offset)
        offsets = torch.from_numpy(np.array(offsets, dtype=np.int32))
        self.register_buffer('offsets', offsets)
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        std = 1e-4
        self.embeddings.data.uniform_(-std, std)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1, max_level=None):
        # inputs: [..., input_dim] in [-bound, bound]; max_level: fraction of levels to evaluate (None = all)
        inputs = (inputs + bound) / (2 * bound)
        prefix_shape = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        outputs = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                              inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id, max_level)
        return outputs.view(prefix_shape + [self.output_dim])

    @torch.amp.autocast('cuda', enabled=False)
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        D = self.input_dim
        C = self.embeddings.shape[1]
        L = self.offsets.shape[0] - 1
        S = float(np.log2(self.per_level_scale))
        H = self.base_resolution
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = (inputs + bound) / (2 * bound)
            inputs = inputs.view(-1, self.input_dim)
            B = inputs.shape[0]
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        inputs = inputs.float().contiguous()
        _lib.call('sdf_grid_grad_total_variation', _lib.ptr(inputs), _lib.ptr(self.embeddings), _lib.ptr(self.embeddings.grad),
                  _lib.ptr(self.offsets), float(weight), B, D, C, L, S, int(H), self.gridtype_id, int(bool(self.align_corners)),
                  _lib.stream())

    @torch.amp.autocast('cuda', enabled=False)
    def grad_weight_decay(self, weight=0.1):
        B = self.embeddings.shape[0]
        C = self.embeddings.shape[1]
        L = self.offsets.shape[0] - 1
        if self.embeddings.grad is None:
            raise ValueError('grad is None, should be called after loss.backward() and before optimizer.step()!')
        _lib.call('sdf_grid_grad_weight_decay', _lib.ptr(self.embeddings), _lib.ptr(self.embeddings.grad), _lib.ptr(self.offsets),
                  float(weight), B, C, L, _lib.stream())
