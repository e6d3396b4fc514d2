"""Data-parallel plumbing for the SDS step: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch; gloo in the
CPU tests).  The path shards over independent views.  Exchanges per step:
  * the NeRF gradient — ONE all-reduce over a flat bucket (GradBucket);
  * (ray-parallel rendering, trainer.py) two 16-KB-per-peer all-to-alls: rendered pixels to the rank that owns the view, and
    the SDS pixel gradients back (exchange_pixels).  Every rank renders 1/W of the rays of EVERY view, so the sample count —
    the only part of a step whose cost varies with the camera — is balanced across ranks by construction and no rank waits
    for a straggler at the gradient all-reduce; the UNet / VAE work stays one whole view per rank.

GradBucket re-homes every parameter's .grad inside ONE flat fp32 buffer (the 48.8 MB hash-table gradient first), so a step
needs a single all-reduce and no gather/scatter copies; the 1/world factor is applied inside the fused Adan kernel."""
import torch
import torch.distributed as dist


class GradBucket:
    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None

    def _build(self):
        ps = [p for p in self.params if p.grad is not None]
        dev = ps[0].device
        self.flat = torch.zeros(sum(p.numel() for p in ps), device=dev, dtype=torch.float32)
        off = 0
        for p in ps:
            n = p.numel()
            view = self.flat[off:off + n].view_as(p)
            view.copy_(p.grad)
            p.grad = view
            off += n
        self.members = ps

    def all_reduce(self):
        """SUM over ranks, in place in every parameter's .grad (which lives inside the bucket after the first call)."""
        if self.flat is None:
            self._build()
        else:
            for p in self.members:      # a gradient that autograd re-allocated must be brought back into the bucket
                if p.grad is not None and (p.grad.data_ptr() < self.flat.data_ptr() or p.grad.data_ptr() >= self.flat.data_ptr() + self.flat.numel() * 4):
                    self.flat = None
                    self._build()
                    break
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self.flat


def broadcast_occupancy(model, src=0):
    """rank `src`'s density bitfield (262 144 B at bound 1) to everyone, so all ranks march the same cells"""
    dist.broadcast(model.density_bitfield, src=src)


class _AllToAll(torch.autograd.Function):
    """y[s] on rank d = x[d] on rank s (equal splits along dim 0).  Its adjoint is the same exchange."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        out = torch.empty_like(x)
        dist.all_to_all_single(out, x)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty_like(g)
        dist.all_to_all_single(out, g)
        return out


def exchange_pixels(local, world):
    """local: [W, B, P, C] — this rank's P = HW / W rendered pixels (pixel p = j * W + rank) of the B views owned by each of the W
    ranks.  Returns [B, HW, C]: the complete images of this rank's own views.  Differentiable (gradients travel back)."""
    W_, B, P, C = local.shape
    assert W_ == world
    got = _AllToAll.apply(local)                      # [src rank s, B, j, C] = pixel j * W + s of my views
    return got.permute(1, 2, 0, 3).reshape(B, P * world, C)
