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
    """Every requires_grad parameter's .grad re-homed inside ONE flat fp32 buffer, built up front over ALL parameters (zeros for
    those that have no gradient yet): membership never depends on which parameters a particular step touched, so every rank's
    bucket has the same layout and size whatever its schedule drew (a background net that first receives a gradient at step 40 is
    already in it).  A gradient that something re-allocated outside the bucket is copied back in before the exchange."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        pad4 = lambda n: (n + 3) // 4 * 4            # every slice starts 16-byte aligned: the fused Adan kernels read float4
        self.flat = torch.zeros(sum(pad4(p.numel()) for p in self.params), device=dev, dtype=torch.float32)
        self.views = []
        off = 0
        for p in self.params:
            n = p.numel()
            v = self.flat[off:off + n].view_as(p)
            if p.grad is not None:
                v.copy_(p.grad)
            p.grad = v
            self.views.append(v)
            off += pad4(n)

    def rehome(self):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
                p.grad = v
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v

    def all_reduce(self):
        """SUM over ranks, in place in every parameter's .grad"""
        self.rehome()
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
