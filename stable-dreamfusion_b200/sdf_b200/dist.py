"""Data-parallel plumbing for the SDS step: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch; gloo in the
CPU tests).  The path shards over independent views; the only exchange is the NeRF gradient.

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
