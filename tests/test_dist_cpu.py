"""CPU, world_size 2 over gloo: the data-parallel host logic (sdf_b200/dist.py) — flat gradient bucket + one all-reduce per
step — gives every rank the SUM of the per-rank gradients (== a single process accumulating both views), keeps .grad views
inside the bucket across steps, and the occupancy broadcast makes rank 1 adopt rank 0's bitfield."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "stable-dreamfusion_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdf_b200.dist import GradBucket, broadcast_occupancy
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    table = torch.nn.Parameter(torch.randn(101, 2))
    late = torch.nn.Parameter(torch.randn(4))          # receives its first gradient at step 1, on rank 1 only (a bg_net that a
    params = [table] + list(model.parameters()) + [late]   # schedule draw switches on later, and not on every rank)
    bucket = GradBucket(params)
    assert late.grad is not None and bucket.flat.numel() == sum((p.numel() + 3) // 4 * 4 for p in params)       # slices are 16-byte aligned
    results = []
    for step in range(3):
        g = torch.Generator().manual_seed(100 * step + rank)          # each rank sees its own "view"
        x = torch.randn(11, 7, generator=g)
        idx = torch.randint(0, 101, (11,), generator=g)
        loss = (model(x) ** 2).sum() + (table[idx] ** 2).sum() * (rank + 1)
        if step >= 1 and rank == 1:
            loss = loss + (late * late).sum() * step
        for p in params:
            if p.grad is not None:
                p.grad.zero_()
        if step == 2:
            model[0].weight.grad = None                                # something dropped a gradient: it must be re-homed, not lost
        loss.backward()
        flat = bucket.all_reduce()
        assert all(p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4 for p in params)
        results.append([p.grad.clone() for p in params])
    holder = type("M", (), {})()
    holder.density_bitfield = torch.full((64,), rank + 1, dtype=torch.uint8)
    broadcast_occupancy(holder, src=0)
    assert int(holder.density_bitfield[0]) == 1
    if rank == 0:
        torch.save(results, out)
    dist.destroy_process_group()


def test_bucket_allreduce_equals_single_process_sum(tmp_path):
    out = str(tmp_path / "r0.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-process restatement: accumulate both ranks' losses
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    table = torch.nn.Parameter(torch.randn(101, 2))
    late = torch.nn.Parameter(torch.randn(4))
    params = [table] + list(model.parameters()) + [late]
    for step in range(3):
        for p in params:
            p.grad = torch.zeros_like(p)
        for rank in range(2):
            g = torch.Generator().manual_seed(100 * step + rank)
            x = torch.randn(11, 7, generator=g)
            idx = torch.randint(0, 101, (11,), generator=g)
            loss = (model(x) ** 2).sum() + (table[idx] ** 2).sum() * (rank + 1)
            if step >= 1 and rank == 1:
                loss = loss + (late * late).sum() * step
            loss.backward()
        for a, p in zip(got[step], params):
            assert torch.allclose(a, p.grad, atol=1e-5), (step, (a - p.grad).abs().max())


def _worker_exchange(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "stable-dreamfusion_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sdf_b200.dist import exchange_pixels
    B, HW, C = 2, 12, 4
    P = HW // world
    # the "scene": pixel p of global view v has value f(v, p, c) * theta; every rank renders pixels rank::world of all views
    theta = torch.tensor(1.5, requires_grad=True)
    v = torch.arange(world * B).view(world, B, 1, 1).float()
    p_idx = (torch.arange(P) * world + rank).view(1, 1, P, 1).float()
    c = torch.arange(C).view(1, 1, 1, C).float()
    local = (100 * v + p_idx + 0.1 * c) * theta                       # [W, B, P, C]
    full = exchange_pixels(local, world)                              # [B, HW, C] complete images of my views
    my_v = (torch.arange(B) + rank * B).view(B, 1, 1).float()
    expect = (100 * my_v + torch.arange(HW).view(1, HW, 1).float() + 0.1 * torch.arange(C).view(1, 1, C).float()) * 1.5
    assert torch.allclose(full.detach(), expect), (rank, full, expect)
    # per-view loss on the owner; its gradient must come back to the ranks that rendered the pixels
    w = torch.linspace(0.5, 2.0, HW).view(1, HW, 1) * (rank + 1)
    (full * w).sum().backward()
    g = theta.grad.clone()
    dist.all_reduce(g)                                                # what the gradient bucket does for the NeRF parameters
    if rank == 0:
        torch.save(g, out)
    dist.destroy_process_group()


def test_pixel_exchange_forward_and_gradient(tmp_path):
    """ray-parallel rendering: all-to-all of rendered pixels reassembles whole views on their owner, and the summed parameter
    gradient equals the single-process gradient of the same per-view losses"""
    out = str(tmp_path / "g.pt")
    world, B, HW, C = 2, 2, 12, 4
    mp.spawn(_worker_exchange, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)
    theta = torch.tensor(1.5, requires_grad=True)
    total = 0.0
    for r in range(world):
        my_v = (torch.arange(B) + r * B).view(B, 1, 1).float()
        img = (100 * my_v + torch.arange(HW).view(1, HW, 1).float() + 0.1 * torch.arange(C).view(1, 1, C).float()) * theta
        total = total + (img * (torch.linspace(0.5, 2.0, HW).view(1, HW, 1) * (r + 1))).sum()
    total.backward()
    assert torch.allclose(got, theta.grad, rtol=1e-6), (got, theta.grad)
