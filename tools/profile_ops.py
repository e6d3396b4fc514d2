"""Per-op CUDA-event timing of the three SD launch lists (UNet forward, VAE forward, VAE data-gradient): every op of a list is
replayed `reps` times back to back (warm L2 for small ops — read the totals as a lower bound of the in-step cost).
    python tools/profile_ops.py [reps] > gpurun_out/ops.txt
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import torch
from guidance.sd_utils import StableDiffusion

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
g = StableDiffusion(dev, weights="random", n_views=1, render_hw=64, seed=0, capture=False)
eng = g.engine
lists = {"unet": eng.unet.runlist, "vae_fwd": eng.vae.fwd, "vae_bwd": eng.vae.bwd}
for lname, rl in lists.items():
    for _, fn in rl.ops:
        fn()
    torch.cuda.synchronize()
    rows = []
    for name, fn in rl.ops:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        plan = getattr(fn, "__self__", None)
        rows.append((name, us, plan))
    total = sum(r[1] for r in rows)
    gemm = sum(r[1] for r in rows if r[2] is not None and hasattr(r[2], "flops"))
    fl = sum(r[2].flops for r in rows if r[2] is not None and hasattr(r[2], "flops"))
    print(f"## {lname}: {len(rows)} ops, {total / 1e3:.2f} ms replayed op by op; GEMM plans {gemm / 1e3:.2f} ms, {fl / 1e9:.0f} GFLOP -> {fl / gemm / 1e6:.0f} TFLOP/s")
    kinds = defaultdict(lambda: [0, 0.0])
    for name, us, plan in rows:
        k = "gemm" if plan is not None and hasattr(plan, "flops") else name.rsplit(".", 1)[-1].split("[")[0]
        kinds[k][0] += 1
        kinds[k][1] += us
    for k, (n, us) in sorted(kinds.items(), key=lambda kv: -kv[1][1])[:12]:
        print(f"   kind {k:28s} n={n:4d} {us:9.1f} us")
    for name, us, plan in sorted(rows, key=lambda r: -r[1])[:45]:
        extra = ""
        if plan is not None and hasattr(plan, "flops"):
            sh = plan.shape
            extra = f"M={sh['M']:7d} N={sh['N']:5d} K={sh['K']:6d} bn={sh['block_n']} sk={sh['splitk']:2d} {plan.flops / us / 1e6:7.1f} TFLOP/s"
        print(f"   {us:8.1f} us  {name:58s} {extra}")
