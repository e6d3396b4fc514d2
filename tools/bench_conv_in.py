"""VAE conv_in (3 -> 128 at 512x512) and its data-gradient: the direct FMA-pipe kernels against the zero-padded implicit GEMM (CUDA events, L2 flushed
before every launch).   python tools/bench_conv_in.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-dreamfusion_b200"))
import torch
from sdf_b200 import _lib, gemm

dev = torch.device("cuda:0")
H = W = 512
C = 128
x = torch.zeros(1, H, W, 8, device=dev, dtype=torch.float16)
x[..., :3] = torch.randn(1, H, W, 3, device=dev).half()
w = (torch.randn(C, 3, 3, 3, device=dev) * 0.2)
b = torch.randn(C, device=dev) * 0.1
y = torch.empty(1, H, W, C, device=dev, dtype=torch.float16)
dy = (torch.randn(1, H, W, C, device=dev) * 0.1).half()
dx = torch.zeros(1, H, W, 8, device=dev, dtype=torch.float16)
flush = torch.empty(64 << 20, device=dev, dtype=torch.float32)
P = _lib.ptr
st = _lib.stream()
plan_f = gemm.conv_plan(x, 3, gemm.pack_conv_weight(w.half()), C, y, taps=9, bias=b)
wflip = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()
plan_b = gemm.conv_plan(dy, C, gemm.pack_conv_weight(wflip.half()), 3, dx, taps=9)


def timeit(fn, n=8):
    ts = []
    for i in range(n + 2):
        flush.fill_(float(i))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    return sum(ts) / len(ts)


print(f"forward  direct {timeit(lambda: _lib.call('sdf_conv3x3_small_cin_forward', P(x), 8, P(w), P(b), P(y), C, 1, H, W, 3, C, st)):7.1f} us   implicit GEMM {timeit(plan_f.run):7.1f} us")
print(f"dgrad    direct {timeit(lambda: _lib.call('sdf_conv3x3_small_cin_dgrad', P(dy), C, P(w), P(dx), 8, 1, H, W, 3, C, st)):7.1f} us   implicit GEMM {timeit(plan_b.run):7.1f} us")
