"""Per-node cost of a CUDA graph replay: N dependent launches of (a) a 64-element copy, (b) a one-tile GEMM, (c) a mid GEMM.
Tells how much of the ~950-launch SD lists is launch floor rather than work."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-dreamfusion_b200"))
import torch
from sdf_b200 import _lib, gemm

dev = torch.device("cuda:0")
N = 200


def graph_time(fn, n=N):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 5 / n


x = torch.zeros(8, 8, device=dev, dtype=torch.float16)
y = torch.zeros(8, 8, device=dev, dtype=torch.float16)
print(f"copy 64 elements          : {graph_time(lambda: _lib.call('sdf_copy2d', _lib.ptr(x), 8, _lib.ptr(y), 8, 8, 8, _lib.stream())):6.2f} us / node")
big = torch.zeros(8192, 320, device=dev, dtype=torch.float16)
big2 = torch.zeros(8192, 320, device=dev, dtype=torch.float16)
print(f"copy 8192x320 (5 MB)      : {graph_time(lambda: _lib.call('sdf_copy2d', _lib.ptr(big), 320, _lib.ptr(big2), 320, 8192, 320, _lib.stream())):6.2f} us / node")
for (M, K, Nn, bn) in ((128, 64, 64, 64), (8192, 64, 320, 160), (8192, 320, 320, 160), (8192, 1280, 320, 160), (2048, 640, 640, 128), (512, 1280, 1280, 64)):
    a = torch.randn(1, 1, M, K, device=dev).half()
    w = gemm.pack_conv_weight(torch.randn(Nn, K, 1, 1, device=dev))
    out = torch.empty(1, 1, M, Nn, device=dev, dtype=torch.float16)
    P = gemm.conv_plan(a, K, w, Nn, out, taps=1, bias=torch.zeros(Nn, device=dev), block_n=bn)
    print(f"gemm M={M:5d} K={K:5d} N={Nn:4d}  : {graph_time(P.run):6.2f} us / node")
xs = torch.randn(2, 4096, 320, device=dev).half()
ys = torch.empty_like(xs)
gm, bt = torch.ones(320, device=dev), torch.zeros(320, device=dev)
st = torch.empty(2, 32, 2, device=dev)
print(f"groupnorm 2x4096x320 (memset + 2 kernels): {graph_time(lambda: _lib.call('sdf_groupnorm_forward', _lib.ptr(xs), 320, _lib.ptr(ys), 320, 2, 4096, 320, 32, _lib.ptr(gm), _lib.ptr(bt), 1e-5, 1, _lib.ptr(st), _lib.stream())):6.2f} us / call")
print(f"layernorm 8192x320        : {graph_time(lambda: _lib.call('sdf_layernorm_forward', _lib.ptr(xs), 320, _lib.ptr(ys), 320, 8192, 320, _lib.ptr(gm), _lib.ptr(bt), 1e-5, _lib.stream())):6.2f} us / node")

for (Ni, HW, C) in ((2, 4096, 640), (2, 1024, 1280), (2, 256, 1280), (1, 16384, 256), (1, 4096, 512)):
    xs = torch.randn(Ni, HW, C, device=dev).half(); ys = torch.empty_like(xs)
    gm, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    st = torch.empty(Ni, 32, 2, device=dev)
    print(f"groupnorm {Ni}x{HW}x{C}: {graph_time(lambda: _lib.call('sdf_groupnorm_forward', _lib.ptr(xs), C, _lib.ptr(ys), C, Ni, HW, C, 32, _lib.ptr(gm), _lib.ptr(bt), 1e-5, 1, _lib.ptr(st), _lib.stream()), 50):6.2f} us / call")
