"""K / tile sweeps of the tcgen05 GEMM: in-stream time per launch (20 back-to-back launches, warm L2) to separate the fixed
per-launch cost (prologue, pipeline fill, epilogue) from the per-k-block slope."""
import math
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-dreamfusion_b200"))
import torch
from sdf_b200 import _lib, gemm

dev = torch.device("cuda:0")


def t_launches(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def run(Nimg, H, W, Cin, Cout, taps, bn, sk=1, pair=0, residual=False):
    a = torch.randn(Nimg, H, W, Cin, device=dev).half()
    w = torch.randn(Cout, Cin, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device=dev) / math.sqrt(taps * Cin)
    wt = gemm.pack_conv_weight(w)
    out = torch.empty(Nimg, H, W, Cout, device=dev, dtype=torch.float16)
    res = torch.randn(Nimg, H, W, Cout, device=dev).half() if residual else None
    P = gemm.conv_plan(a, Cin, wt, Cout, out, taps=taps, bias=torch.randn(Cout, device=dev), residual=res, splitk=sk, block_n=bn, cta_pair=pair)
    us = t_launches(P.run)
    fl = 2.0 * Nimg * H * W * Cout * taps * Cin
    print(f"M={Nimg*H*W:6d} N={Cout:5d} K={taps*Cin:6d} taps={taps} bn={bn:3d} pair={pair} sk={sk:2d} res={int(residual)}  {us:8.1f} us  {fl/us/1e6:8.1f} TFLOP/s", flush=True)


x = torch.zeros(8, device=dev)
print(f"baseline: torch elementwise launch {t_launches(lambda: x.add_(1.0)):.2f} us")
print("# K sweep, linear M=8192 N=320")
for K in (64, 320, 1280, 2560, 5120):
    run(1, 1, 8192, K, 320, 1, 160)
print("# K sweep, conv3x3 M=8192 N=320")
for C in (64, 320, 960):
    run(2, 64, 64, C, 320, 9, 160)
print("# N sweep (epilogue cost), linear M=8192 K=320")
for N in (160, 320, 960, 2560):
    run(1, 1, 8192, 320, N, 1, 160)
    run(1, 1, 8192, 320, N, 1, 160, residual=True)
print("# 16x16 level: M=512 N=1280")
for K, taps in ((1280, 1), (5120, 1), (1280, 9)):
    for bn, pair, sk in ((64, 0, 1), (128, 0, 1), (160, 0, 1), (160, 0, 2), (160, 0, 4), (256, 1, 1), (256, 1, 4), (128, 1, 2)):
        run(2, 16, 16, K, 1280, taps, bn, sk, pair)
print("# 32x32 level: M=2048 N=640")
for K, taps in ((640, 1), (640, 9), (1920, 9)):
    for bn, pair, sk in ((64, 0, 1), (128, 0, 1), (160, 0, 1), (160, 0, 2), (160, 1, 1), (160, 1, 2), (128, 1, 1)):
        run(2, 32, 32, K, 640, taps, bn, sk, pair)
print("# 8x8 level: M=128 N=1280")
for K, taps in ((1280, 1), (1280, 9), (2560, 9)):
    for bn, sk in ((64, 1), (64, 4), (160, 1), (160, 4), (160, 15)):
        run(2, 8, 8, K, 1280, taps, bn, sk, 0)
