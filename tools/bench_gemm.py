"""Times the tcgen05 implicit-GEMM kernel on the UNet / VAE shapes (CUDA events, warm-up, L2 flushed between reps)."""
import math
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-dreamfusion_b200"))
import torch
from sdf_b200 import _lib, gemm

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, Nimg, H, W, Cin, Cout, taps, bn, splitk=1, reps=10, pair=0):
    a = torch.randn(Nimg, H, W, Cin, device=dev).half()
    w = torch.randn(Cout, Cin, 3 if taps == 9 else 1, 3 if taps == 9 else 1, device=dev) / math.sqrt(taps * Cin)
    wt = gemm.pack_conv_weight(w)
    out = torch.empty(Nimg, H, W, Cout, device=dev, dtype=torch.float16)
    bias = torch.randn(Cout, device=dev)
    P = gemm.conv_plan(a, Cin, wt, Cout, out, taps=taps, bias=bias, splitk=splitk, block_n=bn, cta_pair=pair)
    plan = P.handle
    st = _lib.stream()
    for _ in range(3):
        _lib.call("sdf_gemm_run", plan, st)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("sdf_gemm_run", plan, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * Nimg * H * W * Cout * taps * Cin
    print(f"{name:34s} M={Nimg*H*W:7d} N={Cout:5d} K={taps*Cin:6d} bn={bn:3d} pair={pair} sk={splitk:2d}  {ms*1e3:9.1f} us  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
    del P


if __name__ == "__main__":
    cases = [("unet conv3x3 320->320 @64", 2, 64, 64, 320, 320, 9, 160, 160),
             ("unet conv3x3 960->320 @64", 2, 64, 64, 960, 320, 9, 160, 160),
             ("unet conv3x3 640->640 @32", 2, 32, 32, 640, 640, 9, 160, 160),
             ("unet conv3x3 1280->1280 @16", 2, 16, 16, 1280, 1280, 9, 160, 256),
             ("unet linear 320->2560 @64", 1, 1, 8192, 320, 2560, 1, 160, 256),
             ("unet linear 1280->320 @64", 1, 1, 8192, 1280, 320, 1, 160, 160),
             ("unet qkv 320->960 @64", 1, 1, 8192, 320, 960, 1, 160, 160),
             ("vae conv3x3 128->128 @512", 1, 512, 512, 128, 128, 9, 128, 128),
             ("vae conv3x3 256->256 @256", 1, 256, 256, 256, 256, 9, 128, 256),
             ("vae conv3x3 512->512 @128", 1, 128, 128, 512, 512, 9, 128, 256),
             ("vae conv3x3 512->512 @64", 1, 64, 64, 512, 512, 9, 128, 256)]
    for name, Nimg, H, W, Cin, Cout, taps, bn1, bn2 in cases:
        bench(name, Nimg, H, W, Cin, Cout, taps, bn1)
        bench(name, Nimg, H, W, Cin, Cout, taps, bn2, pair=1)
    bench("unet conv3x3 640->640 @32 sk2", 2, 32, 32, 640, 640, 9, 160, 2, pair=1)
    bench("unet conv3x3 1280->1280 @16 sk4", 2, 16, 16, 1280, 1280, 9, 160, 4)
    bench("unet conv3x3 1280->1280 @16 sk4", 2, 16, 16, 1280, 1280, 9, 256, 4, pair=1)
    bench("unet conv3x3 1280->1280 @16 sk8", 2, 16, 16, 1280, 1280, 9, 256, 8, pair=1)
    bench("unet conv3x3 1280->1280 @8 sk16", 2, 8, 8, 1280, 1280, 9, 160, 15)
