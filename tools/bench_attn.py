"""Self-attention at the UNet's shapes: tcgen05 kernel (flash_attn_tc.cu) vs the mma.sync kernel (flash_attn.cu), CUDA events, median of 20."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import torch
from sdf_b200 import _lib

if len(sys.argv) == 1:
    for tc in ("1", "0"):
        subprocess.run([sys.executable, os.path.abspath(__file__), tc], env=dict(os.environ, SDF_FLASH_TC=tc))
    sys.exit(0)
dev = torch.device("cuda:0")
for (B, heads, n, d) in [(2, 8, 4096, 40), (2, 8, 1024, 80), (2, 8, 1024, 40), (4, 8, 4096, 40)]:
    C = heads * d
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B, n, 3 * C, device=dev, generator=g).half()
    o = torch.empty(B, n, C, device=dev, dtype=torch.float16)
    ts = []
    for rep in range(25):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("sdf_flash_attention", qkv.data_ptr(), qkv.data_ptr() + 2 * C, qkv.data_ptr() + 4 * C, o.data_ptr(), B, heads, n, n, d, 3 * C, 3 * C, C, d ** -0.5, _lib.stream())
        e1.record()
        torch.cuda.synchronize()
        if rep >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    fl = 4.0 * B * heads * n * n * d
    print(f"SDF_FLASH_TC={os.environ.get('SDF_FLASH_TC')}  B{B} h{heads} n{n} d{d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s", flush=True)
