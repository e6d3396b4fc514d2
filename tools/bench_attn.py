"""Self-attention at the UNet's shapes: tcgen05 kernels (flash_attn_tc.cu, versions 1 and 2) vs the mma.sync kernel (flash_attn.cu), CUDA events,
median of 20; every variant is also compared with an fp32 PyTorch softmax(QK^T)V on a slice of the rows."""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stable-dreamfusion_b200")]
import torch
from sdf_b200 import _lib

if len(sys.argv) == 1:
    for tc, ver in (("1", "2"), ("1", "4")):
        subprocess.run([sys.executable, os.path.abspath(__file__), tc], env=dict(os.environ, SDF_FLASH_TC=tc, SDF_FLASH_TC_V=ver), timeout=300)
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
    # accuracy on the first 256 query rows of batch 0, every head
    q, k, v = (qkv[0, :, i * C:(i + 1) * C].float().view(n, heads, d).permute(1, 0, 2) for i in range(3))
    ref = torch.softmax(q[:, :256] @ k.transpose(1, 2) * d ** -0.5, -1) @ v
    err = (o[0, :256].float().view(256, heads, d).permute(1, 0, 2) - ref).abs().max().item()
    print(f"SDF_FLASH_TC={os.environ.get('SDF_FLASH_TC')} V={os.environ.get('SDF_FLASH_TC_V')}  B{B} h{heads} n{n} d{d}: {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s  "
          f"max|err| {err:.2e} (ref max {ref.abs().max().item():.2f})", flush=True)
