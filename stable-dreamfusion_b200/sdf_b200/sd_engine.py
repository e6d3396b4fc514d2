"""SD-1.5-shaped UNet (forward) and VAE encoder (forward + data-gradient) as a static list of native launches.

Every dense contraction is a tcgen05 plan (csrc/sd_gemm.cu) — 3x3 / 1x1 convolutions as implicit GEMM over NHWC fp16
activations, linear layers, and the attention products as batched plans straight on the [tokens, heads*d] projections —
glued by the memory-bound kernels of csrc/sd_ops.cu.  Shapes are fixed at construction, so the whole step is a flat
list of ~1000 launches with no Python tensor ops in between; `capture()` records it into a CUDA graph.

Structure follows the modules the reference executes (vendored CompVis code; see oracle/sd_ref.py for the citations):
ResBlock / SpatialTransformer / BasicTransformerBlock / CrossAttention / GEGLU / Downsample / Upsample / UNetModel,
and the VAE Encoder (ResnetBlock / AttnBlock / Downsample) + quant_conv.  State-dict keys are the CompVis ones.

Layout conventions
  activations   NHWC fp16, a `View` = (tensor [Nimg,H,W,ld], channel offset, channels); skip connections are written
                by their producer directly into the consumer's concat buffer (no torch.cat kernels)
  conv weights  [Cout, tap, Cin_iter] fp16, K index = tap*Cin_iter + c (Cin_iter = Cin rounded up to 64)
  attention     S = Q K^T and O = P V are batched plans over (batch, head); the K tail (d_head = 40/80/160) is
                zero-filled by the TMA unit, V is produced transposed ([C, tokens]) by swapping the operands of its
                projection GEMM
"""
import math

import torch

from . import _lib
from .gemm import GemmPlan, conv_plan, linear_plan, pack_conv_weight, pick_block_n

NUM_SMS = 148


def _r(x, m):
    return ((x + m - 1) // m) * m


class View:
    """channels [off, off+C) of an NHWC fp16 tensor [Nimg, H, W, ld]"""

    def __init__(self, t, off=0, C=None):
        assert t.dtype == torch.float16 and t.dim() == 4 and t.is_contiguous()
        self.t, self.off = t, off
        self.C = t.shape[-1] - off if C is None else C
        assert off % 8 == 0

    @property
    def ptr(self):
        return self.t.data_ptr() + 2 * self.off

    @property
    def ld(self):
        return self.t.shape[-1]

    @property
    def Nimg(self):
        return self.t.shape[0]

    @property
    def H(self):
        return self.t.shape[1]

    @property
    def W(self):
        return self.t.shape[2]

    @property
    def rows(self):
        return self.t.shape[0] * self.t.shape[1] * self.t.shape[2]

    def strides(self):
        ld = self.ld
        return (ld, self.W * ld, self.H * self.W * ld)

    def sub(self, off, C):
        return View(self.t, self.off + off, C)

    def torch(self):
        return self.t[..., self.off:self.off + self.C]


class _PtrTensor:
    """duck-typed stand-in so GemmPlan can take a View (pointer with channel offset) where it expects a tensor"""

    def __init__(self, view):
        self.view = view
        self.device = view.t.device

    def data_ptr(self):
        return self.view.ptr


def _choose_splitk(M, N, kblocks, block_n):
    tiles = ((M + 127) // 128) * ((N + block_n - 1) // block_n)
    if tiles >= NUM_SMS // 2 or kblocks < 8:
        return 1
    want = max(1, min(NUM_SMS // tiles, kblocks // 4))
    return want


class Builder:
    """Accumulates launches.  Each op is a zero-argument callable; tensors it touches are kept alive by closures."""

    def __init__(self, device):
        self.device = device
        self.ops = []
        self.flops = 0.0
        self.bytes_act = 0

    def buf(self, Nimg, H, W, C, zero=False):
        f = torch.zeros if zero else torch.empty
        t = f(Nimg, H, W, C, device=self.device, dtype=torch.float16)
        self.bytes_act += t.numel() * 2
        return t

    def f32(self, *shape, zero=False):
        return (torch.zeros if zero else torch.empty)(*shape, device=self.device, dtype=torch.float32)

    def add(self, name, fn):
        self.ops.append((name, fn))

    # ---- dense
    def gemm(self, name, a, c_valid, wt, N, out, *, taps=1, bias=None, temb=None, temb_ld=0, residual=None, act=None, alpha=1.0,
             splitk=None, block_n=None, w_strides=None, w_k_valid=None, n_rows_w=None, geom=None, a_strides=None, o_strides=None,
             r_strides=None, cin_iter=None):
        """a / out / residual: View (or (ptr-holder, strides) through a_strides/o_strides with geom=(Nimg,H,W))."""
        Nimg, H, W = geom if geom is not None else (a.Nimg, a.H, a.W)
        a_str = a_strides if a_strides is not None else a.strides()
        o_str = o_strides if o_strides is not None else out.strides()
        if residual is not None:
            r_str = r_strides if r_strides is not None else residual.strides()
        else:
            r_str = (0, 0, 0)
        if cin_iter is None:
            cin_iter = wt.shape[-1] // taps
        w_str = w_strides if w_strides is not None else (wt.shape[-1], 0, 0)
        bn = pick_block_n(N) if block_n is None else block_n
        M = Nimg * H * W
        kb = taps * cin_iter // 64
        sk = _choose_splitk(M, N, kb, bn) if splitk is None else splitk
        wrap = lambda v: _PtrTensor(v) if isinstance(v, View) else v
        plan = GemmPlan(wrap(a), a_str, c_valid, wrap(wt), w_str, (taps * cin_iter if w_k_valid is None else w_k_valid),
                        (wt.shape[0] if n_rows_w is None else n_rows_w), Nimg, H, W, cin_iter, taps, N, wrap(out), o_str, bias=bias,
                        temb=wrap(temb) if temb is not None else None, temb_ld=temb_ld, residual=wrap(residual) if residual is not None else None,
                        r_strides=r_str, act=act, alpha=alpha, splitk=sk, block_n=bn)
        self.flops += 2.0 * M * N * taps * c_valid
        self.add(name, plan.run)
        return plan

    # ---- memory-bound
    def groupnorm(self, name, x, y, gamma, beta, eps, silu, stats=None):
        stats = self.f32(x.Nimg, 32, 2) if stats is None else stats
        args = (x.ptr, x.ld, y.ptr, y.ld, x.Nimg, x.H * x.W, x.C, 32, _lib.ptr(gamma), _lib.ptr(beta), float(eps), int(silu), _lib.ptr(stats))
        keep = (x, y, gamma, beta, stats)
        self.add(name, lambda a=args, k=keep: _lib.call('sdf_groupnorm_forward', *a, _lib.stream()))
        return stats

    def groupnorm_bwd(self, name, x, dy, dx, gamma, beta, eps, silu, stats, accumulate):
        bstats = self.f32(x.Nimg, 32, 2)
        args = (x.ptr, x.ld, dy.ptr, dy.ld, dx.ptr, dx.ld, x.Nimg, x.H * x.W, x.C, 32, _lib.ptr(gamma), _lib.ptr(beta), float(eps), int(silu),
                _lib.ptr(stats), _lib.ptr(bstats), int(accumulate))
        keep = (x, dy, dx, gamma, beta, stats, bstats)
        self.add(name, lambda a=args, k=keep: _lib.call('sdf_groupnorm_backward', *a, _lib.stream()))

    def layernorm(self, name, x, y, gamma, beta, eps=1e-5):
        args = (x.ptr, x.ld, y.ptr, y.ld, x.rows, x.C, _lib.ptr(gamma), _lib.ptr(beta), float(eps))
        keep = (x, y, gamma, beta)
        self.add(name, lambda a=args, k=keep: _lib.call('sdf_layernorm_forward', *a, _lib.stream()))

    def softmax(self, name, s, rows, cols, ld, scale=1.0):
        args = (s.data_ptr(), s.data_ptr(), rows, cols, ld, float(scale))
        self.add(name, lambda a=args, k=s: _lib.call('sdf_softmax_rows', *a, _lib.stream()))

    def softmax_bwd(self, name, p, dp, ds, rows, cols, ld, scale):
        args = (p.data_ptr(), dp.data_ptr(), ds.data_ptr(), rows, cols, ld, float(scale))
        self.add(name, lambda a=args, k=(p, dp, ds): _lib.call('sdf_softmax_rows_backward', *a, _lib.stream()))

    def geglu(self, name, x, y, inner):
        args = (x.ptr, x.ld, y.ptr, y.ld, x.rows, inner)
        self.add(name, lambda a=args, k=(x, y): _lib.call('sdf_geglu', *a, _lib.stream()))

    def upsample2(self, name, x, y):
        args = (x.ptr, x.ld, y.ptr, y.ld, x.Nimg, x.H, x.W, x.C)
        self.add(name, lambda a=args, k=(x, y): _lib.call('sdf_upsample_nearest2', *a, _lib.stream()))

    def im2col_s2(self, name, x, col, pt, pl):
        args = (x.ptr, x.ld, col.data_ptr(), x.Nimg, x.H, x.W, x.C, col.shape[1], col.shape[2], pt, pl)
        self.add(name, lambda a=args, k=(x, col): _lib.call('sdf_im2col_s2', *a, _lib.stream()))

    def col2im_s2(self, name, dcol, dx, pt, pl):
        args = (dcol.data_ptr(), dx.ptr, dx.ld, dx.Nimg, dx.H, dx.W, dx.C, dcol.shape[1], dcol.shape[2], pt, pl)
        self.add(name, lambda a=args, k=(dcol, dx): _lib.call('sdf_col2im_s2', *a, _lib.stream()))

    def copy(self, name, x, y):
        args = (x.ptr, x.ld, y.ptr, y.ld, x.rows, x.C)
        self.add(name, lambda a=args, k=(x, y): _lib.call('sdf_copy2d', *a, _lib.stream()))

    def add2(self, name, a_, b_, y):
        args = (a_.ptr, a_.ld, b_.ptr, b_.ld, y.ptr, y.ld, y.rows, y.C)
        self.add(name, lambda a=args, k=(a_, b_, y): _lib.call('sdf_add2d', *a, _lib.stream()))

    def transpose(self, name, x_t, ldx, y_t, ldy, batch, rows, C):
        args = (x_t.data_ptr(), ldx, y_t.data_ptr(), ldy, batch, rows, C)
        self.add(name, lambda a=args, k=(x_t, y_t): _lib.call('sdf_transpose2d', *a, _lib.stream()))

    def torch_op(self, name, fn):
        """escape hatch for tiny host-scheduled torch ops (scalars, a few KB); never on a hot tensor"""
        self.add(name, fn)


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def _pack_linear(w, device, rows_multiple=1):
    """[out, in] -> fp16 [rows, in_iter] (in rounded up to 64, zero padded)"""
    out_f, in_f = w.shape
    it = _r(in_f, 64)
    rows = _r(out_f, rows_multiple)
    p = torch.zeros(rows, it, device=device, dtype=torch.float16)
    p[:out_f, :in_f] = w.detach().to(device=device, dtype=torch.float16)
    return p


class RunList:
    """A flat list of launches, optionally replayed as a CUDA graph."""

    def __init__(self, ops):
        self.ops = ops
        self.graph = None

    def run(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            for _, fn in self.ops:
                fn()

    def capture(self):
        # warm-up outside capture (sets function attributes, touches every buffer)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _, fn in self.ops:
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _, fn in self.ops:
                fn()
        self.graph = g
        return self
