"""Python handle of a tcgen05 implicit-GEMM plan (csrc/sd_gemm.cu, include/sdf_b200.h: sdf_gemm_plan_create)."""
import torch

from . import _lib

ACT = {None: 0, 'none': 0, 'silu': 1, 'gelu': 2}


def pick_block_n(N):
    if N % 160 == 0:
        return 160
    if N <= 64:
        return 64
    return 128


def pick_tile(M, N, batched=False):
    """(block_n, cta_pair): CTA pairs (tcgen05 cta_group::2, M = 256) whenever the product has at least two M tiles and one shared
    weight matrix — the pair stages each weight tile once for 256 rows, which is what the L2 -> SM fabric needs."""
    if batched or M <= 128 or N <= 64:
        return pick_block_n(N), 0
    if N % 256 == 0:
        return 256, 1
    if N % 160 == 0:
        return 160, 1
    if N % 128 == 0 or N < 256:
        return 128, 1
    return 256, 1


def pack_conv_weight(w, cin_iter=None, rows_multiple=1):
    """[Cout, Cin, kh, kw] (any float dtype) -> fp16 [rows, taps*cin_iter] with K index = tap*cin_iter + c (zero padded)."""
    Cout, Cin, kh, kw = w.shape
    taps = kh * kw
    cin_iter = ((Cin + 63) // 64) * 64 if cin_iter is None else cin_iter
    rows = ((Cout + rows_multiple - 1) // rows_multiple) * rows_multiple
    out = torch.zeros(rows, taps, cin_iter, device=w.device, dtype=torch.float16)
    out[:Cout, :, :Cin] = w.permute(0, 2, 3, 1).reshape(Cout, taps, Cin).to(torch.float16)
    return out.reshape(rows, taps * cin_iter)


class GemmPlan:
    """Keeps the operand tensors alive and owns the native plan handle."""

    def __init__(self, a, a_strides, a_c_valid, wt, w_strides, w_k_valid, n_rows_w, Nimg, H, W, Cin, taps, N, out, o_strides,
                 bias=None, temb=None, temb_ld=0, residual=None, r_strides=(0, 0, 0), act=None, alpha=1.0, splitk=1, block_n=None, cta_pair=0):
        self.keep = (a, wt, out, bias, temb, residual)
        block_n = pick_block_n(N) if block_n is None else block_n
        self.workspace = torch.empty(Nimg * H * W, N, device=out.device, dtype=torch.float32) if splitk > 1 else None
        self.flops = 2.0 * Nimg * H * W * N * taps * Cin
        self.shape = dict(M=Nimg * H * W, N=N, K=taps * Cin, taps=taps, block_n=block_n, splitk=splitk, pair=int(cta_pair))
        h = _lib.lib().cdll.sdf_gemm_plan_create(
            _lib.ptr(a), *[int(s) for s in a_strides], int(a_c_valid), _lib.ptr(wt), *[int(s) for s in w_strides], int(w_k_valid), int(n_rows_w),
            int(Nimg), int(H), int(W), int(Cin), int(taps), int(N), _lib.ptr(out), *[int(s) for s in o_strides],
            _lib.ptr(bias), _lib.ptr(temb), int(temb_ld), _lib.ptr(residual), *[int(s) for s in r_strides],
            ACT[act], float(alpha), int(splitk), _lib.ptr(self.workspace), int(block_n), int(cta_pair))
        if h < 0:
            raise RuntimeError(f"sdf_gemm_plan_create failed ({h}): {_lib.lib().last_error()}")
        self.handle = h

    def run(self):
        _lib.call('sdf_gemm_run', self.handle, _lib.stream())

    def __del__(self):
        try:
            _lib.lib().cdll.sdf_gemm_plan_destroy(self.handle)
        except Exception:
            pass


def conv_plan(a, c_valid, wt, N, out, *, taps, bias=None, temb=None, residual=None, act=None, alpha=1.0, splitk=1, block_n=None, cta_pair=0):
    """a: [Nimg, H, W, lda] fp16 (channels [0, c_valid) are read); wt packed by pack_conv_weight; out: [Nimg, H, W, ldo]."""
    Nimg, H, W, lda = a.shape
    ldo = out.shape[-1]
    cin_iter = wt.shape[1] // taps
    r_str = (0, 0, 0)
    if residual is not None:
        ldr = residual.shape[-1]
        r_str = (ldr, W * ldr, H * W * ldr)
    return GemmPlan(a, (lda, W * lda, H * W * lda), c_valid, wt, (wt.shape[1], 0, 0), wt.shape[1], wt.shape[0], Nimg, H, W, cin_iter, taps, N,
                    out, (ldo, W * ldo, H * W * ldo), bias=bias, temb=temb, temb_ld=0 if temb is None else temb.shape[-1],
                    residual=residual, r_strides=r_str, act=act, alpha=alpha, splitk=splitk, block_n=block_n, cta_pair=cta_pair)


def linear_plan(a, k_valid, wt, N, out, *, bias=None, residual=None, act=None, alpha=1.0, splitk=1, block_n=None, cta_pair=0):
    """a: [rows, lda]; wt: [n_rows, K_iter] fp16; out: [rows, ldo]."""
    rows, lda = a.shape
    ldo = out.shape[-1]
    r_str = (0, 0, 0)
    if residual is not None:
        ldr = residual.shape[-1]
        r_str = (ldr, rows * ldr, rows * ldr)
    return GemmPlan(a, (lda, rows * lda, rows * lda), k_valid, wt, (wt.shape[1], 0, 0), min(k_valid, wt.shape[1]), wt.shape[0], 1, 1, rows,
                    ((k_valid + 63) // 64) * 64, 1, N, out, (ldo, rows * ldo, rows * ldo), bias=bias, residual=residual, r_strides=r_str,
                    act=act, alpha=alpha, splitk=splitk, block_n=block_n, cta_pair=cta_pair)
