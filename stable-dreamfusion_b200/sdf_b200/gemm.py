"""Python handle of a tcgen05 implicit-GEMM plan (csrc/sd_gemm.cu, include/sdf_b200.h: sdf_gemm_plan_create)."""
import torch

from . import _lib

ACT = {None: 0, 'none': 0, 'silu': 1, 'gelu': 2, 'geglu': 3}


def pick_block_n(N):
    if N % 160 == 0:
        return 160
    if N <= 64:
        return 64
    return 128


import torch as _torch
NUM_SMS = _torch.cuda.get_device_properties(0).multi_processor_count if _torch.cuda.is_available() else 148      # B200: 148

# Cost model of one plan launch, fitted to tools/gemm_sweep.py on a B200 (profiles/r01_gemm_sweep.txt): the persistent kernel
# retires one 64-deep k-block per CTA every ~0.32 us whatever the tile width (0.44 us for the 256-wide CTA-pair tile), a tile's
# epilogue costs 2.5-7 us when it cannot hide behind the next tile, and split-K adds a memset + a reduction launch.
_T_KB = {(64, 0): 0.30, (128, 0): 0.32, (160, 0): 0.33, (128, 1): 0.33, (160, 1): 0.33, (256, 1): 0.44}
_T_EPI = {64: 2.5, 128: 4.5, 160: 5.5, 256: 7.0}


def estimate_us(M, N, kb, bn, pair, sk):
    m_tiles = (M + 127) // 128
    units = ((m_tiles + 1) // 2 if pair else m_tiles) * ((N + bn - 1) // bn) * sk
    workers = NUM_SMS // 2 if pair else NUM_SMS
    waves = (units + workers - 1) // workers
    kb_per = (kb + sk - 1) // sk
    t = 4.0 + waves * kb_per * _T_KB[(bn, pair)] + _T_EPI[bn] + (waves - 1) * max(0.0, _T_EPI[bn] - kb_per * _T_KB[(bn, pair)])
    if sk > 1:
        t += 5.0 + 2.0 * M * N * 4 * (sk + 1) / 3.0e6          # memset + atomics + reduction pass at ~3 TB/s of L2 traffic
    return t


def choose_config(M, N, kb, batched=False, allow_pair=True):
    """(block_n, cta_pair, splitk) minimising the modelled launch time.  kb = K / 64."""
    cands = [(64, 0), (128, 0)]
    if N % 160 == 0:
        cands.append((160, 0))
    if allow_pair and not batched and M > 128 and N % 256 == 0:
        cands.append((256, 1))
    if N <= 64:
        cands = [(64, 0)]
    best = None
    for bn, pair in cands:
        for sk in (1, 2, 3, 4, 6, 8, 12, 16):
            if sk > 1 and (kb // sk < 8 or batched):
                continue
            t = estimate_us(M, N, kb, bn, pair, sk)
            if best is None or t < best[0] - 1e-9:
                best = (t, bn, pair, sk)
    return best[1], best[2], best[3]


def pick_tile(M, N, batched=False):
    """(block_n, cta_pair) for callers that fix split-K themselves."""
    bn, pair, _ = choose_config(M, N, 64, batched)
    return bn, pair


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
                 bias=None, temb=None, temb_ld=0, residual=None, r_strides=(0, 0, 0), act=None, alpha=1.0, splitk=1, block_n=None, cta_pair=0,
                 stride=1, pad_lo=1):
        """H, W: OUTPUT geometry.  stride = 2 (taps = 9 only): a is [Nimg, 2H, 2W, .] and is read through an element-strided TMA box
        (input pixel = 2 * output pixel + tap - pad_lo): the UNet's Downsample is pad_lo = 1, the VAE encoder's (0,1,0,1)-padded one pad_lo = 0."""
        self.keep = (a, wt, out, bias, temb, residual)
        block_n = pick_block_n(N) if block_n is None else block_n
        self.workspace = torch.empty(Nimg * H * W, N, device=out.device, dtype=torch.float32) if splitk > 1 else None
        self.flops = 2.0 * Nimg * H * W * N * taps * Cin
        self.shape = dict(M=Nimg * H * W, N=N, K=taps * Cin, taps=taps, block_n=block_n, splitk=splitk, pair=int(cta_pair))
        h = _lib.lib().cdll.sdf_gemm_plan_create_strided(
            _lib.ptr(a), *[int(s) for s in a_strides], int(a_c_valid), _lib.ptr(wt), *[int(s) for s in w_strides], int(w_k_valid), int(n_rows_w),
            int(Nimg), int(H), int(W), int(Cin), int(taps), int(N), _lib.ptr(out), *[int(s) for s in o_strides],
            _lib.ptr(bias), _lib.ptr(temb), int(temb_ld), _lib.ptr(residual), *[int(s) for s in r_strides],
            ACT[act], float(alpha), int(splitk), _lib.ptr(self.workspace), int(block_n), int(cta_pair), int(stride), int(pad_lo))
        if h < 0:
            raise RuntimeError(f"sdf_gemm_plan_create_strided failed ({h}): {_lib.lib().last_error()}")
        self.handle = h
        self.gn_slots = 0
        tw = 128 if W >= 128 else W
        th = 1 if (W >= 128 or w_strides[1] != 0 or w_strides[2] != 0) else max(1, min(H, 128 // tw))
        self._stats_ok = (splitk == 1 and N % 32 == 0 and act != 'geglu' and all(int(v) % 8 == 0 for v in o_strides) and (tw * th) % 32 == 0)

    def can_carry_stats(self):
        """may this plan's epilogue accumulate one more consumer's GroupNorm statistics? (csrc/sd_gemm.cu: sdf_gemm_plan_set_gn_stats)"""
        return self._stats_ok and self.gn_slots < 2

    def add_gn_stats(self, stats, channels_per_group, channel_offset):
        rc = _lib.lib().cdll.sdf_gemm_plan_set_gn_stats(self.handle, self.gn_slots, _lib.ptr(stats), int(channels_per_group), int(channel_offset))
        if rc != 0:
            raise RuntimeError(f"sdf_gemm_plan_set_gn_stats failed ({rc}): {_lib.lib().last_error()}")
        self.keep = self.keep + (stats,)
        self.gn_slots += 1

    def run(self):
        _lib.call('sdf_gemm_run', self.handle, _lib.stream())

    def __del__(self):
        try:
            _lib.lib().cdll.sdf_gemm_plan_destroy(self.handle)
        except Exception:
            pass


def conv_plan(a, c_valid, wt, N, out, *, taps, bias=None, temb=None, residual=None, act=None, alpha=1.0, splitk=1, block_n=None, cta_pair=0,
              stride=1, pad_lo=1):
    """a: [Nimg, H * stride, W * stride, lda] fp16 (channels [0, c_valid) are read); wt packed by pack_conv_weight; out: [Nimg, H, W, ldo]."""
    Nimg, Hin, Win, lda = a.shape
    H, W = Hin // stride, Win // stride
    ldo = out.shape[-1]
    cin_iter = wt.shape[1] // taps
    r_str = (0, 0, 0)
    if residual is not None:
        ldr = residual.shape[-1]
        r_str = (ldr, W * ldr, H * W * ldr)
    return GemmPlan(a, (lda, Win * lda, Hin * Win * lda), c_valid, wt, (wt.shape[1], 0, 0), wt.shape[1], wt.shape[0], Nimg, H, W, cin_iter, taps, N,
                    out, (ldo, W * ldo, H * W * ldo), bias=bias, temb=temb, temb_ld=0 if temb is None else temb.shape[-1],
                    residual=residual, r_strides=r_str, act=act, alpha=alpha, splitk=splitk, block_n=block_n, cta_pair=cta_pair,
                    stride=stride, pad_lo=pad_lo)


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
