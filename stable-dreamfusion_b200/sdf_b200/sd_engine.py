"""SD-1.5-shaped UNet (forward) and VAE encoder (forward + data-gradient) as a static list of native launches.

Every dense contraction is a tcgen05 plan (csrc/sd_gemm.cu) — 3x3 / 1x1 convolutions as implicit GEMM over NHWC fp16
activations and linear layers (q|k|v / k|v projections fused into one plan, GEGLU fused into its projection's epilogue) — the
UNet attention is the fused flash kernel of csrc/flash_attn.cu straight on the [tokens, heads*d] projections, everything is
glued by the memory-bound kernels of csrc/sd_ops.cu.  Shapes are fixed at construction, so the whole step is a flat
list of ~500 launches with no Python tensor ops in between; `capture()` records it into a CUDA graph.

Structure follows the modules the reference executes (vendored CompVis code; see oracle/sd_ref.py for the citations):
ResBlock / SpatialTransformer / BasicTransformerBlock / CrossAttention / GEGLU / Downsample / Upsample / UNetModel,
and the VAE Encoder (ResnetBlock / AttnBlock / Downsample) + quant_conv.  State-dict keys are the CompVis ones.

Layout conventions
  activations   NHWC fp16, a `View` = (tensor [Nimg,H,W,ld], channel offset, channels); skip connections are written
                by their producer directly into the consumer's concat buffer (no torch.cat kernels)
  conv weights  [Cout, tap, Cin_iter] fp16, K index = tap*Cin_iter + c (Cin_iter = Cin rounded up to 64)
  attention     UNet: sdf_flash_attention on q / k / v views of the fused projection buffers (no score matrix in HBM).
                VAE mid block (one head, d = 512, forward + backward): S = Q K^T and O = P V as batched tcgen05 plans with a
                row-softmax kernel in between
"""
import math

import torch

from . import _lib
import os

from .gemm import GemmPlan, choose_config, conv_plan, linear_plan, pack_conv_weight, pick_block_n

NUM_SMS = torch.cuda.get_device_properties(0).multi_processor_count if torch.cuda.is_available() else 148      # B200: 148
USE_CTA_PAIRS = os.environ.get('SDF_GEMM_CTA_PAIRS', '1') != '0'
# GroupNorm statistics in the producing GEMM's epilogue.  OFF by default: measured on B200 (gpurun_out/r2_bench4*.log, launches_r02.csv) the
# statistics-carrying epilogue costs more than the separate statistics pass it removes (k_gemm<128> 107 us vs 26 us: the per-chunk group
# reduction runs on the critical path of single-tile CTAs and spills); kept behind the switch with its parity tests.
FUSE_GN_STATS = os.environ.get('SDF_FUSE_GN_STATS', '0') != '0'
# A/B switch: FMA-pipe conv_in / conv_in^T kernels instead of the zero-padded implicit GEMM.  Measured (tools/bench_conv_in.py, 512x512, L2 flushed):
# forward 210 us direct vs 92 us GEMM, data-gradient 248 us vs 123 us -> the tcgen05 path wins even at 3/64 useful k-columns; off by default.
DIRECT_CONV_IN = os.environ.get('SDF_DIRECT_CONV_IN', '0') != '0'
FUSE_GEGLU = os.environ.get('SDF_FUSE_GEGLU', '1') != '0'                 # A/B switch: GEGLU in the projection GEMM's epilogue      # A/B switch for the cta_group::2 GEMM variant
STRIDED_TMA_CONV = os.environ.get('SDF_STRIDED_TMA_CONV', '1') != '0'    # A/B switch: stride-2 convolutions through an element-strided TMA box (0: im2col + 1-tap GEMM)


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


def _choose_splitk(M, N, kblocks, block_n, pair=0):
    """split-K only when the output tiles cannot fill a third of the SMs (CTA pairs: of the 74 pairs) AND K is long (3x3 convs at
    16x16 / 8x8, the M=2 embedding products): every split costs a memset + a reduction launch."""
    m_tiles = (M + 127) // 128
    units = ((m_tiles + 1) // 2 if pair else m_tiles) * ((N + block_n - 1) // block_n)
    workers = NUM_SMS // 2 if pair else NUM_SMS
    if units > workers // 3 or kblocks < 16:
        return 1
    return max(1, min(workers // units, kblocks // 8))


class Builder:
    """Accumulates launches.  Each op is a zero-argument callable; tensors it touches are kept alive by closures."""

    def __init__(self, device):
        self.device = device
        self.ops = []
        self.flops = 0.0
        self.flops_attn = 0.0
        self.bytes_act = 0
        # which GEMM plan last wrote which channel range of which tensor (GroupNorm statistics ride in that plan's epilogue)
        self.producers = {}
        self.gn_arena = None             # fp32 arena of all epilogue-accumulated statistics, zeroed by the first op of the list
        self.gn_used = 0
        self.gn_fused = 0

    # ---- producer tracking for the GroupNorm-statistics fusion
    def _touch(self, view, plan=None, N=None, geom=None):
        """record `plan` as the writer of channels [off, off+N) of view.t (or forget the range when a non-GEMM kernel writes it)"""
        key = view.t.data_ptr()
        lo, hi = view.off, view.off + (view.C if N is None else N)
        ent = [e for e in self.producers.get(key, []) if e['hi'] <= lo or e['lo'] >= hi]
        if plan is not None:
            ent.append(dict(lo=lo, hi=hi, plan=plan, geom=geom))
        self.producers[key] = ent

    def _arena_stats(self, Nimg):
        """[Nimg, 32, 2] fp32 statistics slot inside ONE arena that the first op of the list zeroes (instead of a memset node per norm)"""
        n = Nimg * 32 * 2
        if self.gn_arena is None:
            self.gn_arena = torch.zeros(64 * 1024, device=self.device, dtype=torch.float32)
            arena = self.gn_arena
            self.ops.insert(0, ('zero_gn_stats', lambda a=arena: a.zero_()))
        if self.gn_used + n > self.gn_arena.numel():
            return None
        stats = self.gn_arena[self.gn_used:self.gn_used + n].view(Nimg, 32, 2)
        self.gn_used += n
        return stats

    def _gn_stats_from_producers(self, x):
        """-> stats tensor filled by the epilogues of the plans that wrote x, or None when x is not completely covered by such plans"""
        if not FUSE_GN_STATS:
            return None
        ent = sorted([e for e in self.producers.get(x.t.data_ptr(), []) if e['lo'] >= x.off and e['hi'] <= x.off + x.C], key=lambda e: e['lo'])
        pos = x.off
        for e in ent:
            if e['lo'] != pos or e['geom'] != (x.Nimg, x.H, x.W) or not e['plan'].can_carry_stats():
                return None
            pos = e['hi']
        if pos != x.off + x.C or not ent:
            return None
        stats = self._arena_stats(x.Nimg)
        if stats is None:
            return None
        cpg = x.C // 32
        for e in ent:
            e['plan'].add_gn_stats(stats, cpg, e['lo'] - x.off)
        self.gn_fused += 1
        return stats

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
             r_strides=None, cin_iter=None, stride=1, pad_lo=1):
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
        M = Nimg * H * W
        batched = w_str[1] != 0 or w_str[2] != 0
        kb = taps * cin_iter // 64
        if block_n is None and splitk is None:
            bn, pair, sk = choose_config(M, N, kb, batched, allow_pair=USE_CTA_PAIRS)
        elif block_n is None and splitk == 1:
            bn, pair, sk = choose_config(M, N, kb, True, allow_pair=False)[0], 0, 1      # batched=True disables split-K in the model
            if USE_CTA_PAIRS and not batched and M > 128 and N % 256 == 0:
                bn, pair = 256, 1
        else:
            bn, pair = (pick_block_n(N) if block_n is None else block_n), 0
            sk = _choose_splitk(M, N, kb, bn, pair) if splitk is None else splitk
        wrap = lambda v: _PtrTensor(v) if isinstance(v, View) else v
        plan = GemmPlan(wrap(a), a_str, c_valid, wrap(wt), w_str, (taps * cin_iter if w_k_valid is None else w_k_valid),
                        (wt.shape[0] if n_rows_w is None else n_rows_w), Nimg, H, W, cin_iter, taps, N, wrap(out), o_str, bias=bias,
                        temb=wrap(temb) if temb is not None else None, temb_ld=temb_ld, residual=wrap(residual) if residual is not None else None,
                        r_strides=r_str, act=act, alpha=alpha, splitk=sk, block_n=bn, cta_pair=pair, stride=stride, pad_lo=pad_lo)
        self.flops += 2.0 * M * N * taps * c_valid
        self.add(name, plan.run)
        if isinstance(out, View):
            n_out = N // 2 if act == 'geglu' else N
            self._touch(out, plan if (o_strides is None and geom is None and act != 'geglu') else None, n_out, (Nimg, H, W))
        return plan

    # ---- memory-bound
    def groupnorm(self, name, x, y, gamma, beta, eps, silu, stats=None):
        fused = self._gn_stats_from_producers(x) if stats is None else None
        fn = 'sdf_groupnorm_forward'
        if fused is not None:
            stats, fn = fused, 'sdf_groupnorm_apply'      # statistics accumulated by the producing GEMMs' epilogues: normalise(+SiLU) only
        elif stats is None:
            stats = self._arena_stats(x.Nimg)             # zeroed once per list run together with every other norm's statistics
            if stats is not None:
                fn = 'sdf_groupnorm_forward_prezeroed'
            else:
                stats = self.f32(x.Nimg, 32, 2)
        args = (x.ptr, x.ld, y.ptr, y.ld, x.Nimg, x.H * x.W, x.C, 32, _lib.ptr(gamma), _lib.ptr(beta), float(eps), int(silu), _lib.ptr(stats))
        keep = (x, y, gamma, beta, stats)
        self.add(name + ('(apply)' if fused is not None else ''), lambda a=args, k=keep, f=fn: _lib.call(f, *a, _lib.stream()))
        self._touch(y)
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
        self._touch(y)

    def softmax(self, name, s, rows, cols, ld, scale=1.0):
        args = (s.data_ptr(), s.data_ptr(), rows, cols, ld, float(scale))
        self.add(name, lambda a=args, k=s: _lib.call('sdf_softmax_rows', *a, _lib.stream()))

    def flash_attention(self, name, q, k, v, o, B, heads, n, nkv, d):
        args = (q.ptr, k.ptr, v.ptr, o.ptr, B, heads, n, nkv, d, q.ld, k.ld, o.ld, float(d) ** -0.5)
        assert k.ld == v.ld
        self.flops_attn += 4.0 * B * heads * n * nkv * d
        self.add(name, lambda a=args, keep=(q, k, v, o): _lib.call('sdf_flash_attention', *a, _lib.stream()))
        self._touch(o)

    def softmax_bwd(self, name, p, dp, ds, rows, cols, ld, scale):
        args = (p.data_ptr(), dp.data_ptr(), ds.data_ptr(), rows, cols, ld, float(scale))
        self.add(name, lambda a=args, k=(p, dp, ds): _lib.call('sdf_softmax_rows_backward', *a, _lib.stream()))

    def geglu(self, name, x, y, inner):
        args = (x.ptr, x.ld, y.ptr, y.ld, x.rows, inner)
        self.add(name, lambda a=args, k=(x, y): _lib.call('sdf_geglu', *a, _lib.stream()))
        self._touch(y)

    def upsample2(self, name, x, y):
        args = (x.ptr, x.ld, y.ptr, y.ld, x.Nimg, x.H, x.W, x.C)
        self.add(name, lambda a=args, k=(x, y): _lib.call('sdf_upsample_nearest2', *a, _lib.stream()))
        self._touch(y)

    def conv3x3_small_cin(self, name, x, cin, w, bias, y):
        """direct 3x3 convolution whose input has <= 4 real channels (the VAE's conv_in); w fp32 [Cout, Cin, 3, 3]"""
        args = (x.ptr, x.ld, w.data_ptr(), bias.data_ptr() if bias is not None else None, y.ptr, y.ld, x.Nimg, x.H, x.W, cin, y.C)
        self.flops += 2.0 * x.rows * 9 * cin * y.C
        self.add(name, lambda a=args, k=(x, w, bias, y): _lib.call('sdf_conv3x3_small_cin_forward', *a, _lib.stream()))
        self._touch(y)

    def conv3x3_small_cin_dgrad(self, name, dy, w, cin, dx):
        args = (dy.ptr, dy.ld, w.data_ptr(), dx.ptr, dx.ld, dy.Nimg, dy.H, dy.W, cin, dy.C)
        self.flops += 2.0 * dy.rows * 9 * cin * dy.C
        self.add(name, lambda a=args, k=(dy, w, dx): _lib.call('sdf_conv3x3_small_cin_dgrad', *a, _lib.stream()))
        self._touch(dx)

    def im2col_s2(self, name, x, col, pt, pl):
        args = (x.ptr, x.ld, col.data_ptr(), x.Nimg, x.H, x.W, x.C, col.shape[1], col.shape[2], pt, pl)
        self.add(name, lambda a=args, k=(x, col): _lib.call('sdf_im2col_s2', *a, _lib.stream()))

    def col2im_s2(self, name, dcol, dx, pt, pl):
        args = (dcol.data_ptr(), dx.ptr, dx.ld, dx.Nimg, dx.H, dx.W, dx.C, dcol.shape[1], dcol.shape[2], pt, pl)
        self.add(name, lambda a=args, k=(dcol, dx): _lib.call('sdf_col2im_s2', *a, _lib.stream()))
        self._touch(dx)

    def copy(self, name, x, y):
        args = (x.ptr, x.ld, y.ptr, y.ld, x.rows, x.C)
        self.add(name, lambda a=args, k=(x, y): _lib.call('sdf_copy2d', *a, _lib.stream()))
        self._touch(y)

    def add2(self, name, a_, b_, y):
        args = (a_.ptr, a_.ld, b_.ptr, b_.ld, y.ptr, y.ld, y.rows, y.C)
        self.add(name, lambda a=args, k=(a_, b_, y): _lib.call('sdf_add2d', *a, _lib.stream()))
        self._touch(y)

    def transpose(self, name, x_t, ldx, y_t, ldy, batch, rows, C):
        args = (x_t.data_ptr(), ldx, y_t.data_ptr(), ldy, batch, rows, C)
        self.add(name, lambda a=args, k=(x_t, y_t): _lib.call('sdf_transpose2d', *a, _lib.stream()))

    def torch_op(self, name, fn):
        """escape hatch for tiny host-scheduled torch ops (scalars, a few KB); never on a hot tensor"""
        self.add(name, fn)


def _f32(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def geglu_row_permutation(inner, device):
    """row order of a [2*inner, k] GEGLU projection for the fused epilogue: chunk c = rows [16c, 16c+16) of the value half followed
    by the same rows of the gate half"""
    idx = torch.arange(inner, device=device).view(-1, 16)
    return torch.cat([idx, idx + inner], dim=1).reshape(-1)


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


# =============================================================================================== UNet
UNET_SD15 = dict(in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2, attention_resolutions=(4, 2, 1),
                 channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768)


def unet_structure(cfg):
    """Block lists of the CompVis UNetModel constructor (openaimodel.py:530-712): every entry is
    ('conv_in', cin, cout) | ('res', cin, cout) | ('attn', ch) | ('down', ch) | ('up', ch)."""
    mc, mult, nrb, ar = cfg['model_channels'], cfg['channel_mult'], cfg['num_res_blocks'], cfg['attention_resolutions']
    inp = [[('conv_in', cfg['in_channels'], mc)]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [('res', ch, m * mc)]
            ch = m * mc
            if ds in ar:
                layers.append(('attn', ch))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            inp.append([('down', ch)])
            chans.append(ch)
            ds *= 2
    mid = [('res', ch, ch), ('attn', ch), ('res', ch, ch)]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [('res', ch + ich, mc * m)]
            skip = (ch, ich)
            ch = mc * m
            if ds in ar:
                layers.append(('attn', ch))
            if level and i == nrb:
                layers.append(('up', ch))
                ds //= 2
            out.append((layers, skip))
    return inp, mid, out


class UNetEngine:
    """Forward pass of the SD-1.5-shaped UNet for a fixed (batch, latent size, context length).

    inputs  : self.x_in  [B, hw, hw, 8] fp16 (channels 0..3 = noisy latents), self.t_in int32 [B], self.ctx [1,1,B*ctx_len,ctx_dim] fp16
    output  : self.eps   [B, hw, hw, 8] fp16 (channels 0..3 = predicted noise)
    """

    def __init__(self, sd, device, cfg=UNET_SD15, batch=2, hw=64, ctx_len=77):
        self.cfg, self.B, self.hw, self.ctx_len = cfg, batch, hw, ctx_len
        self.dev = device
        b = self.b = Builder(device)
        self.sd = sd
        self._wcache = {}
        self._sbuf = {}
        B, mc = batch, cfg['model_channels']
        ted = 4 * mc
        heads = cfg['num_heads']
        self.heads = heads
        inp, mid, out = unet_structure(cfg)

        self.x_in = b.buf(B, hw, hw, 8, zero=True)
        self.t_in = torch.zeros(B, dtype=torch.int32, device=device)
        self.ctx = b.buf(1, 1, B * ctx_len, cfg['context_dim'], zero=True)
        ctxv = View(self.ctx)

        # ---- time embedding: sinusoid -> Linear+SiLU -> Linear (+ the SiLU that opens every emb_layers)
        t0 = b.buf(1, 1, B, mc)
        b.add('t_emb', lambda: _lib.call('sdf_timestep_embedding', _lib.ptr(self.t_in), B, mc, t0.data_ptr(), mc, _lib.stream()))
        e1 = b.buf(1, 1, B, ted)
        b.gemm('time_embed.0', View(t0), mc, self._lin('time_embed.0.weight'), ted, View(e1), bias=self._f32('time_embed.0.bias'), act='silu')
        emb = b.buf(1, 1, B, ted)
        b.gemm('time_embed.2', View(e1), ted, self._lin('time_embed.2.weight'), ted, View(emb), bias=self._f32('time_embed.2.bias'), act='silu')
        # all ResBlock emb_layers as ONE product
        res_prefixes = []
        for i, layers in enumerate(inp):
            for j, l in enumerate(layers):
                if l[0] == 'res':
                    res_prefixes.append((f'input_blocks.{i}.{j}', l[2]))
        for j, l in enumerate(mid):
            if l[0] == 'res':
                res_prefixes.append((f'middle_block.{j}', l[2]))
        for i, (layers, _) in enumerate(out):
            for j, l in enumerate(layers):
                if l[0] == 'res':
                    res_prefixes.append((f'output_blocks.{i}.{j}', l[2]))
        w_all = torch.cat([sd[p + '.emb_layers.1.weight'] for p, _ in res_prefixes], 0)
        b_all = torch.cat([sd[p + '.emb_layers.1.bias'] for p, _ in res_prefixes], 0)
        sumc = w_all.shape[0]
        self.temb_all = b.buf(1, 1, B, sumc)
        b.gemm('emb_layers(all)', View(emb), ted, _pack_linear(w_all, device), sumc, View(self.temb_all), bias=_f32(b_all, device))
        self.temb_off = {}
        off = 0
        for p, c in res_prefixes:
            self.temb_off[p] = off
            off += c

        # ---- every cross-attention's k|v projection of the (layer-independent) context as ONE product: 16 launches of M = B*ctx_len rows
        #      (154: two row tiles each) become one with a few hundred column tiles
        attn_prefixes = []
        for i, layers in enumerate(inp):
            attn_prefixes += [(f'input_blocks.{i}.{j}', l[1]) for j, l in enumerate(layers) if l[0] == 'attn']
        attn_prefixes += [(f'middle_block.{j}', l[1]) for j, l in enumerate(mid) if l[0] == 'attn']
        for i, (layers, _) in enumerate(out):
            attn_prefixes += [(f'output_blocks.{i}.{j}', l[1]) for j, l in enumerate(layers) if l[0] == 'attn']
        self.kv_off = {}
        if attn_prefixes and os.environ.get('SDF_HOIST_CTX_KV', '1') != '0':
            ws_, off = [], 0
            for p_, c_ in attn_prefixes:
                a2 = p_ + '.transformer_blocks.0.attn2'
                ws_ += [self._lin(a2 + '.to_k.weight'), self._lin(a2 + '.to_v.weight')]
                self.kv_off[a2] = off
                off += 2 * c_
            self.kv_all = View(b.buf(1, 1, B * ctx_len, off))
            b.gemm('attn2.to_kv(all)', View(self.ctx), cfg['context_dim'], torch.cat(ws_, 0).contiguous(), off, self.kv_all)

        # ---- buffers for the skip concatenations: output block k reads cat_k = [h (ch) | skip (ich)]
        # spatial size of every input block's output
        sizes = []
        s = hw
        for layers in inp:
            if layers[0][0] == 'down':
                s //= 2
            sizes.append(s)
        cats = []
        for k, (layers, (chh, ich)) in enumerate(out):
            j = len(inp) - 1 - k
            cats.append(b.buf(B, sizes[j], sizes[j], chh + ich))
        self.cats = cats

        def skip_view(j):      # where input block j must leave its output
            k = len(inp) - 1 - j
            chh, ich = out[k][1]
            return View(cats[k], chh, ich)

        # ---- down path
        h = None
        for i, layers in enumerate(inp):
            dst = skip_view(i)
            for j, l in enumerate(layers):
                last = j == len(layers) - 1
                p = f'input_blocks.{i}.{j}'
                if l[0] == 'conv_in':
                    w = pack_conv_weight(sd[p + '.weight'].to(device))
                    b.gemm(p, View(self.x_in), l[1], w, l[2], dst, taps=9, bias=self._f32(p + '.bias'))
                    h = dst
                elif l[0] == 'res':
                    o = dst if last else View(b.buf(B, h.H, h.W, l[2]))
                    self._resblock(p, h, l[2], o)
                    h = o
                elif l[0] == 'attn':
                    o = dst if last else View(b.buf(B, h.H, h.W, l[1]))
                    self._transformer(p, h, o, ctxv)
                    h = o
                elif l[0] == 'down':
                    self._down(p, h, dst)
                    h = dst
        # ---- middle
        dst0 = View(cats[0], 0, out[0][1][0])
        m0 = View(b.buf(B, h.H, h.W, mid[0][2]))
        self._resblock('middle_block.0', h, mid[0][2], m0)
        m1 = View(b.buf(B, h.H, h.W, mid[1][1]))
        self._transformer('middle_block.1', m0, m1, ctxv)
        self._resblock('middle_block.2', m1, mid[2][2], dst0)
        # ---- up path
        final = None
        for k, (layers, (chh, ich)) in enumerate(out):
            h = View(cats[k])
            last_block = k == len(out) - 1
            nxt = None if last_block else View(cats[k + 1], 0, out[k + 1][1][0])
            for j, l in enumerate(layers):
                last = j == len(layers) - 1
                p = f'output_blocks.{k}.{j}'
                if last and not last_block:
                    o = nxt
                elif l[0] == 'up':
                    o = View(b.buf(B, 2 * h.H, 2 * h.W, l[1]))
                else:
                    o = View(b.buf(B, h.H, h.W, l[2] if l[0] == 'res' else l[1]))
                if l[0] == 'res':
                    self._resblock(p, h, l[2], o)
                elif l[0] == 'attn':
                    self._transformer(p, h, o, ctxv)
                else:
                    self._up(p, h, o)
                h = o
            final = h
        # ---- out: GroupNorm32 + SiLU + conv3x3
        tfin = View(b.buf(B, final.H, final.W, final.C))
        b.groupnorm('out.0', final, tfin, self._f32('out.0.weight'), self._f32('out.0.bias'), 1e-5, True)
        self.eps = b.buf(B, hw, hw, 8, zero=True)
        b.gemm('out.2', tfin, final.C, pack_conv_weight(sd['out.2.weight'].to(device)), cfg['out_channels'], View(self.eps), taps=9,
               bias=self._f32('out.2.bias'))
        self.runlist = RunList(b.ops)
        self.gn_fused = b.gn_fused          # GroupNorms whose statistics ride in their producers' epilogues
        self.flops = b.flops + b.flops_attn
        self.flops_gemm, self.flops_attn = b.flops, b.flops_attn
        self.sd = None          # fp32 originals are no longer needed
        self._wcache = None

    # ---------------------------------------------------------------- helpers
    def _f32(self, key):
        return _f32(self.sd[key], self.dev)

    def _lin(self, key, rows_multiple=1):
        return _pack_linear(self.sd[key], self.dev, rows_multiple)

    def _conv1x1_w(self, key):
        w = self.sd[key]
        return _pack_linear(w.reshape(w.shape[0], -1), self.dev)

    def _resblock(self, p, x, cout, out):
        b, sd, dev, B = self.b, self.sd, self.dev, self.B
        cin = x.C
        t1 = View(b.buf(B, x.H, x.W, cin))
        b.groupnorm(p + '.in_layers.0', x, t1, self._f32(p + '.in_layers.0.weight'), self._f32(p + '.in_layers.0.bias'), 1e-5, True)
        h1 = View(b.buf(B, x.H, x.W, cout))
        temb = View(self.temb_all, self.temb_off[p], cout)
        b.gemm(p + '.in_layers.2', t1, cin, pack_conv_weight(sd[p + '.in_layers.2.weight'].to(dev)), cout, h1, taps=9,
               bias=self._f32(p + '.in_layers.2.bias'), temb=temb, temb_ld=self.temb_all.shape[-1])
        t2 = View(b.buf(B, x.H, x.W, cout))
        b.groupnorm(p + '.out_layers.0', h1, t2, self._f32(p + '.out_layers.0.weight'), self._f32(p + '.out_layers.0.bias'), 1e-5, True)
        if cin != cout:
            xs = View(b.buf(B, x.H, x.W, cout))
            b.gemm(p + '.skip_connection', x, cin, self._conv1x1_w(p + '.skip_connection.weight'), cout, xs, bias=self._f32(p + '.skip_connection.bias'))
        else:
            xs = x
        b.gemm(p + '.out_layers.3', t2, cout, pack_conv_weight(sd[p + '.out_layers.3.weight'].to(dev)), cout, out, taps=9,
               bias=self._f32(p + '.out_layers.3.bias'), residual=xs)

    def _lin_cat(self, keys):
        """rows of several linear weights stacked (one GEMM produces q|k|v or k|v side by side)"""
        return torch.cat([self._lin(k) for k in keys], 0).contiguous()

    def _attention(self, p, ln, kv_src, kv_rows_per_batch, kv_dim, u, C):
        """u += to_out(softmax(q k^T / sqrt(d)) v);  q from ln [B*n, C]; k, v from kv_src ([B*kv_rows, kv_dim] View).
        Projections are one GEMM (q|k|v for self-attention, q and k|v for cross-attention); the attention itself is the fused
        flash kernel on the [tokens, heads*d] layout (ldm/modules/attention.py:170-193)."""
        b, B, heads = self.b, self.B, self.heads
        n = ln.H * ln.W if ln.Nimg == B else ln.rows // B
        d = C // heads
        nkv = kv_rows_per_batch
        lnf = View(ln.t.view(1, 1, B * n, ln.ld), ln.off, ln.C)
        if kv_src is ln:
            qkv = View(b.buf(1, 1, B * n, 3 * C))
            b.gemm(p + '.to_qkv', lnf, C, self._lin_cat([p + '.to_q.weight', p + '.to_k.weight', p + '.to_v.weight']), 3 * C, qkv)
            q, k, v = qkv.sub(0, C), qkv.sub(C, C), qkv.sub(2 * C, C)
        else:
            q = View(b.buf(1, 1, B * n, C))
            b.gemm(p + '.to_q', lnf, C, self._lin(p + '.to_q.weight'), C, q)
            if p in self.kv_off:         # projected once for all layers at the top of the list
                k, v = self.kv_all.sub(self.kv_off[p], C), self.kv_all.sub(self.kv_off[p] + C, C)
            else:
                kvf = View(kv_src.t.view(1, 1, B * nkv, kv_src.ld), kv_src.off, kv_src.C)
                kv = View(b.buf(1, 1, B * nkv, 2 * C))
                b.gemm(p + '.to_kv', kvf, kv_dim, self._lin_cat([p + '.to_k.weight', p + '.to_v.weight']), 2 * C, kv)
                k, v = kv.sub(0, C), kv.sub(C, C)
        o = View(b.buf(1, 1, B * n, C))
        b.flash_attention(p + '.attention', q, k, v, o, B, heads, n, nkv, d)
        uf = View(u.t.view(1, 1, B * n, u.ld), u.off, u.C)
        b.gemm(p + '.to_out', o, C, self._lin(p + '.to_out.0.weight'), C, uf, bias=self._f32(p + '.to_out.0.bias'), residual=uf)

    def _transformer(self, p, x, out, ctxv):
        b, B, dev = self.b, self.B, self.dev
        C = x.C
        t = View(b.buf(B, x.H, x.W, C))
        b.groupnorm(p + '.norm', x, t, self._f32(p + '.norm.weight'), self._f32(p + '.norm.bias'), 1e-6, False)
        u = View(b.buf(B, x.H, x.W, C))
        b.gemm(p + '.proj_in', t, C, self._conv1x1_w(p + '.proj_in.weight'), C, u, bias=self._f32(p + '.proj_in.bias'))
        ln = View(b.buf(B, x.H, x.W, C))
        tb = p + '.transformer_blocks.0'
        b.layernorm(tb + '.norm1', u, ln, self._f32(tb + '.norm1.weight'), self._f32(tb + '.norm1.bias'))
        self._attention(tb + '.attn1', ln, ln, x.H * x.W, C, u, C)
        b.layernorm(tb + '.norm2', u, ln, self._f32(tb + '.norm2.weight'), self._f32(tb + '.norm2.bias'))
        self._attention(tb + '.attn2', ln, ctxv, self.ctx_len, self.cfg['context_dim'], u, C)
        b.layernorm(tb + '.norm3', u, ln, self._f32(tb + '.norm3.weight'), self._f32(tb + '.norm3.bias'))
        inner = 4 * C
        n = x.H * x.W
        lnf = View(ln.t.view(1, 1, B * n, C))
        gg = View(b.buf(1, 1, B * n, inner))
        if FUSE_GEGLU and inner % 16 == 0:
            # GEGLU inside the projection's epilogue (ldm/modules/attention.py:37-45): weight / bias rows interleaved in 32-row chunks
            # [16 value | 16 gate], so the [tokens, 2*inner] intermediate (42 MB per 64x64 layer) never exists
            perm = geglu_row_permutation(inner, self.dev)
            wp = self._lin(tb + '.ff.net.0.proj.weight')[perm].contiguous()
            bp = self._f32(tb + '.ff.net.0.proj.bias')[perm].contiguous()
            b.gemm(tb + '.ff.net.0.proj+geglu', lnf, C, wp, 2 * inner, gg, bias=bp, act='geglu', splitk=1,
                   block_n=None)
        else:
            f = View(b.buf(1, 1, B * n, 2 * inner))
            b.gemm(tb + '.ff.net.0.proj', lnf, C, self._lin(tb + '.ff.net.0.proj.weight'), 2 * inner, f, bias=self._f32(tb + '.ff.net.0.proj.bias'))
            b.geglu(tb + '.ff.geglu', f, gg, inner)
        uf = View(u.t.view(1, 1, B * n, C))
        b.gemm(tb + '.ff.net.2', gg, inner, self._lin(tb + '.ff.net.2.weight'), C, uf, bias=self._f32(tb + '.ff.net.2.bias'), residual=uf)
        b.gemm(p + '.proj_out', u, C, self._conv1x1_w(p + '.proj_out.weight'), C, out, bias=self._f32(p + '.proj_out.bias'), residual=x)

    def _down(self, p, x, out):
        b, B = self.b, self.B
        C = x.C
        w4 = self.sd[p + '.op.weight'].to(self.dev)
        if STRIDED_TMA_CONV:
            # 3x3, stride 2, pad 1 (openaimodel.py:130-138) read straight out of x through an element-strided TMA box: no im2col buffer
            b.gemm(p + '.op', x, C, pack_conv_weight(w4), C, out, taps=9, bias=self._f32(p + '.op.bias'), geom=(B, x.H // 2, x.W // 2),
                   a_strides=x.strides(), stride=2, pad_lo=1)
            return
        col = torch.empty(B, x.H // 2, x.W // 2, 9 * C, device=self.dev, dtype=torch.float16)
        b.im2col_s2(p + '.im2col', x, col, 1, 1)
        w = _pack_linear(w4.permute(0, 2, 3, 1).reshape(w4.shape[0], 9 * C), self.dev)                # [Cout, C, 3, 3] -> [Cout, tap*C + c]
        b.gemm(p + '.op', View(col), 9 * C, w, C, out, bias=self._f32(p + '.op.bias'))

    def _up(self, p, x, out):
        b, B = self.b, self.B
        up = View(b.buf(B, 2 * x.H, 2 * x.W, x.C))
        b.upsample2(p + '.nearest', x, up)
        b.gemm(p + '.conv', up, x.C, pack_conv_weight(self.sd[p + '.conv.weight'].to(self.dev)), x.C, out, taps=9, bias=self._f32(p + '.conv.bias'))

    # ---------------------------------------------------------------- API
    def set_inputs(self, latents_nchw, t, context):
        """latents [B,4,hw,hw], t int [B], context [B, ctx_len, ctx_dim] (any float dtype, on the engine's device)."""
        self.x_in[..., :latents_nchw.shape[1]].copy_(latents_nchw.permute(0, 2, 3, 1))
        self.t_in.copy_(t.to(torch.int32))
        self.ctx.view(self.B, self.ctx_len, -1).copy_(context)

    def forward(self):
        self.runlist.run()
        return self.eps[..., :self.cfg['out_channels']].permute(0, 3, 1, 2)


class _OffsetPtr:
    """rows [row0, ...) of a 2-D View (as a pointer holder for GemmPlan)"""

    def __init__(self, view, row0):
        self.view, self.row0 = view, row0
        self.device = view.t.device

    def data_ptr(self):
        return self.view.ptr + 2 * self.row0 * self.view.ld


class _OffsetRaw:
    def __init__(self, t, elem_off):
        self.t, self.off = t, elem_off
        self.device = t.device

    def data_ptr(self):
        return self.t.data_ptr() + 2 * self.off


# =============================================================================================== VAE encoder
VAE_SD15 = dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4)


def _flip_conv_weight(w):
    """weights of the data-gradient of a stride-1 3x3 convolution: W'[cin, cout, ky, kx] = W[cout, cin, 2-ky, 2-kx]"""
    return w.flip(2, 3).permute(1, 0, 2, 3).contiguous()


class VaeEncoderEngine:
    """VAE encoder forward (image -> moments) and its data-gradient (d moments -> d image); weights frozen.

    inputs : self.img [B, R, R, 8] fp16 (channels 0..2 = 2*rgb-1)      outputs: self.moments [B, R/8, R/8, 8]
    backward: self.d_moments [B, R/8, R/8, 8] -> self.d_img [B, R, R, 8]
    """

    def __init__(self, sd, device, cfg=VAE_SD15, batch=1, res=512):
        self.cfg, self.B, self.res, self.dev, self.sd = cfg, batch, res, device, sd
        fb = self.fb = Builder(device)
        bb = self.bb = Builder(device)
        self._bwd = []          # closures that append backward ops, run in reverse
        B, ch = batch, cfg['ch']
        self.img = fb.buf(B, res, res, 8, zero=True)
        x = View(fb.buf(B, res, res, ch))
        self._w_in = _f32(sd['conv_in.weight'], device).contiguous() if DIRECT_CONV_IN and cfg['in_channels'] <= 4 and ch % 128 == 0 else None
        if self._w_in is not None:
            fb.conv3x3_small_cin('conv_in', View(self.img), cfg['in_channels'], self._w_in, self._f32('conv_in.bias'), x)
        else:
            fb.gemm('conv_in', View(self.img), cfg['in_channels'], pack_conv_weight(sd['conv_in.weight'].to(device)), ch, x, taps=9, bias=self._f32('conv_in.bias'))
        x0 = x
        in_mult = (1,) + tuple(cfg['ch_mult'])
        nlev = len(cfg['ch_mult'])
        chain = []              # (kind, prefix, x_in View, x_out View, extras)
        for i in range(nlev):
            bin_, bout = ch * in_mult[i], ch * cfg['ch_mult'][i]
            for j in range(cfg['num_res_blocks']):
                x = self._res(f'down.{i}.block.{j}', x, bout)
                bin_ = bout
            if i != nlev - 1:
                x = self._down(f'down.{i}.downsample', x)
        x = self._res('mid.block_1', x, x.C)
        x = self._attn('mid.attn_1', x)
        x = self._res('mid.block_2', x, x.C)
        # norm_out + swish + conv_out + quant_conv
        tn = View(fb.buf(B, x.H, x.W, x.C))
        g, be = self._f32('norm_out.weight'), self._f32('norm_out.bias')
        st = fb.groupnorm('norm_out', x, tn, g, be, 1e-6, True)
        zc = 2 * cfg['z_channels']
        co = View(fb.buf(B, x.H, x.W, 8, zero=True))
        fb.gemm('conv_out', tn, x.C, pack_conv_weight(sd['conv_out.weight'].to(device)), zc, co, taps=9, bias=self._f32('conv_out.bias'))
        self.moments = fb.buf(B, x.H, x.W, 8, zero=True)
        wq = sd['quant_conv.weight'].reshape(zc, zc)
        fb.gemm('quant_conv', co, zc, _pack_linear(wq, device), zc, View(self.moments), bias=self._f32('quant_conv.bias'))
        self.fwd = RunList(fb.ops)
        self.gn_fused = fb.gn_fused

        # ---------------- backward list
        self.d_moments = bb.buf(B, x.H, x.W, 8, zero=True)
        dco = View(bb.buf(B, x.H, x.W, 8, zero=True))
        bb.gemm('quant_conv^T', View(self.d_moments), zc, _pack_linear(wq.t().contiguous(), device), zc, dco)
        dtn = View(bb.buf(B, x.H, x.W, x.C))
        bb.gemm('conv_out^T', dco, zc, pack_conv_weight(_flip_conv_weight(sd['conv_out.weight'].to(device))), x.C, dtn, taps=9)
        dx = View(bb.buf(B, x.H, x.W, x.C))
        bb.groupnorm_bwd('norm_out^T', x, dtn, dx, g, be, 1e-6, True, st, 0)
        for fn in reversed(self._bwd):
            dx = fn(dx)
        self.d_img = bb.buf(B, res, res, 8, zero=True)
        if self._w_in is not None:
            bb.conv3x3_small_cin_dgrad('conv_in^T', dx, self._w_in, cfg['in_channels'], View(self.d_img))
        else:
            bb.gemm('conv_in^T', dx, ch, pack_conv_weight(_flip_conv_weight(sd['conv_in.weight'].to(device))), cfg['in_channels'], View(self.d_img), taps=9)
        self.bwd = RunList(bb.ops)
        self.flops_fwd, self.flops_bwd = fb.flops, bb.flops
        self.sd = None

    def _f32(self, key):
        return _f32(self.sd[key], self.dev)

    def _w1x1(self, key, transpose=False):
        w = self.sd[key].reshape(self.sd[key].shape[0], -1)
        return _pack_linear(w.t().contiguous() if transpose else w, self.dev)

    def _res(self, p, x, cout):
        fb, bb, sd, dev, B = self.fb, self.bb, self.sd, self.dev, self.B
        cin = x.C
        g1, b1 = self._f32(p + '.norm1.weight'), self._f32(p + '.norm1.bias')
        g2, b2 = self._f32(p + '.norm2.weight'), self._f32(p + '.norm2.bias')
        t1 = View(fb.buf(B, x.H, x.W, cin))
        st1 = fb.groupnorm(p + '.norm1', x, t1, g1, b1, 1e-6, True)
        h1 = View(fb.buf(B, x.H, x.W, cout))
        fb.gemm(p + '.conv1', t1, cin, pack_conv_weight(sd[p + '.conv1.weight'].to(dev)), cout, h1, taps=9, bias=self._f32(p + '.conv1.bias'))
        t2 = t1 if cin == cout else View(fb.buf(B, x.H, x.W, cout))       # t1 is dead once conv1 has run
        st2 = fb.groupnorm(p + '.norm2', h1, t2, g2, b2, 1e-6, True)
        if cin != cout:
            xs = View(fb.buf(B, x.H, x.W, cout))
            fb.gemm(p + '.nin_shortcut', x, cin, self._w1x1(p + '.nin_shortcut.weight'), cout, xs, bias=self._f32(p + '.nin_shortcut.bias'))
        else:
            xs = x
        out = View(fb.buf(B, x.H, x.W, cout))
        fb.gemm(p + '.conv2', t2, cout, pack_conv_weight(sd[p + '.conv2.weight'].to(dev)), cout, out, taps=9, bias=self._f32(p + '.conv2.bias'), residual=xs)
        w2t = pack_conv_weight(_flip_conv_weight(sd[p + '.conv2.weight'].to(dev)))
        w1t = pack_conv_weight(_flip_conv_weight(sd[p + '.conv1.weight'].to(dev)))
        wst = self._w1x1(p + '.nin_shortcut.weight', transpose=True) if cin != cout else None

        def backward(dout):
            dt2 = View(bb.buf(B, x.H, x.W, cout))
            bb.gemm(p + '.conv2^T', dout, cout, w2t, cout, dt2, taps=9)
            dh1 = View(bb.buf(B, x.H, x.W, cout))
            bb.groupnorm_bwd(p + '.norm2^T', h1, dt2, dh1, g2, b2, 1e-6, True, st2, 0)
            dt1 = View(bb.buf(B, x.H, x.W, cin))
            bb.gemm(p + '.conv1^T', dh1, cout, w1t, cin, dt1, taps=9)
            if cin != cout:
                dxv = View(bb.buf(B, x.H, x.W, cin))
                bb.gemm(p + '.nin_shortcut^T', dout, cout, wst, cin, dxv)
            else:
                dxv = dout
            bb.groupnorm_bwd(p + '.norm1^T', x, dt1, dxv, g1, b1, 1e-6, True, st1, 1)
            return dxv

        self._bwd.append(backward)
        return out

    def _down(self, p, x):
        fb, bb, dev, B = self.fb, self.bb, self.dev, self.B
        C = x.C
        Ho = x.H // 2
        w4 = self.sd[p + '.conv.weight'].to(dev)
        w2 = w4.permute(0, 2, 3, 1).reshape(C, 9 * C)              # [Cout, tap*C + c]
        out = View(fb.buf(B, Ho, Ho, C))
        if STRIDED_TMA_CONV:
            # (0,1,0,1) zero pad then stride 2 (model.py:67-79): input pixel = 2 * output pixel + tap, the TMA zero-fills the row / column
            # past the border; no im2col buffer (302 MB at 512x512x128)
            fb.gemm(p + '.conv', x, C, pack_conv_weight(w4), C, out, taps=9, bias=self._f32(p + '.conv.bias'), geom=(B, Ho, Ho),
                    a_strides=x.strides(), stride=2, pad_lo=0)
        else:
            col = torch.empty(B, Ho, Ho, 9 * C, device=dev, dtype=torch.float16)
            fb.im2col_s2(p + '.im2col', x, col, 0, 0)
            fb.gemm(p + '.conv', View(col), 9 * C, _pack_linear(w2, dev), C, out, bias=self._f32(p + '.conv.bias'))
        wt = w2.t().contiguous()                                   # [9C, C] : d col = d out . W

        def backward(dout):
            dcol = torch.empty(B, Ho, Ho, 9 * C, device=dev, dtype=torch.float16)
            bb.gemm(p + '.conv^T', dout, C, _pack_linear(wt, dev), 9 * C, View(dcol))
            dxv = View(bb.buf(B, x.H, x.W, C))
            bb.col2im_s2(p + '.col2im', dcol, dxv, 0, 0)
            return dxv

        self._bwd.append(backward)
        return out

    def _attn(self, p, x):
        """single-head self-attention over H*W tokens (model.py:150-204); batch handled per image"""
        fb, bb, dev, B = self.fb, self.bb, self.dev, self.B
        C, n = x.C, x.H * x.W
        scale = float(C) ** -0.5
        g, be = self._f32(p + '.norm.weight'), self._f32(p + '.norm.bias')
        hn = View(fb.buf(B, x.H, x.W, C))
        st = fb.groupnorm(p + '.norm', x, hn, g, be, 1e-6, False)
        q, k, v = (View(fb.buf(B, x.H, x.W, C)) for _ in range(3))
        for nm, dst in (('q', q), ('k', k), ('v', v)):
            fb.gemm(f'{p}.{nm}', hn, C, self._w1x1(f'{p}.{nm}.weight'), C, dst, bias=self._f32(f'{p}.{nm}.bias'))
        S = torch.empty(B, 1, n, n, device=dev, dtype=torch.float16)
        geom = (B, 1, n)
        tokstr = (C, n * C, n * C)
        sstr = (n, n * n, n * n)
        fb.gemm(p + '.qk', q, C, k, n, S, geom=geom, a_strides=tokstr, o_strides=sstr, w_strides=(C, 0, n * C), w_k_valid=C, n_rows_w=n, cin_iter=C,
                alpha=scale, block_n=128)
        fb.softmax(p + '.softmax', S, B * n, n, n)
        vt = torch.empty(B, C, n, device=dev, dtype=torch.float16)
        fb.transpose(p + '.v^T', v.t, C, vt, n, B, n, C)
        o = View(fb.buf(B, x.H, x.W, C))
        fb.gemm(p + '.pv', S, n, vt, C, o, geom=geom, a_strides=sstr, o_strides=tokstr, w_strides=(n, 0, C * n), w_k_valid=n, n_rows_w=C, cin_iter=n, block_n=128)
        out = View(fb.buf(B, x.H, x.W, C))
        fb.gemm(p + '.proj_out', o, C, self._w1x1(p + '.proj_out.weight'), C, out, bias=self._f32(p + '.proj_out.bias'), residual=x)
        wp_t = self._w1x1(p + '.proj_out.weight', transpose=True)
        wq_t, wk_t, wv_t = (self._w1x1(f'{p}.{nm}.weight', transpose=True) for nm in 'qkv')

        def backward(dout):
            do = View(bb.buf(B, x.H, x.W, C))
            bb.gemm(p + '.proj_out^T', dout, C, wp_t, C, do)
            dP = torch.empty(B, 1, n, n, device=dev, dtype=torch.float16)
            bb.gemm(p + '.dP', do, C, v, n, dP, geom=geom, a_strides=tokstr, o_strides=sstr, w_strides=(C, 0, n * C), w_k_valid=C, n_rows_w=n, cin_iter=C, block_n=128)
            dS = torch.empty(B, 1, n, n, device=dev, dtype=torch.float16)
            bb.softmax_bwd(p + '.softmax^T', S, dP, dS, B * n, n, n, scale)
            kt = torch.empty(B, C, n, device=dev, dtype=torch.float16)
            qt = torch.empty(B, C, n, device=dev, dtype=torch.float16)
            dot = torch.empty(B, C, n, device=dev, dtype=torch.float16)
            bb.transpose(p + '.k^T', k.t, C, kt, n, B, n, C)
            bb.transpose(p + '.q^T', q.t, C, qt, n, B, n, C)
            bb.transpose(p + '.dO^T', do.t, C, dot, n, B, n, C)
            dq, dk, dv = (View(bb.buf(B, x.H, x.W, C)) for _ in range(3))
            wT = (n, 0, C * n)
            bb.gemm(p + '.dQ', dS, n, kt, C, dq, geom=geom, a_strides=sstr, o_strides=tokstr, w_strides=wT, w_k_valid=n, n_rows_w=C, cin_iter=n, block_n=128)
            dSt = dP                                                # dP is dead: reuse it for dS^T, then for P^T
            bb.transpose(p + '.dS^T', dS, n, dSt, n, B, n, n)
            bb.gemm(p + '.dK', dSt, n, qt, C, dk, geom=geom, a_strides=sstr, o_strides=tokstr, w_strides=wT, w_k_valid=n, n_rows_w=C, cin_iter=n, block_n=128)
            Pt = dS                                                 # dS is dead now
            bb.transpose(p + '.P^T', S, n, Pt, n, B, n, n)
            bb.gemm(p + '.dV', Pt, n, dot, C, dv, geom=geom, a_strides=sstr, o_strides=tokstr, w_strides=wT, w_k_valid=n, n_rows_w=C, cin_iter=n, block_n=128)
            dhn = View(bb.buf(B, x.H, x.W, C))
            bb.gemm(p + '.q^T.w', dq, C, wq_t, C, dhn)
            bb.gemm(p + '.k^T.w', dk, C, wk_t, C, dhn, residual=dhn)
            bb.gemm(p + '.v^T.w', dv, C, wv_t, C, dhn, residual=dhn)
            bb.groupnorm_bwd(p + '.norm^T', x, dhn, dout, g, be, 1e-6, False, st, 1)
            return dout

        self._bwd.append(backward)
        return out


# =============================================================================================== SDS step
def alphas_cumprod(n=1000, linear_start=0.00085, linear_end=0.012):
    """'scaled_linear' betas (sqrt-linear), as DDIMScheduler for SD / ldm make_beta_schedule('linear', ...) (util.py:21-25)"""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0).float()


VAE_SCALING = 0.18215


class SDSEngine:
    """pred_rgb -> (bilinear 512) -> VAE encode -> latents -> add noise -> UNet x2 (CFG) -> SDS gradient -> VAE data-gradient
    -> d pred_rgb, as three launch lists (encode, unet, backward) around two tiny glue kernels."""

    def __init__(self, unet_sd, vae_sd, device, unet_cfg=UNET_SD15, vae_cfg=VAE_SD15, n_views=1, render_hw=64, ctx_len=77, vae_res=512,
                 capture=False):
        self.dev, self.nv, self.rhw, self.vae_res = device, n_views, render_hw, vae_res
        self.lat_hw = vae_res // 8
        self.unet = UNetEngine(unet_sd, device, unet_cfg, batch=2 * n_views, hw=self.lat_hw, ctx_len=ctx_len)
        self.vae = VaeEncoderEngine(vae_sd, device, vae_cfg, batch=n_views, res=vae_res) if vae_sd is not None else None
        self.acp = alphas_cumprod().to(device)
        B, hw = n_views, self.lat_hw
        self.t = torch.zeros(B, dtype=torch.int32, device=device)
        self.noise = torch.zeros(B, 4, hw, hw, device=device)
        self.eps_post = torch.zeros(B, 4, hw, hw, device=device)
        self.latents = torch.zeros(B, 4, hw, hw, device=device)
        self.latents_in = torch.zeros(B, 4, hw, hw, device=device)
        self.grad = torch.zeros(B, 4, hw, hw, device=device)
        self.loss = torch.zeros(1, device=device)
        self.pred_rgb = torch.zeros(B, 3, render_hw, render_hw, device=device)
        self.d_pred_rgb = torch.zeros(B, 3, render_hw, render_hw, device=device)
        self.guidance_scale, self.grad_scale = 100.0, 1.0
        if capture:
            self.unet.runlist.capture()
            if self.vae is not None:
                self.vae.fwd.capture()
                self.vae.bwd.capture()

    def set_text(self, text_embeddings):
        """[2*n_views, ctx_len, ctx_dim]: unconditional rows first, as sd_utils.train_step concatenates them"""
        self.unet.ctx.view(2 * self.nv, self.unet.ctx_len, -1).copy_(text_embeddings)

    def step(self, as_latent=False):
        """Consumes self.pred_rgb / self.latents_in, self.t, self.noise, self.eps_post; fills self.loss, self.grad, self.d_pred_rgb."""
        self.encode(as_latent)
        self.unet.runlist.run()
        self.finish(as_latent)

    def encode(self, as_latent=False):
        """pred_rgb -> (resize, VAE encoder, posterior sample | latent resize) -> latents -> add_noise into both CFG halves of the UNet input"""
        st = _lib.stream()
        B, hw = self.nv, self.lat_hw
        u = self.unet
        u.t_in[:B].copy_(self.t)
        u.t_in[B:].copy_(self.t)
        if as_latent:
            _lib.call('sdf_sds_prepare', None, 0, _lib.ptr(self.latents_in), None, _lib.ptr(self.noise), _lib.ptr(self.t), _lib.ptr(self.acp), B, hw * hw,
                      _lib.ptr(self.latents), _lib.ptr(u.x_in), 8, VAE_SCALING, st)
        else:
            v = self.vae
            _lib.call('sdf_bilinear_forward', _lib.ptr(self.pred_rgb), B, 3, self.rhw, self.rhw, _lib.ptr(v.img), 8, self.vae_res, self.vae_res, 2.0, -1.0, st)
            v.fwd.run()
            _lib.call('sdf_sds_prepare', _lib.ptr(v.moments), 8, None, _lib.ptr(self.eps_post), _lib.ptr(self.noise), _lib.ptr(self.t), _lib.ptr(self.acp), B,
                      hw * hw, _lib.ptr(self.latents), _lib.ptr(u.x_in), 8, VAE_SCALING, st)

    def finish(self, as_latent=False, view_scale=None):
        """CFG + w(t)(eps_hat - eps) + loss on self.unet.eps, then the VAE data-gradient and the resize adjoint; view_scale: optional
        device float [n_views] multiplying the gradient per image"""
        st = _lib.stream()
        B, hw = self.nv, self.lat_hw
        u = self.unet
        if as_latent:
            _lib.call('sdf_sds_grad', _lib.ptr(u.eps), 8, _lib.ptr(self.noise), _lib.ptr(self.t), _lib.ptr(self.acp), B, hw * hw, float(self.guidance_scale),
                      float(self.grad_scale), _lib.ptr(view_scale), None, 0, None, VAE_SCALING, _lib.ptr(self.grad), None, _lib.ptr(self.loss), st)
        else:
            v = self.vae
            _lib.call('sdf_sds_grad', _lib.ptr(u.eps), 8, _lib.ptr(self.noise), _lib.ptr(self.t), _lib.ptr(self.acp), B, hw * hw, float(self.guidance_scale),
                      float(self.grad_scale), _lib.ptr(view_scale), _lib.ptr(v.moments), 8, _lib.ptr(self.eps_post), VAE_SCALING, _lib.ptr(self.grad),
                      _lib.ptr(v.d_moments), _lib.ptr(self.loss), st)
            v.bwd.run()
            _lib.call('sdf_bilinear_backward', _lib.ptr(v.d_img), 8, self.vae_res, self.vae_res, _lib.ptr(self.d_pred_rgb), B, 3, self.rhw, self.rhw, 2.0, st)


def random_state(shapes, device, seed=0, dtype=torch.float16):
    """Random-init weights of a given {name: shape} table (fan-in scaled uniform, like nn.Conv2d / nn.Linear defaults;
    norm weights 1, norm/other biases small).  Used by bench.py: synthetic weights of the SD-1.5 architecture."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if len(shp) >= 2:
            fan_in = 1
            for s_ in shp[1:]:
                fan_in *= s_
            bound = 1.0 / math.sqrt(fan_in)
            out[k] = ((torch.rand(shp, generator=g, device=device) * 2 - 1) * bound).to(dtype)
        elif 'norm' in k and k.endswith('weight') or k.endswith('in_layers.0.weight') or k.endswith('out_layers.0.weight') or k == 'out.0.weight':
            out[k] = torch.ones(shp, device=device, dtype=dtype)
        else:
            out[k] = ((torch.rand(shp, generator=g, device=device) * 2 - 1) * 0.05).to(dtype)
    return out
