#!/usr/bin/env python
"""SDS steps/s of the -O preset (64x64 render, SD-1.5-shaped UNet guidance) on N B200s — BASELINE.json's metric on config C2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3]   ours (sm_100a kernels through the C ABI)
  python bench.py --impl reference ...                           the reference's CPU path (-O2 vanilla NeRF + PyTorch UNet/VAE on the
                                                                 host cores; oracle port, see oracle/nerf_o2.py, oracle/sd_ref.py)
  python bench.py --impl reference-cuda ...                      the reference's own -O training loop on this GPU: its unmodified Trainer /
                                                                 renderer / network Python on its own CUDA extensions + PyTorch fp16 SD
                                                                 (oracle/ref_harness.py); at N = 1 the main line runs it too and reports
                                                                 vs_baseline = ours / that (the north star's target: >= 1.5)
Under torchrun (N > 1) every rank renders its own view and the NeRF gradients are all-reduced once per step (weak scaling:
value = views processed by all ranks / s).  One JSON line is printed by rank 0.

A "step" = nerf/utils.py:1032-1072 loop body: occupancy-grid refresh every 16 steps, render, SDS loss (VAE encode + 2x UNet),
backward into the hash table / MLPs, Adan step.  Shading follows the reference schedule mix (20 % latent/normal, 64 %
lambertian, 16 % textureless) in a fixed 25-step cycle.  Synthetic data: random orbit cameras, default-initialised NeRF,
seeded random weights of the SD-1.5 architecture (no network access: no checkpoints).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "stable-dreamfusion_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "SDS steps/sec (64x64 render, SD-1.5 UNet)"
CONFIGS = {"C2": dict(hw=64, note="BASELINE.json configs[1]: -O backbone, 64x64 render, 1 view/step/GPU"),
           "C3": dict(hw=128, note="BASELINE.json configs[2]: -O backbone, 128x128 render, 1 view/step/GPU (4 views over 4 GPUs)"),
           "C4": dict(hw=64, zero123=True, note="BASELINE.json configs[3]: Zero-1-to-3-shaped UNet guidance (8-channel input, 1-token context, 32x32 "
                                               "latents, VAE 256x256), 64x64 render, 1 view/step/GPU"),
           "C5": dict(hw=512, dmtet=True, note="BASELINE.json configs[4]: DMTet fine-tuning stage (tet lattice of the tets/128 size class), 512x512 rasterised "
                                               "render + SD-1.5 SDS, 1 view/step/GPU")}
def _schedule_cycle():
    """The reference's shading schedule (20 % latent warm-up, then 80 % lambertian / 20 % textureless: main.py:150-153, nerf/utils.py:503-535) as a
    25-step cycle with the modes INTERLEAVED — every 5 consecutive steps hold one latent step, every 25 hold 5 / 16 / 4 — so that a timed window of
    any length K (the driver picks K) sees the schedule's mix to within one step, and the resident and end-to-end windows see the same one."""
    out, shaded = [], 0
    for i in range(25):
        if i % 5 == 0:
            out.append("latent")
        else:
            out.append("textureless" if shaded % 5 == 4 else "lambertian")
            shaded += 1
    return out


CYCLE = _schedule_cycle()        # 20 % / 64 % / 16 %


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the two field kernels, from the ncu --set full capture summarised in
# profiles/ (the fp16 table is L2-resident: DRAM traffic is ~11 % of the 4.8 GB of algorithmic bytes, most of it the feature stash)
FIELD_DRAM_TRAFFIC = {"fwd_dram_bytes": 211.5e6, "bwd_dram_bytes": 287.8e6, "at_samples": 432365,
                      "source": "profiles/r02_kernels.md addendum (ncu --set full on the marched default-view sample set; 178 MB of the forward's writes and 194 MB "
                                "of the backward's reads are the feature stash)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:6]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)
def host_core_limit():
    """cores this process may really use: affinity mask capped by the cgroup CPU quota (containers report the machine's count)"""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, q // per))
        except Exception:
            pass
    return max(1, n)


def pick_host_threads(torch):
    """torchrun exports OMP_NUM_THREADS=1 and containers over-report cores: time a small conv at a few thread counts and keep
    the fastest, so the CPU arm really uses all the host threads it can."""
    limit = host_core_limit()
    cands = sorted({c for c in (4, 8, 16, 32, 64, limit) if c <= limit} | {limit})
    x = torch.randn(1, 128, 128, 128)
    w = torch.randn(128, 128, 3, 3)
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        torch.nn.functional.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if dt < best_t * 0.95:
            best, best_t = c, dt
    return best


def run_reference(args):
    """The reference's own CPU implementation of the path: -O2 pure-PyTorch NeRF + PyTorch UNet/VAE on the host cores
    (BASELINE.json config C1: 32x32 render).  Bounded sample: each 'step' is one full CPU SDS step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import nerf_o2, sd_ref
    from sdf_b200 import synth
    cores = pick_host_threads(torch)
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = nerf_o2.VanillaNeRF()
    unet = sd_ref.UNet(**sd_ref.UNET_SD15).eval()
    sd_ref.reinit_zero_modules(unet)
    vae = sd_ref.VaeEncoder(**sd_ref.VAE_SD15).eval()
    for p in list(unet.parameters()) + list(vae.parameters()):
        p.requires_grad_(False)
    from oracle.ref_cuda_path import TorchAdan                  # the reference's optimizer (main.py:368: Adan, foreach=False), restated in PyTorch
    opt_ = TorchAdan([{'params': list(model.parameters()), 'lr': 5e-3}], eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0)
    acp = sd_ref.alphas_cumprod()
    text = torch.randn(2, 77, 768)
    rng = __import__("numpy").random.default_rng(0)

    def step(i):
        pose, _ = synth.rand_pose(rng)
        ro, rd = synth.get_rays(pose, 32, 32, float(rng.uniform(10, 30)))
        shading = CYCLE[i % len(CYCLE)]
        as_latent = shading == "latent"
        out = model.render(torch.from_numpy(ro), torch.from_numpy(rd), 0.55, "normal" if as_latent else shading)
        if as_latent:
            pred = torch.cat([out["image"], out["weights_sum"].unsqueeze(-1)], -1).reshape(1, 32, 32, 4).permute(0, 3, 1, 2)
        else:
            pred = out["image"].reshape(1, 32, 32, 3).permute(0, 3, 1, 2)
        t = torch.randint(20, 981, (1,))
        loss, _, _ = sd_ref.sds_train_step(unet, vae, acp, text, pred, t, torch.randn(1, 4, 64, 64), torch.randn(1, 4, 64, 64), 100.0, as_latent)
        if "loss_orient" in out:
            loss = loss + 1e-2 * out["loss_orient"]
        opt_.zero_grad()
        loss.backward()
        opt_.step()
        return float(loss.detach())

    # bounded sample: the CPU path takes ~10-20 s per step; measure as many of the K requested steps as fit the time budget
    budget_s = float(os.environ.get("SDF_CPU_BASELINE_BUDGET_S", "150"))
    t0 = time.perf_counter()
    step(1)                                   # warm-up (thread pools, allocator) on a shaded step — the costly kind — also calibrates the sample size
    est = time.perf_counter() - t0
    measured = max(1, min(args.steps, int(budget_s / max(est, 1e-3))))
    t0 = time.perf_counter()
    for i in range(measured):
        step(i)                               # the interleaved cycle from its start: latent, 4 x shaded, latent, ...
    dt = time.perf_counter() - t0
    v = measured / dt
    line = {"metric": METRIC, "value": v, "unit": "steps/s", "impl": "reference", "n_gpus": args.gpus, "steps": measured, "steps_requested": args.steps,
            "warmup": 1,
            "ms_per_step": 1e3 * dt / measured, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "reference -O2 CPU path: vanilla NeRF 32x32 (64+32 samples/ray) + SD-1.5-shaped UNet/VAE in PyTorch fp32, 1 view/step"},
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": "port",
                             "sample": f"{measured} full CPU SDS steps (of {args.steps} requested) after 1 warm-up step, config C1: 32x32 render, schedule-mix shading (interleaved cycle: 1 latent step in every 5)"},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ reference, CUDA path (SURVEY §8d arm A)
def time_reference_cuda(hw=64, steps=30, warmup=10):
    """The reference's own training loop on this GPU, unmodified: nerf/utils.py Trainer.train_one_epoch (GradScaler, reference Adan, tqdm,
    loss.item() per step) + nerf/renderer.py + nerf/network_grid.py on the reference's CUDA extensions, SD-1.5-shaped UNet/VAE in PyTorch
    fp16 (cuDNN / cuBLAS / SDPA), in a fresh interpreter (oracle/ref_harness.py `time`).  Two windows, one per phase of the reference
    schedule (latent: the first 20 % of the iterations; shaded: lambertian / textureless mix), each `warmup` untimed + `steps` timed steps with
    cudnn.benchmark OFF (with it on, the first ~30 steps are autotuning: round 1 measured 1.35 steps/s over 13 such steps against 3.19
    once settled); combined with the 20/80 weights of the schedule the main line cycles through."""
    import tempfile
    from oracle import ref_harness as RH
    tmp = tempfile.mkdtemp(prefix="sdf_refcuda_")
    res = {}
    specs = []
    for phase, gstep in (("latent", 0), ("shaded", 2096)):
        specs.append(dict(cmd="time", workspace=os.path.join(tmp, "ws_" + phase), steps=steps, warmup=warmup, global_step=gstep, seed=0,
                          opt=dict(h=hw, w=hw), out=os.path.join(tmp, phase + ".json")))
    RH.run_subprocess(dict(cmd="multi", ops="reference", specs=specs), timeout=3000)
    for phase in ("latent", "shaded"):
        res[phase] = json.load(open(os.path.join(tmp, phase + ".json")))
    ms = 0.2 * res["latent"]["ms_per_step"] + 0.8 * res["shaded"]["ms_per_step"]
    return {"value": 1e3 / ms, "unit": "steps/s", "ms_per_step": ms, "ms_per_step_latent": res["latent"]["ms_per_step"],
            "ms_per_step_shaded": res["shaded"]["ms_per_step"], "steps": 2 * steps, "warmup": 2 * warmup,
            "what": "reference Trainer.train_one_epoch, unmodified, on its own CUDA extensions + PyTorch fp16 UNet/VAE; 0.2 x latent-phase + 0.8 x "
                    "shaded-phase step time; cudnn.benchmark off"}


def run_reference_cuda(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    hw = CONFIGS[args.config]["hw"]
    r = time_reference_cuda(hw, steps=max(args.steps // 2, 10), warmup=max(args.warmup, 10))
    print(json.dumps({"metric": METRIC, "value": r["value"], "unit": "steps/s", "impl": "reference-cuda", "n_gpus": 1, "steps": r["steps"], "warmup": r["warmup"],
                      "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                      "config": {"workload": f"reference -O path on this GPU ({hw}x{hw}, 1 view/step): " + r["what"]}, "detail": r}), flush=True)


# ------------------------------------------------------------------------------------------------ multi-GPU equivalence
def ray_parallel_selfcheck(opt, dev, rank, world):
    """N GPUs x 1 view == 1 GPU x N views.  One ray-parallel step (every rank renders pixels rank::N of every view, all-to-all of
    pixels / pixel gradients, ONE all-reduce of the flat gradient bucket) against a single-process render of the same N whole views on this
    rank, same cameras / lights / weights, march jitter off, with a linear per-view loss <pred_rgb, G_view> so that the comparison isolates the
    distributed plumbing.  Returns the relative L2 distance of the summed parameter gradients (every rank computes it)."""
    import copy
    import torch
    from sdf_b200.render import render_train
    from sdf_b200.trainer import SDSTrainer
    o = copy.copy(opt)
    o.lambda_entropy, o.lambda_orient, o.lambda_opacity = 0.0, 0.0, 0.0      # their means run over per-rank sample sets
    H, W, Bv = o.h, o.w, o.batch_size
    gen = torch.Generator(device="cpu").manual_seed(4242)
    G_all = torch.randn(world * Bv, 3, H, W, generator=gen).to(dev)

    class LinearGuidance:
        def get_text_embeds(self, prompt):
            return torch.zeros(len(prompt), 1, 1, device=dev)

        def train_step(self, text_z, pred_rgb, as_latent=False, **kw):
            return (pred_rgb * G_all[rank * Bv:(rank + 1) * Bv, :pred_rgb.shape[1]]).sum()

    tr = SDSTrainer(o, dev, LinearGuidance(), seed=123, rank=rank, world_size=world, ema_decay=None)
    tr.perturb = False
    tr.optimizer.step = lambda **k: None                                      # keep the reduced gradients for inspection
    tr.train_step(shading="lambertian")
    dist_grad = tr.bucket.flat.clone()
    tr.bucket.flat.zero_()
    L = tr._last_shared
    rays_o, rays_d = tr._rays(L["poses"], L["fov"])
    light = L["light"].repeat_interleave(H * W, dim=0)
    out = render_train(tr.model, rays_o, rays_d, light_d=light, ambient_ratio=L["ambient"], shading=L["mode"], bg_color=L["bg_color"], perturb=False,
                       as_latent=False, B=world * Bv, H=H, W=W, direct_grads=True)
    (out["pred_rgb"] * G_all).sum().backward()
    single = tr.bucket.flat
    rel = float(((dist_grad - single).norm() / (single.norm() + 1e-30)).item())
    cos = float((dist_grad @ single / (dist_grad.norm() * single.norm() + 1e-30)).item())
    return {"table_and_mlp_grad_rel_l2": rel, "cosine": cos, "grad_norm": float(single.norm().item()),
            "what": f"{world}-rank ray-parallel step (all-to-all pixels + one bucket all-reduce) vs one rank rendering the same {world * Bv} whole views"}


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args):
    import numpy as np
    import torch
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from sdf_b200 import _lib
    from sdf_b200.options import default_opt
    from sdf_b200.trainer import SDSTrainer
    from guidance.sd_utils import StableDiffusion

    hw = CONFIGS[args.config]["hw"]
    opt = default_opt(h=hw, w=hw, batch_size=1)
    is_dmtet = bool(CONFIGS[args.config].get("dmtet"))
    if is_dmtet:
        from sdf_b200.options import dmtet_opt
        opt = dmtet_opt(batch_size=1)                                   # main.py:253-274: 64 x dmtet_reso_scale 8 = 512, no latent / albedo phases
        assert opt.h == hw
        guidance = StableDiffusion(dev, weights="random", n_views=1, render_hw=hw, seed=0, capture=True)
        trainer = SDSTrainer(opt, dev, guidance, seed=0, rank=rank, world_size=world)
        trainer.model.update_extra_state()                              # stands for the coarse stage's trained density: the initial blob
        trainer.model.init_tet()                                        # main.py:317-324 (init_with a checkpoint -> init_tet)
    elif CONFIGS[args.config].get("zero123"):
        from guidance.zero123_utils import Zero123
        opt.guidance_scale, opt.latent_iter_ratio = 5.0, 0.0            # main.py:196-219 defaults of an image-only run
        opt.zero123_grad_scale = "angle"
        guidance = Zero123(dev, opt=opt, weights="random", n_views=1, render_hw=hw, seed=0, capture=True)
        ref_img = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(0)).to(dev)
        c, v = guidance.get_img_embeds(ref_img)
        emb = {"c_crossattn": c, "c_concat": v, "ref_polars": [90.0], "ref_azimuths": [0.0], "ref_radii": [3.2], "zero123_ws": [1.0]}
        trainer = SDSTrainer(opt, dev, guidance, seed=0, rank=rank, world_size=world, image_embeddings=emb)
    else:
        guidance = StableDiffusion(dev, weights="random", n_views=1, render_hw=hw, seed=0, capture=True)
        trainer = SDSTrainer(opt, dev, guidance, seed=0, rank=rank, world_size=world)
    cycle = [c for c in CYCLE if c != "latent"] if (CONFIGS[args.config].get("zero123") or is_dmtet) else CYCLE
    eng = guidance.engine
    launches = {"n": 0}
    orig_call = _lib.call

    def counting_call(name, *a):
        launches["n"] += 1
        return orig_call(name, *a)
    _lib.call = counting_call
    graph_ops = {"unet": len(eng.unet.runlist.ops), "vae_fwd": len(eng.vae.fwd.ops), "vae_bwd": len(eng.vae.bwd.ops)}

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, first_index, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_steps):
            trainer.train_step(shading=cycle[(first_index + i) % len(cycle)], read_loss=read_loss)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(max(3, args.warmup)):
        trainer.train_step(shading=cycle[i % len(cycle)], read_loss=False)
    sampler = ClockSampler(local)
    sampler.start()
    launches["n"] = 0
    ms_res = timed(args.steps, args.warmup, read_loss=False)         # inputs resident: no loss read-back
    n_calls = launches["n"]
    ms_e2e = timed(args.steps, args.warmup + args.steps, read_loss=True)   # pinned-host pose in, loss out, every step
    sampler.stop_flag = True
    sampler.join(timeout=2)

    hbm, tf_burst, tf_sus, peak_kind = peaks()
    # --- roofline of the dominant kernel (tcgen05 implicit GEMM): all GEMM launches of one SDS step, timed alone
    gemm_plans = []
    flops = 0.0
    for rl in (eng.vae.fwd, eng.unet.runlist, eng.vae.bwd):
        for name, fn in rl.ops:
            plan = getattr(fn, "__self__", None)
            if plan is not None and hasattr(plan, "flops") and hasattr(plan, "handle"):
                gemm_plans.append(plan)
                flops += plan.flops
    for p in gemm_plans[:50]:
        p.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        for p in gemm_plans:
            p.run()
    e1.record()
    torch.cuda.synchronize()
    gemm_ms = e0.elapsed_time(e1) / reps
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    roof_gemm = {"bound": "tensor", "kernel": "k_gemm (tcgen05 implicit GEMM: all conv/linear/attention products of one SDS step)", "achieved": achieved,
            "peak": tf_sus, "peak_kind": f"bf16_tflops_sustained ({peak_kind})", "unit": "TFLOP/s", "frac": achieved / tf_sus,
            "launches_per_step": len(gemm_plans), "flops_per_step": flops, "ms_per_step_alone": gemm_ms, "traffic": None}

    # --- fused hashgrid+MLP field kernels (the kernel BASELINE.json's metric names): algorithmic bytes (540 B forward /
    # 1052 B backward per point-eval, 7 point-evals per lambertian sample) / CUDA-event time of the bare C-ABI launches on a typical sample
    # set; L2 is flushed (256 MB write) before every timed launch, so the 24 MB fp16 table starts in HBM as it does inside a step
    fld = {}
    try:
        # the sample set of the default view (radius 3.2, polar 90, azimuth 0, fovy 20: M = 432 k at the step-0 occupancy) marched
        # through the model's current occupancy grid: real samples cluster along rays, which is what the kernels see inside a step
        import raymarching as _rm
        from sdf_b200 import synth as _sy
        g = torch.Generator(device=dev).manual_seed(1)
        _ro, _rd = _sy.get_rays(_sy.circle_pose(3.2, 90.0, 0.0), 64, 64, 20.0)
        _ro, _rd = torch.from_numpy(_ro).to(dev), torch.from_numpy(_rd).to(dev)
        _n, _f = _rm.near_far_from_aabb(_ro, _rd, trainer.model.aabb_train, 0.2)
        xyz = _rm.march_rays_train(_ro, _rd, trainer.model.bound, trainer.model.density_bitfield, trainer.model.cascade, trainer.model.grid_size, _n, _f,
                                   True, opt.dt_gamma, opt.max_steps)[0].contiguous()
        M = int(xyz.shape[0])
        l = torch.nn.functional.normalize(torch.randn(M, 3, device=dev, generator=g), dim=-1).contiguous()
        m = trainer.model
        c = m.field_cfg()
        sn = m.sigma_net.net
        P = _lib.ptr
        table = m.table_half()
        sig, col, nrm, aux = (torch.empty(M, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 3, device=dev), torch.empty(M, 10, device=dev))
        gs, gc = torch.full((M,), 1e-3, device=dev), torch.randn(M, 3, device=dev, generator=g)
        gt = torch.zeros_like(m.encoder.embeddings)
        gw = [torch.zeros_like(t) for t in (sn[0].weight, sn[0].bias, sn[1].weight, sn[1].bias, sn[2].weight, sn[2].bias)]
        wts = [P(t) for t in (sn[0].weight, sn[0].bias, sn[1].weight, sn[1].bias, sn[2].weight, sn[2].bias)]
        fargs = (P(xyz), M, None, P(table), P(c["offsets"]), c["L"], c["levels_active"], c["S"], int(c["H"]), int(c["smoothstep"]), *wts, m.bound,
                 c["blob_density"], c["blob_radius"], 1, P(l), 1, 0.5)
        feat = torch.empty(_lib.query("sdf_field_feat_bytes", M, 1) // 4, device=dev, dtype=torch.int32)      # forward -> backward feature stash (as in a step)
        flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)
        st = _lib.stream()
        tf_, tb_ = [], []
        for rep in range(7):
            flush.fill_(float(rep))
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e[0].record()
            orig_call("sdf_field_forward", *fargs, P(sig), P(col), P(nrm), P(aux), P(feat), st)
            e[1].record()
            flush.fill_(float(rep) + 0.5)
            e[2].record()
            orig_call("sdf_field_backward", *fargs, P(aux), P(gs), P(gc), None, P(gt), *[P(t) for t in gw], P(feat), st)
            e[3].record()
            torch.cuda.synchronize()
            if rep >= 2:
                tf_.append(e[0].elapsed_time(e[1]) * 1e-3)
                tb_.append(e[2].elapsed_time(e[3]) * 1e-3)
        tfw, tbw = sum(tf_) / len(tf_), sum(tb_) / len(tb_)
        bf, bb = M * (7 * 540.0), M * (7 * 1052.0)
        fld = {"bound": "hbm", "kernel": "k_field_forward + k_field_backward (fused hashgrid gather + 32-64-64-4 MLP + 7-point finite-difference normal + "
                                          "shading; lambertian), bare C-ABI launches",
               "achieved": (bf + bb) / (tfw + tbw) / 1e9, "peak": hbm, "peak_kind": f"hbm_gbs ({peak_kind})", "unit": "GB/s",
               "frac": (bf + bb) / (tfw + tbw) / 1e9 / hbm, "samples": M, "sample_set": "default view (r 3.2, polar 90, azimuth 0, fovy 20, 64x64) marched through the current occupancy grid",
               "launches_timed": len(tf_),
               "algorithmic_bytes_per_launch": {"fwd": bf, "bwd": bb, "rule": "SURVEY.md 8d: 540 B fwd / 1052 B bwd per point-eval x 7 point-evals per sample"},
               "fwd": {"ms": tfw * 1e3, "GBps": bf / tfw / 1e9, "frac": bf / tfw / 1e9 / hbm},
               "bwd": {"ms": tbw * 1e3, "GBps": bb / tbw / 1e9, "frac": bb / tbw / 1e9 / hbm},
               "l2": "flushed before every timed launch",
               "traffic": FIELD_DRAM_TRAFFIC["fwd_dram_bytes"] + FIELD_DRAM_TRAFFIC["bwd_dram_bytes"], "traffic_detail": FIELD_DRAM_TRAFFIC}
        del flush, gt, aux
    except Exception as e:      # never let the auxiliary measurement kill the bench line
        fld = {"error": repr(e)}

    # --- secondary figure: inference render rate at 800x800 with the device-side loop (no host sync per iteration,
    #     on-device alive-ray compaction); the reference's only published number is "~10 FPS at 800x800" on a V100 (readme.md:28)
    eval_fps = None
    if world == 1 and not args.no_eval and not is_dmtet:
        try:
            from sdf_b200 import synth as _synth
            m = trainer.model
            m.eval()
            ro, rd = _synth.get_rays(_synth.circle_pose(3.2, 80.0, 30.0), 800, 800, 20.0)
            ro_t, rd_t = torch.from_numpy(ro).to(dev)[None], torch.from_numpy(rd).to(dev)[None]
            light = torch.nn.functional.normalize(torch.tensor([[0.3, 0.8, 0.5]], device=dev), dim=-1)
            eval_fps = {}
            for sh in ("albedo", "lambertian"):
                for rep in range(4):
                    if rep == 1:
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                    out_e = m.render(ro_t, rd_t, None, 800, 800, staged=False, perturb=False, bg_color=None, ambient_ratio=0.5, shading=sh, light_d=light)
                torch.cuda.synchronize()
                eval_fps[sh] = 3.0 / (time.perf_counter() - t0)
            eval_fps["coverage"] = float((out_e["weights_sum"] > 0.5).float().mean())
            eval_fps["what"] = "800x800 frames/s, default view, scene as trained so far (blob); wall clock over 3 frames after 1 warm-up"
            m.train()
        except Exception as e:
            eval_fps = {"error": repr(e)[-300:]}
            trainer.model.train()

    mg_check = None
    if world > 1 and not is_dmtet:
        try:
            mg_check = ray_parallel_selfcheck(opt, dev, rank, world)
        except Exception as e:
            mg_check = {"error": repr(e)[-300:]}

    stages = None
    if args.breakdown:
        stages = {}
        for sh in ("latent", "lambertian", "textureless"):
            acc = {}
            for rep in range(3):
                trainer.stage_events = []
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                trainer.train_step(shading=sh)
                host_ms = (time.perf_counter() - t0) * 1e3
                torch.cuda.synchronize()
                ev = trainer.stage_events
                for (_, a), (name, b_) in zip(ev[:-1], ev[1:]):
                    acc[name] = acc.get(name, 0.0) + a.elapsed_time(b_) / 3
                acc["host issue time"] = acc.get("host issue time", 0.0) + host_ms / 3
                acc["samples"] = trainer.last_M
            stages[sh] = acc
        trainer.stage_events = None

    if rank == 0:
        # --- CPU baseline (oracle port) on the host cores: bounded sample
        cpu = None
        if not args.no_cpu_baseline and world == 1:          # the CPU leg is reported at N = 1 only
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                                     capture_output=True, text=True, timeout=900, env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
                cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception as e:
                cpu = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}"}
        # --- the north star's comparison: the reference's own -O loop on this same GPU (N = 1 only; skipped with --no-ref-cuda)
        refc, vs = None, None
        if world == 1 and not args.no_ref_cuda and args.config not in ("C4", "C5"):        # the reference's Zero123 needs ldm + a checkpoint, its DMTet stage nvdiffrast: no same-GPU arm
            try:
                torch.cuda.empty_cache()
                refc = time_reference_cuda(hw, steps=max(args.steps, 25), warmup=max(args.warmup, 10))
            except Exception as e:
                refc = {"value": None, "error": repr(e)[-400:]}
        mesh_counts = trainer.model.lattice.mesh_counts() if is_dmtet else None
        value = world * args.steps / (ms_res * 1e-3)
        if refc and refc.get("value"):
            vs = value / refc["value"]
        e2e_v = world * args.steps / (ms_e2e * 1e-3)
        per_step_calls = n_calls / args.steps
        # graph replays stand for all captured launches
        frac_latent = cycle.count("latent") / len(cycle)
        launches_per_step = per_step_calls + graph_ops["unet"] + (1 - frac_latent) * (graph_ops["vae_fwd"] + graph_ops["vae_bwd"])
        line = {"metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": vs,
                "vs_baseline_kind": "value / steps-per-second of the reference's own -O CUDA-extension training loop measured on this same GPU in this run "
                                    "(`reference_cuda` below; BASELINE.md publishes no number, its 3.1 defines this arm A)",
                "dtype": "f16", "data": "synthetic",
                "config": {"workload": (f"{args.config}: DMTet fine-tuning stage: marching tetrahedra on a {trainer.model.lattice.N}-vertex / {trainer.model.lattice.F}-tet lattice "
                                        f"(mesh {mesh_counts[0]} vertices / {mesh_counts[1]} faces), {hw}x{hw} rasterised + antialiased render textured by the -O hash-grid "
                                        "network, 1 view/step/GPU, SD-1.5-shaped UNet (B=2 CFG) + VAE encoder 512x512, shading mix 80% lambertian / 20% textureless, mesh "
                                        "normal-consistency + Laplacian regularisers, Adan step over table + MLP + bg net + sdf + deform") if is_dmtet else
                                       f"{args.config}: -O instant-NGP backbone, {hw}x{hw} render, 1 view/step/GPU, " +
                                       ("Zero-1-to-3-shaped UNet (8-ch input, 1-token context, 32x32 latents, B=2 CFG) + VAE encoder 256x256, " if CONFIGS[args.config].get("zero123")
                                        else "SD-1.5-shaped UNet (B=2 CFG) + VAE encoder 512x512, ") +
                                       ("shading mix 80% lambertian / 20% textureless (no latent phase with image guidance), " if CONFIGS[args.config].get("zero123")
                                        else "reference shading schedule mix (20% latent, 64% lambertian, 16% textureless; interleaved so that any timed window sees the mix), ") + "Adan step, grid refresh every 16 steps",
                           "rays_per_view": hw * hw, "samples_last_step": trainer.last_M, "l2": "per-step working set (UNet weights 1.7 GB + activations) exceeds the 126 MB L2",
                           "parallelism": f"dp{world}" + (" (one view per GPU for the UNet/VAE; every GPU renders 1/N of the rays of every view, pixels and pixel "
                                                         "gradients exchanged by all-to-all; one gradient all-reduce)" if world > 1 else "")},
                "e2e": {"value": e2e_v, "unit": "steps/s", "h2d_bytes_per_step": (128 if is_dmtet else 64) * (world if trainer.ray_parallel else 1) + (12 * world if trainer.ray_parallel else 0),
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                        "how": "every step: pose (+ per-view light offset when ray-parallel) copied from pinned host memory, the step's loss copied to a pinned word on a "
                               "side stream as soon as it exists (before the backward) and waited for by the host before the step returns"},
                "gpu_launches": int(round(launches_per_step * args.steps)),
                "roofline": fld, "roofline_gemm": roof_gemm, "reference_cuda": refc, "cpu_baseline": cpu, "clocks": sampler.summary()}
        if stages is not None:
            line["stages"] = stages
        if mg_check is not None:
            line["multi_gpu_check"] = mg_check
        if eval_fps is not None:
            line["eval_render"] = eval_fps
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"],
                    help="reference = the reference's CPU path (driver contract); reference-cuda = arm (A) of SURVEY.md §8d: the reference's own "
                         "unmodified Python + CUDA extensions + PyTorch fp16 UNet/VAE on this GPU (oracle/ref_harness.py)")
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS), help="BASELINE.json config: C2 = 64x64 (the metric's), C3 = 128x128, C4 = Zero123-shaped guidance, C5 = DMTet stage at 512x512")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eval", action="store_true", help="skip the 800x800 inference-render figure")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the same-GPU reference arm (vs_baseline stays null)")
    ap.add_argument("--breakdown", action="store_true", help="add per-stage CUDA-event times of one step per shading mode ('stages')")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-cuda":
        run_reference_cuda(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
