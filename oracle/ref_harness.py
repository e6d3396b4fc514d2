"""Runs the reference's OWN, UNMODIFIED host Python — TEST / BENCH INFRASTRUCTURE ONLY, never imported by the product.

oracle/build_ref.py byte-compiles nerf/{renderer,network_grid,utils,provider}.py, encoding.py, activation.py, optimizer.py and the
four operator wrapper packages of /root/reference into sourceless .pyc trees under oracle/_ref/ (build outputs, git-ignored, they
travel to the GPU box like the reference's compiled .so files).  This module puts them on sys.path in one of two ways:

  ops='reference'  nerf/* + encoding + activation + optimizer   on   the reference's raymarching/gridencoder/freqencoder/shencoder
                   wrappers bound to oracle/_ref/_*.so (its own CUDA kernels), guidance = oracle.sd_ref (PyTorch, cuDNN/cuBLAS).
                   This is the all-reference arm: the parity comparand and bench.py's `--impl reference-cuda`.
  ops='dropin'     the SAME nerf/* + encoding + activation + optimizer   on   stable-dreamfusion_b200/{raymarching,gridencoder,
                   freqencoder,shencoder,guidance}: the drop-in claim of the north star, executed instead of asserted.

It must run in its own process (module names like `raymarching` and `nerf` are global): tests call it through `run_subprocess`.
Packages the path never executes (mesh export, GUI, logging) are stubbed exactly like tests/golden/make_golden_o2.py does;
torch_ema (a third-party dependency of the Trainer, not vendored by the reference, v0.3 algorithm restated below) is provided.

Commands (JSON spec on argv[1]):
  points : NeRFNetwork.forward / .density of the reference class at given points, with parameter gradients (fused-field comparand).
  render : one NeRFNetwork (reference class), fixed weights / occupancy / rays -> image, weights_sum, depth, weights, per-ray counts,
           parameter gradients for a fixed upstream gradient, for each requested (shading, H, pose) case.
  steps  : the reference Trainer (nerf/utils.py) wired as main.py:363-410 wires it, train_one_epoch over n steps; per step the
           inputs the Trainer drew (pose, shading, ambient ratio, background) and pred_rgb, loss, d loss/d pred_rgb; final parameters.
  time   : `steps` without capture, timed with CUDA events after a warm-up (bench.py --impl reference-cuda).
  dmtet  : (ops='dropin' only) the reference's NeRFNetwork with opt.dmtet: its own DMTet class, run_dmtet, normal_consistency and
           laplacian_smooth_loss (nerf/renderer.py:94-254, 862-954) on the drop-in gridencoder AND the drop-in nvdiffrast package; image +
           parameter gradients per case.  The lattice file it loads (tets/<size>_tets.npz) is written from sdf_b200/tetgrid.py.
"""
import json
import os
import sys
import types
import weakref

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFDIR = os.path.join(HERE, "_ref")
PKG = os.path.join(ROOT, "stable-dreamfusion_b200")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        m = _Stub(self.__name__ + "." + k)
        setattr(self, k, m)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        return self


def _torch_ema_module():
    """torch_ema 0.3 (requirements.txt: `torch-ema`, un-vendored): shadow <- shadow - (1 - d)(shadow - p), d = min(decay, (1+n)/(10+n))"""
    import torch
    m = types.ModuleType("torch_ema")

    class ExponentialMovingAverage:
        def __init__(self, parameters, decay, use_num_updates=True):
            parameters = list(parameters)
            self.decay = decay
            self.num_updates = 0 if use_num_updates else None
            self.shadow_params = [p.clone().detach() for p in parameters]
            self.collected_params = None
            self._refs = [weakref.ref(p) for p in parameters]

        def _params(self, parameters):
            return [r() for r in self._refs] if parameters is None else list(parameters)

        @torch.no_grad()
        def update(self, parameters=None):
            decay = self.decay
            if self.num_updates is not None:
                self.num_updates += 1
                decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
            for s, p in zip(self.shadow_params, self._params(parameters)):
                s.sub_((s - p) * (1.0 - decay))

        @torch.no_grad()
        def copy_to(self, parameters=None):
            for s, p in zip(self.shadow_params, self._params(parameters)):
                p.data.copy_(s.data)

        def store(self, parameters=None):
            self.collected_params = [p.clone() for p in self._params(parameters)]

        @torch.no_grad()
        def restore(self, parameters=None):
            for c, p in zip(self.collected_params, self._params(parameters)):
                p.data.copy_(c.data)

        def state_dict(self):
            return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params, "collected_params": self.collected_params}

        def load_state_dict(self, sd):
            self.decay, self.num_updates = sd["decay"], sd["num_updates"]
            self.shadow_params, self.collected_params = sd["shadow_params"], sd["collected_params"]

    m.ExponentialMovingAverage = ExponentialMovingAverage
    return m


def install(ops):
    """arrange sys.path / sys.modules so that `import nerf.network_grid` etc. resolve to the reference's compiled Python"""
    assert ops in ("reference", "dropin")
    for d in ("refpy", "refpy_ops"):
        if not os.path.isdir(os.path.join(REFDIR, d)):
            raise RuntimeError(f"oracle/_ref/{d} missing: run `python oracle/build_ref.py` where /root/reference exists")
    stubs = ["mcubes", "trimesh", "meshutils", "cubvh", "xatlas", "pymeshlab", "tensorboardX", "imageio",
             "torchmetrics", "matplotlib", "matplotlib.pyplot", "lpips", "dearpygui", "dearpygui.dearpygui"]
    if ops == "reference":
        stubs += ["nvdiffrast", "nvdiffrast.torch"]          # not vendored by the reference and not installed; the drop-in arm binds the repository's package
    for name in stubs:
        sys.modules.setdefault(name, _Stub(name))
    sys.modules["torch_ema"] = _torch_ema_module()
    paths = [os.path.join(REFDIR, "refpy")]
    if ops == "reference":
        paths = [os.path.join(REFDIR, "refpy_ops")] + paths + [REFDIR]       # _raymarching.so etc. import as top-level modules
    else:
        paths = [PKG] + paths
    for p in reversed(paths + [ROOT]):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import torch
    # nerf/utils.py:183-195 decorates two colour-space helpers (never called on this path) with torch.jit.script, which needs source text;
    # the trees are sourceless by design, so scripting is the identity here
    torch.jit.script = lambda fn, *a, **k: fn
    import warnings
    warnings.filterwarnings("ignore", category=FutureWarning)        # torch.cuda.amp.custom_fwd deprecation in the reference's wrappers


def make_opt(**over):
    """the option namespace `main.py --text <prompt> -O` produces (tests/golden/options_O.json was written by its own argparse block)"""
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "options_O.json")))
    d = d.get("opt", d)
    # what main.py:186-254 derives from the parsed arguments for a text-only run (no --image / --image_config / --dmtet)
    d.setdefault("images", None)
    for k in ("ref_radii", "ref_polars", "ref_azimuths", "zero123_ws"):
        d.setdefault(k, [])
    d.setdefault("default_zero123_w", 1)
    d["exp_start_iter"] = d.get("exp_start_iter") or 0
    d["exp_end_iter"] = d.get("exp_end_iter") or d["iters"]
    if d.get("text") is None:
        d["text"] = "a hamburger"
    d.update(over)
    return types.SimpleNamespace(**d)


# ------------------------------------------------------------------------------------------------------------------ guidance arms
def make_guidance(ops, device, opt, sd_seed=0, small=False):
    """'dropin': the product's guidance.sd_utils.StableDiffusion; 'reference': the same weights in the PyTorch restatement of the
    CompVis modules (oracle/sd_ref.py, pinned to the vendored ldm code), fp16 on cuDNN / cuBLAS / SDPA, drawing its random numbers
    in the order guidance/sd_utils.py:95-103 does (posterior sample -> t -> noise) from the default CUDA generator."""
    import torch
    if PKG not in sys.path:
        sys.path.append(PKG)          # LAST: only sdf_b200.* / guidance.* may resolve there in the all-reference arm (weights + shapes)
    from sdf_b200 import sd_engine as E
    from guidance import sd_utils as S
    ucfg, vcfg = (SMALL_UNET, SMALL_VAE) if small else (E.UNET_SD15, E.VAE_SD15)
    usd = E.random_state(S.unet_param_shapes(ucfg), device, seed=sd_seed)
    vsd = E.random_state(S.vae_param_shapes(vcfg), device, seed=sd_seed + 1)
    if ops == "dropin":
        if small:
            g = S.StableDiffusion.__new__(S.StableDiffusion)
            torch.nn.Module.__init__(g)
            g.device = device
            g.engine = E.SDSEngine(usd, vsd, device, ucfg, vcfg, n_views=opt.batch_size, render_hw=opt.h, ctx_len=77, vae_res=512, capture=True)
            g.min_step, g.max_step = 20, 980
            g.text_encoder, g._synthetic_text = None, True
            return g
        return S.StableDiffusion(device, opt.fp16, opt.vram_O, opt.sd_version, opt.hf_key, opt.t_range, weights={"unet": usd, "vae": vsd},
                                 n_views=opt.batch_size, render_hw=opt.h, synthetic_text=True)
    from oracle import sd_ref

    class RefSD(torch.nn.Module):
        def __init__(self):
            super().__init__()
            with torch.device(device):
                self.unet = sd_ref.UNet(**ucfg)
                self.vae = sd_ref.VaeEncoder(**vcfg)
            self.unet.load_state_dict(usd)
            self.vae.load_state_dict(vsd)
            self.unet = self.unet.half().eval().requires_grad_(False)
            self.vae = self.vae.half().eval().requires_grad_(False)
            sd_ref.CrossAttention.use_sdpa = True
            self.acp = sd_ref.alphas_cumprod().to(device)
            self.min_step, self.max_step = 20, 980
            self.device = device
            self.text_encoder, self._synthetic_text = None, True

        get_text_embeds = S.StableDiffusion.get_text_embeds

        def train_step(self, text_embeddings, pred_rgb, guidance_scale=100, as_latent=False, grad_scale=1, save_guidance_path=None):
            B = pred_rgb.shape[0]
            post = None if as_latent else torch.randn(B, 4, 64, 64, device=device)
            t = torch.randint(self.min_step, self.max_step + 1, (B,), dtype=torch.long, device=device)
            noise = torch.randn(B, 4, 64, 64, device=device)
            with torch.autocast("cuda", dtype=torch.float16):
                loss, _, _ = sd_ref.sds_train_step(self.unet, self.vae, self.acp, text_embeddings.half(), pred_rgb, t, noise, post,
                                                   float(guidance_scale), bool(as_latent), float(grad_scale))
            return loss

    return RefSD()


SMALL_UNET = dict(in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1, attention_resolutions=(1, 2), channel_mult=(1, 2),
                  num_heads=2, context_dim=768)
SMALL_VAE = dict(ch=32, ch_mult=(1, 2, 4, 4), num_res_blocks=1, in_channels=3, z_channels=4)


# ------------------------------------------------------------------------------------------------------------------ commands
def _np(t):
    return t.detach().float().cpu().numpy()


def _load_state(model, path):
    import numpy as np
    import torch
    z = np.load(path)
    sd = model.state_dict()
    for k in sd:
        if k in z.files:
            sd[k].copy_(torch.from_numpy(z[k]).to(sd[k].device).to(sd[k].dtype))
    model.mean_density = float(z["mean_density"]) if "mean_density" in z.files else 0.0


def cmd_render(spec):
    """spec: ops, out, state (npz with the parameters, density_grid, density_bitfield [optional: written by this run when absent]),
    cases: [{shading, H, pose:[16], fovy, ambient, bg:[3]|null, seed}], table_init ('default' | 'wide'), grad_scale"""
    import numpy as np
    import torch
    install(spec["ops"])
    from nerf.network_grid import NeRFNetwork
    from nerf.utils import get_rays
    import raymarching
    dev = torch.device("cuda:0")
    opt = make_opt(**spec.get("opt", {}))
    torch.manual_seed(spec.get("seed", 0))
    model = NeRFNetwork(opt).to(dev)
    out = {}
    if spec.get("state") and os.path.exists(spec["state"]):
        _load_state(model, spec["state"])
    else:
        with torch.no_grad():
            if spec.get("table_init", "default") == "wide":
                g = torch.Generator(device="cpu").manual_seed(5)
                model.encoder.embeddings.copy_(((torch.rand(model.encoder.embeddings.shape, generator=g) - 0.5) * spec.get("table_amp", 1.0)).to(dev))
        with torch.autocast("cuda", dtype=torch.float16):
            for _ in range(spec.get("occupancy_updates", 1)):
                model.update_extra_state()
        if spec.get("state"):
            sd = {k: _np(v) if v.dtype != torch.uint8 else v.cpu().numpy() for k, v in model.state_dict().items()}
            np.savez(spec["state"], mean_density=np.float32(model.mean_density), **sd)
    if spec.get("refresh_occupancy"):
        # this arm's own update_extra_state on the loaded parameters (compared across arms with a tolerance)
        torch.manual_seed(11)
        model.density_grid.zero_()
        with torch.autocast("cuda", dtype=torch.float16):
            model.update_extra_state()
        out["refresh.density_grid"] = _np(model.density_grid)
        out["refresh.bitfield"] = model.density_bitfield.cpu().numpy()
        out["refresh.mean_density"] = np.float32(model.mean_density)
        _load_state(model, spec["state"])
    model.train()
    S = float(spec.get("grad_scale", 128.0))
    for ci, c in enumerate(spec["cases"]):
        H = W = int(c["H"])
        pose = torch.tensor(c["pose"], dtype=torch.float32, device=dev).view(1, 4, 4)
        focal = H / (2 * np.tan(np.deg2rad(c["fovy"]) / 2))
        rays = get_rays(pose, np.array([focal, focal, H / 2, W / 2]), H, W, -1)
        ro, rd = rays["rays_o"], rays["rays_d"]
        for p in model.parameters():
            p.grad = None
        bg = None if c.get("bg") is None else torch.tensor(c["bg"], dtype=torch.float32, device=dev)
        torch.manual_seed(c.get("seed", 0))
        with torch.autocast("cuda", dtype=torch.float16):
            res = model.render(ro, rd, None, H, W, staged=False, perturb=bool(c.get("perturb", True)), bg_color=bg, ambient_ratio=float(c["ambient"]),
                               shading=c["shading"], binarize=False)
        img, ws, dep, wts = res["image"], res["weights_sum"], res["depth"], res["weights"]
        g = torch.Generator(device="cpu").manual_seed(100 + ci)
        G = torch.randn(img.shape, generator=g).to(dev)
        G2 = torch.randn(ws.shape, generator=g).to(dev)
        loss = (img.float() * G).sum() + (ws.float() * G2).sum()
        if "loss_orient" in res:
            loss = loss + 10.0 * res["loss_orient"]
        k = f"c{ci}."
        out[k + "image"], out[k + "weights_sum"], out[k + "depth"], out[k + "M"] = _np(img), _np(ws), _np(dep), np.int64(wts.shape[0])
        if "loss_orient" in res:
            out[k + "loss_orient"] = _np(res["loss_orient"])
        if c.get("grad", True):
            (loss * S).backward()
            for n, p in model.named_parameters():
                if p.grad is not None:
                    out[k + "grad." + n] = _np(p.grad) / S
        # per-ray sample counts of the same march (same seed -> same jitter)
        torch.manual_seed(c.get("seed", 0))
        _ = torch.randn(3, device=dev)                  # the light direction draw that precedes the march in run_cuda
        nears, fars = raymarching.near_far_from_aabb(ro.view(-1, 3).contiguous(), rd.view(-1, 3).contiguous(), model.aabb_train)
        _, _, _, rr = raymarching.march_rays_train(ro.view(-1, 3).contiguous(), rd.view(-1, 3).contiguous(), model.bound, model.density_bitfield, model.cascade,
                                                   model.grid_size, nears, fars, bool(c.get("perturb", True)), opt.dt_gamma, opt.max_steps)
        out[k + "counts"] = rr[:, 1].cpu().numpy()
    if spec.get("eval_cases"):
        model.eval()
        for ci, c in enumerate(spec["eval_cases"]):
            H = W = int(c["H"])
            pose = torch.tensor(c["pose"], dtype=torch.float32, device=dev).view(1, 4, 4)
            focal = H / (2 * np.tan(np.deg2rad(c["fovy"]) / 2))
            rays = get_rays(pose, np.array([focal, focal, H / 2, W / 2]), H, W, -1)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                light = torch.tensor(c["light"], dtype=torch.float32, device=dev).view(1, 3) if c.get("light") else None
                res = model.render(rays["rays_o"], rays["rays_d"], None, H, W, staged=True, perturb=False, bg_color=None, light_d=light,
                                   ambient_ratio=float(c["ambient"]), shading=c["shading"])
            out[f"e{ci}.image"], out[f"e{ci}.weights_sum"], out[f"e{ci}.depth"] = _np(res["image"]), _np(res["weights_sum"]), _np(res["depth"])
    torch.cuda.synchronize()
    np.savez(spec["out"], **out)


def _sparse_rows(t):
    """table gradient as (row indices, rows): a few thousand points touch a small part of the 6.1 M-row table"""
    import numpy as np
    a = t.detach().float().cpu().numpy()
    idx = np.flatnonzero(np.abs(a).sum(-1) > 0).astype(np.int64)
    return idx, a[idx]


def cmd_points(spec):
    """NeRFNetwork.forward(x, d, l, ratio, shading) of the reference class at given points (+ parameter gradients for given upstream
    gradients): the comparand of the fused field kernels.  spec: ops, state, points (npz: x, l, gs, gc, gn), runs: [{shading, autocast,
    max_level, grad}], ratio, out."""
    import numpy as np
    import torch
    install(spec["ops"])
    from nerf.network_grid import NeRFNetwork
    dev = torch.device("cuda:0")
    opt = make_opt(**spec.get("opt", {}))
    model = NeRFNetwork(opt).to(dev)
    _load_state(model, spec["state"])
    model.train()
    z = np.load(spec["points"])
    T = lambda k: torch.from_numpy(z[k]).to(dev)
    x, l, gs, gc, gn = T("x"), T("l"), T("gs"), T("gc"), T("gn")
    d = torch.nn.functional.normalize(torch.ones_like(x), dim=-1)
    out = {}
    for ri, r in enumerate(spec["runs"]):
        model.max_level = r.get("max_level")
        for p in model.parameters():
            p.grad = None
        n = int(r.get("n", x.shape[0]))
        with torch.autocast("cuda", dtype=torch.float16, enabled=bool(r["autocast"])):
            if r["shading"] == "density":
                sig, col, nrm = model.density(x[:n])["sigma"], None, None
            else:
                sig, col, nrm = model(x[:n], d[:n], l[:n], ratio=float(spec.get("ratio", 0.3)), shading=r["shading"])
        k = f"r{ri}."
        out[k + "sigma"] = _np(sig)
        if col is not None:
            out[k + "color"] = _np(col)
        if nrm is not None:
            out[k + "normal"] = _np(nrm)
        if r.get("grad"):
            S = float(spec.get("grad_scale", 1.0)) if r["autocast"] else 1.0
            loss = (sig.float() * gs[:n]).sum() + (col.float() * gc[:n]).sum()
            if nrm is not None:
                loss = loss + (nrm.float() * gn[:n]).sum()
            (loss * S).backward()
            for name, p in model.named_parameters():
                if p.grad is None:
                    continue
                if name == "encoder.embeddings":
                    out[k + "grad.table.idx"], rows = _sparse_rows(p.grad)
                    out[k + "grad.table.rows"] = rows / S
                else:
                    out[k + "grad." + name] = _np(p.grad) / S
    torch.cuda.synchronize()
    np.savez(spec["out"], **out)


def build_trainer(spec, dev):
    """main.py:363-410: model, Adan through the optimizer lambda, guidance ModuleDict, Trainer(..., ema_decay=0.95, fp16=opt.fp16)"""
    import torch
    from nerf.network_grid import NeRFNetwork
    from nerf.provider import NeRFDataset
    from nerf.utils import Trainer, seed_everything
    from optimizer import Adan
    opt = make_opt(**spec.get("opt", {}))
    opt.workspace = spec["workspace"]
    seed_everything(int(spec.get("seed", 0)))
    model = NeRFNetwork(opt).to(dev)
    loader = NeRFDataset(opt, device=dev, type="train", H=opt.h, W=opt.w, size=int(spec["n_steps"]) * opt.batch_size).dataloader()
    optimizer = lambda m: Adan(m.get_params(5 * opt.lr), eps=1e-8, weight_decay=2e-5, max_grad_norm=5.0, foreach=False)
    scheduler = lambda o: torch.optim.lr_scheduler.LambdaLR(o, lambda it: 1)
    guidance = torch.nn.ModuleDict()
    guidance["SD"] = make_guidance(spec["ops"], dev, opt, small=bool(spec.get("small_sd", False)))
    trainer = Trainer("ref_harness", "df", opt, model, guidance, device=dev, workspace=opt.workspace, optimizer=optimizer, ema_decay=0.95,
                      fp16=opt.fp16, lr_scheduler=scheduler, use_checkpoint="scratch", scheduler_update_every_step=True, use_tensorboardX=False,
                      mute=True)
    if spec.get("state") and os.path.exists(spec["state"]):
        _load_state(model, spec["state"])
    return opt, model, loader, trainer


def cmd_steps(spec):
    import numpy as np
    import torch
    install(spec["ops"])
    dev = torch.device("cuda:0")
    opt, model, loader, trainer = build_trainer(spec, dev)
    if spec.get("state") and not os.path.exists(spec["state"]):
        sd = {k: _np(v) if v.dtype != torch.uint8 else v.cpu().numpy() for k, v in model.state_dict().items()}
        np.savez(spec["state"], **sd)
    if spec.get("global_step"):
        trainer.global_step = int(spec["global_step"])       # e.g. past the latent phase
    if spec.get("init_scale"):
        # GradScaler's default 65536 overflows the fp16 backward of the first steps in BOTH arms (the reference skips those steps and halves
        # the scale); a smaller initial scale makes the very first steps comparable instead of skipped
        trainer.scaler = torch.cuda.amp.GradScaler(init_scale=float(spec["init_scale"]), enabled=trainer.fp16)
    out, rec = {}, {"i": 0}
    real_render = model.render
    real_train_step = trainer.train_step

    def render(rays_o, rays_d, mvp, H, W, **kw):
        i = rec["i"]
        out[f"s{i}.rays_o"], out[f"s{i}.rays_d"] = _np(rays_o), _np(rays_d)
        out[f"s{i}.shading"] = np.array(kw["shading"])
        out[f"s{i}.ambient"] = np.float32(kw["ambient_ratio"])
        out[f"s{i}.bg"] = np.zeros(0, np.float32) if kw["bg_color"] is None else _np(kw["bg_color"])
        out[f"s{i}.bitfield"] = model.density_bitfield.cpu().numpy()
        torch.manual_seed(1000 + i)              # pin the draws inside render + guidance (light, jitter, posterior, t, noise)
        res = real_render(rays_o, rays_d, mvp, H, W, **kw)
        out[f"s{i}.M"] = np.int64(res["weights"].shape[0])
        return res

    def train_step(data, save_guidance_path=None):
        i = rec["i"]
        out[f"s{i}.poses_azimuth"] = _np(data["azimuth"]) if torch.is_tensor(data["azimuth"]) else np.asarray(data["azimuth"], np.float32)
        pred_rgb, pred_depth, loss = real_train_step(data, save_guidance_path=save_guidance_path)
        out[f"s{i}.pred_rgb"], out[f"s{i}.loss"] = _np(pred_rgb), _np(loss)
        scale = float(trainer.scaler.get_scale()) if trainer.fp16 else 1.0

        def hook(g, i=i, scale=scale):
            out[f"s{i}.d_pred_rgb"] = _np(g) / scale
        pred_rgb.register_hook(hook)
        rec["i"] = i + 1
        return pred_rgb, pred_depth, loss

    model.render = render
    trainer.train_step = train_step
    # building the two arms consumes the global generators differently (PyTorch module initialisers vs the engine's own generator):
    # both arms start the epoch from the same state, so the loader draws the same cameras and the Trainer the same schedule
    from nerf.utils import seed_everything
    seed_everything(int(spec.get("seed", 0)) + 17)
    trainer.train_one_epoch(loader, 1)
    torch.cuda.synchronize()
    for n, p in model.named_parameters():
        out["final." + n] = _np(p)
    out["final.density_bitfield"] = model.density_bitfield.cpu().numpy()
    if trainer.ema is not None:
        for (n, _), s in zip(model.named_parameters(), trainer.ema.shadow_params):
            out["ema." + n] = _np(s)
    out["scale"] = np.float32(trainer.scaler.get_scale()) if trainer.fp16 else np.float32(1)
    np.savez(spec["out"], **out)


def cmd_time(spec):
    """the reference's own training loop (Trainer.train_one_epoch, unmodified) timed on this GPU: warm-up epoch, then a timed epoch"""
    import torch
    install(spec["ops"])
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = bool(spec.get("cudnn_benchmark", False))
    spec = dict(spec, n_steps=int(spec["warmup"]))
    opt, model, loader, trainer = build_trainer(spec, dev)
    from nerf.provider import NeRFDataset
    trainer.global_step = int(spec.get("global_step", 0))
    trainer.train_one_epoch(loader, 1)                      # warm-up: cuDNN heuristics, allocator, first occupancy refresh
    timed = NeRFDataset(opt, device=dev, type="train", H=opt.h, W=opt.w, size=int(spec["steps"]) * opt.batch_size).dataloader()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    trainer.train_one_epoch(timed, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / int(spec["steps"])
    json.dump({"ms_per_step": ms, "steps": int(spec["steps"]), "warmup": int(spec["warmup"]), "ops": spec["ops"],
               "global_step_start": int(spec.get("global_step", 0))}, open(spec["out"], "w"))


def cmd_dmtet(spec):
    """spec: ops='dropin', out, state (npz: parameters incl. sdf / deform), tet_grid_size, cases: [{shading, H, pose:[16], fovy, ambient, bg:[3],
    light:[3], g_seed}], lambda_n, lambda_l"""
    import tempfile
    import numpy as np
    import torch
    assert spec["ops"] == "dropin", "nvdiffrast exists here only as the drop-in package"
    install("dropin")
    from sdf_b200 import tetgrid
    size = int(spec["tet_grid_size"])
    verts, tets = tetgrid.make_tet_grid(tetgrid.cells_for(size))
    work = tempfile.mkdtemp(prefix="sdf_dmtet_")
    os.makedirs(os.path.join(work, "tets"))
    np.savez(os.path.join(work, "tets", f"{size}_tets.npz"), vertices=(-verts / 2).astype(np.float32), indices=tets)     # nerf/renderer.py:293-294 flips it back
    os.chdir(work)
    from nerf.network_grid import NeRFNetwork
    from nerf.utils import get_rays
    if spec.get("no_antialias"):               # debugging aid: isolate the rasterise / interpolate half
        import nvdiffrast.torch as _dr
        _dr.antialias = lambda color, *a, **k: color
    dev = torch.device("cuda:0")
    opt = make_opt(dmtet=True, tet_grid_size=size, lock_geo=False, **spec.get("opt", {}))
    torch.manual_seed(spec.get("seed", 0))
    model = NeRFNetwork(opt).to(dev)
    _load_state(model, spec["state"])
    model.train()
    out = {}
    for ci, c in enumerate(spec["cases"]):
        H = W = int(c["H"])
        pose = torch.tensor(c["pose"], dtype=torch.float32, device=dev).view(1, 4, 4)
        focal = H / (2 * np.tan(np.deg2rad(c["fovy"]) / 2))
        rays = get_rays(pose, np.array([focal, focal, H / 2, W / 2]), H, W, -1)
        near, far = float(opt.min_near), 1000.0
        proj = torch.tensor([[2 * focal / W, 0, 0, 0], [0, -2 * focal / H, 0, 0], [0, 0, -(far + near) / (far - near), -(2 * far * near) / (far - near)],
                             [0, 0, -1, 0]], dtype=torch.float32, device=dev).unsqueeze(0)                           # nerf/provider.py:222-229
        mvp = proj @ torch.inverse(pose)
        for p in model.parameters():
            p.grad = None
        bg = torch.tensor(c["bg"], dtype=torch.float32, device=dev)
        light = torch.tensor(c["light"], dtype=torch.float32, device=dev).view(1, 1, 1, 3)
        # autocast=False isolates the algorithm: under fp16 autocast the reference's clip-space bmm (nerf/renderer.py:893-894) rounds vertex positions
        # to half precision, which moves a handful of silhouette pixels to the neighbouring triangle / the background
        with torch.autocast("cuda", dtype=torch.float16, enabled=bool(spec.get("autocast", True))):
            res = model.render(rays["rays_o"], rays["rays_d"], mvp, H, W, staged=False, bg_color=bg, light_d=light, ambient_ratio=float(c["ambient"]),
                               shading=c["shading"])
        img = res["image"]
        g = torch.Generator(device="cpu").manual_seed(int(c.get("g_seed", 100 + ci)))
        G = torch.randn(1, H, W, 3, generator=g).to(dev) * float(spec.get("g_scale", 1.0))
        loss = (img.float() * G).sum() + float(spec.get("lambda_n", 0.0)) * res["normal_loss"] + float(spec.get("lambda_l", 0.0)) * res["lap_loss"]
        loss.backward()
        k = f"c{ci}."
        out[k + "image"], out[k + "weights_sum"] = _np(img), _np(res["weights_sum"])
        out[k + "normal_loss"], out[k + "lap_loss"] = _np(res["normal_loss"]), _np(res["lap_loss"])
        for n, p in model.named_parameters():
            if p.grad is not None:
                out[k + "grad." + n] = _np(p.grad)
    torch.cuda.synchronize()
    np.savez(spec["out"], **out)


def run_subprocess(spec, timeout=1800):
    """used by tests / bench: run one command in a fresh interpreter (module names are global)"""
    import subprocess
    p = subprocess.run([sys.executable, os.path.abspath(__file__), json.dumps(spec)], capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"ref_harness {spec.get('cmd')} ({spec.get('ops')}) failed:\n{p.stdout[-3000:]}\n{p.stderr[-6000:]}")
    return p


COMMANDS = {"render": cmd_render, "points": cmd_points, "steps": cmd_steps, "time": cmd_time, "dmtet": cmd_dmtet}

if __name__ == "__main__":
    if HERE in sys.path:
        sys.path.remove(HERE)             # the script directory would make `import oracle` resolve to oracle/oracle.py instead of the package
    _spec = json.loads(sys.argv[1])
    if _spec["cmd"] == "multi":                 # several commands of ONE arm in one interpreter (torch import + CUDA init paid once)
        for _s in _spec["specs"]:
            COMMANDS[_s["cmd"]](dict(_s, ops=_spec["ops"]))
    else:
        COMMANDS[_spec["cmd"]](_spec)
