"""Generates tests/golden/sd_small.npz by running the reference's OWN vendored CompVis modules
(/root/reference/ldm/...: UNetModel, Encoder) on CPU with seeded weights/inputs at a reduced size, and
checks oracle/sd_ref.py against them on the way (same state dict -> same outputs).

Run in the authoring container only (needs /root/reference):   python tests/golden/make_golden_sd.py
The fixture stores weights + inputs + the reference outputs; tests/test_sd_ref_golden.py replays it
anywhere (no reference tree needed).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SDF_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)


def _shim(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_ldm():
    # packages the vendored ldm imports at module top but the UNet/Encoder forward never touches
    for n in ["matplotlib", "matplotlib.pyplot", "omegaconf", "omegaconf.listconfig", "pytorch_lightning", "taming", "kornia", "clip"]:
        if n not in sys.modules:
            try:
                __import__(n)
            except Exception:
                _shim(n)
    sys.modules["omegaconf.listconfig"].__dict__.setdefault("ListConfig", type("ListConfig", (), {}))
    sys.modules["omegaconf"].__dict__.setdefault("ListConfig", sys.modules["omegaconf.listconfig"].ListConfig)
    sys.path.insert(0, REF)
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from ldm.modules.diffusionmodules.model import Encoder
    from ldm.modules.diffusionmodules.util import make_beta_schedule
    return UNetModel, Encoder, make_beta_schedule


def main():
    UNetModel, Encoder, make_beta_schedule = import_ldm()
    from oracle import sd_ref

    torch.manual_seed(0)
    cfg = dict(in_channels=4, model_channels=32, out_channels=4, num_res_blocks=1, attention_resolutions=(1, 2),
               channel_mult=(1, 2), num_heads=2, context_dim=16)
    ref_unet = UNetModel(image_size=16, use_spatial_transformer=True, transformer_depth=1, legacy=False, use_checkpoint=False, **cfg).eval()
    mine = sd_ref.UNet(**cfg).eval()
    # re-randomise the zero-initialised layers through OUR helper, then copy the weights into the reference
    mine.load_state_dict(ref_unet.state_dict())
    sd_ref.reinit_zero_modules(mine, seed=1)
    ref_unet.load_state_dict(mine.state_dict())
    x = torch.randn(2, 4, 16, 16)
    t = torch.tensor([37, 911])
    ctx = torch.randn(2, 5, 16)
    with torch.no_grad():
        y_ref = ref_unet(x, t, context=ctx)
        y_mine = mine(x, t, ctx)
    assert y_ref.abs().max() > 1e-3
    err = (y_ref - y_mine).abs().max().item()
    print("unet max|ref - restatement| =", err, " |y|max =", y_ref.abs().max().item())
    assert err < 1e-5

    vcfg = dict(ch=32, ch_mult=(1, 2, 2), num_res_blocks=1, in_channels=3, z_channels=4)
    ref_enc = Encoder(out_ch=3, attn_resolutions=[], dropout=0.0, resolution=32, double_z=True, **vcfg).eval()
    quant = torch.nn.Conv2d(8, 8, 1)
    my_enc = sd_ref.VaeEncoder(**vcfg).eval()
    sd = {k: v for k, v in ref_enc.state_dict().items()}
    sd["quant_conv.weight"], sd["quant_conv.bias"] = quant.weight.data, quant.bias.data
    my_enc.load_state_dict(sd)
    img = torch.rand(1, 3, 32, 32) * 2 - 1
    img.requires_grad_(True)
    m_ref = quant(ref_enc(img))
    g_up = torch.randn_like(m_ref)
    (gi_ref,) = torch.autograd.grad((m_ref * g_up).sum(), img)
    img2 = img.detach().clone().requires_grad_(True)
    m_mine = my_enc(img2)
    (gi_mine,) = torch.autograd.grad((m_mine * g_up).sum(), img2)
    print("vae  max|ref - restatement| =", (m_ref - m_mine).abs().max().item(), " grad:", (gi_ref - gi_mine).abs().max().item())
    assert (m_ref - m_mine).abs().max().item() < 1e-5 and (gi_ref - gi_mine).abs().max().item() < 1e-5

    betas = make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    acp_ref = np.cumprod(1.0 - betas, axis=0)
    acp = sd_ref.alphas_cumprod().numpy()
    assert np.abs(acp - acp_ref).max() < 1e-6
    temb_ref = None
    from ldm.modules.diffusionmodules.util import timestep_embedding
    assert torch.equal(timestep_embedding(t, 32), sd_ref.timestep_embedding(t, 32))

    out = {"unet_x": x.numpy(), "unet_t": t.numpy(), "unet_ctx": ctx.numpy(), "unet_y": y_ref.numpy(),
           "vae_img": img.detach().numpy(), "vae_moments": m_ref.detach().numpy(), "vae_gup": g_up.numpy(), "vae_gimg": gi_ref.numpy(),
           "acp_idx": np.array([0, 20, 500, 980, 999]), "acp_val": acp_ref[[0, 20, 500, 980, 999]].astype(np.float64)}
    for k, v in mine.state_dict().items():
        out["unet_w/" + k] = v.numpy()
    for k, v in my_enc.state_dict().items():
        out["vae_w/" + k] = v.numpy()
    path = os.path.join(HERE, "sd_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
