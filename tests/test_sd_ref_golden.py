"""CPU: oracle/sd_ref.py (restatement of the CompVis UNet / VAE-encoder on the SDS path) replayed against the
golden vectors tests/golden/sd_small.npz, which were produced by the reference's own vendored ldm modules
(tests/golden/make_golden_sd.py)."""
import os

import numpy as np
import torch

from oracle import sd_ref

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sd_small.npz"))


def _sd(prefix):
    return {k[len(prefix):]: torch.from_numpy(G[k]) for k in G.files if k.startswith(prefix)}


def test_unet_matches_reference_vectors():
    m = sd_ref.UNet(in_channels=4, model_channels=32, out_channels=4, num_res_blocks=1, attention_resolutions=(1, 2),
                    channel_mult=(1, 2), num_heads=2, context_dim=16).eval()
    m.load_state_dict(_sd("unet_w/"))
    with torch.no_grad():
        y = m(torch.from_numpy(G["unet_x"]), torch.from_numpy(G["unet_t"]), torch.from_numpy(G["unet_ctx"]))
    assert np.abs(y.numpy() - G["unet_y"]).max() < 2e-5
    assert np.abs(G["unet_y"]).max() > 0.1


def test_vae_encoder_forward_and_input_gradient():
    m = sd_ref.VaeEncoder(ch=32, ch_mult=(1, 2, 2), num_res_blocks=1, in_channels=3, z_channels=4).eval()
    m.load_state_dict(_sd("vae_w/"))
    img = torch.from_numpy(G["vae_img"]).requires_grad_(True)
    mom = m(img)
    assert np.abs(mom.detach().numpy() - G["vae_moments"]).max() < 2e-5
    (gi,) = torch.autograd.grad((mom * torch.from_numpy(G["vae_gup"])).sum(), img)
    assert np.abs(gi.numpy() - G["vae_gimg"]).max() < 2e-5


def test_schedule_and_full_size_parameter_counts():
    acp = sd_ref.alphas_cumprod().double().numpy()
    np.testing.assert_allclose(acp[G["acp_idx"]], G["acp_val"], rtol=1e-6)
    # SD-1.5 sizes (SURVEY.md §8c probes): 859 520 964 UNet parameters, 34 163 592 encoder (+72 quant_conv)
    with torch.device("meta"):
        u = sd_ref.UNet(**sd_ref.UNET_SD15)
        v = sd_ref.VaeEncoder(**sd_ref.VAE_SD15)
    assert sum(p.numel() for p in u.parameters()) == 859_520_964
    assert sum(p.numel() for p in v.parameters()) == 34_163_592 + 72


def test_sds_step_gradient_identity():
    """d loss / d latents == w(t) (eps_hat - eps)   (guidance/sd_utils.py:160-163)"""
    torch.manual_seed(0)
    unet = sd_ref.UNet(in_channels=4, model_channels=32, out_channels=4, num_res_blocks=1, attention_resolutions=(1, 2),
                       channel_mult=(1, 2), num_heads=2, context_dim=16).eval()
    sd_ref.reinit_zero_modules(unet)
    acp = sd_ref.alphas_cumprod()
    rgb = torch.rand(1, 4, 64, 64, requires_grad=True)     # as_latent path: 4-channel "image"
    t = torch.tensor([300]); noise = torch.randn(1, 4, 64, 64); ctx = torch.randn(2, 3, 16)
    loss, lat, grad = sd_ref.sds_train_step(unet, None, acp, ctx, rgb, t, noise, None, guidance_scale=7.0, as_latent=True)
    (g,) = torch.autograd.grad(loss, lat, retain_graph=True)
    assert torch.allclose(g, grad, atol=1e-6)
