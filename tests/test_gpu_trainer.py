"""GPU: the SDS training step end to end (sdf_b200/trainer.py) on a reduced SD configuration: every shading mode of the
schedule runs, the loss is finite, every NeRF parameter receives a gradient and moves, and the drop-in guidance.sd_utils
StableDiffusion.train_step back-propagates the engine's gradient."""
import pytest
import torch

from sdf_b200 import sd_engine as E
from sdf_b200.options import default_opt
from sdf_b200.trainer import SDSTrainer

pytestmark = pytest.mark.gpu

SMALL_UNET = dict(in_channels=4, model_channels=64, out_channels=4, num_res_blocks=1, attention_resolutions=(1, 2), channel_mult=(1, 2),
                  num_heads=2, context_dim=768)
SMALL_VAE = dict(ch=32, ch_mult=(1, 2, 4, 4), num_res_blocks=1, in_channels=3, z_channels=4)


class SmallGuidance(torch.nn.Module):
    """guidance.sd_utils.StableDiffusion with a reduced architecture (same code path, fewer channels)"""

    def __init__(self, device, render_hw=64):
        super().__init__()
        from guidance import sd_utils as S
        self.S = S
        self.device = device
        usd = E.random_state(S.unet_param_shapes(SMALL_UNET), device, seed=0)
        vsd = E.random_state(S.vae_param_shapes(SMALL_VAE), device, seed=1)
        self.engine = E.SDSEngine(usd, vsd, device, SMALL_UNET, SMALL_VAE, n_views=1, render_hw=render_hw, ctx_len=77, vae_res=512, capture=True)
        self.min_step, self.max_step = 20, 980
        self.text_encoder, self._synthetic_text = None, True

    def get_text_embeds(self, prompt):
        return self.S.StableDiffusion.get_text_embeds(self, prompt)

    def train_step(self, text_embeddings, pred_rgb, guidance_scale=100, as_latent=False, grad_scale=1, save_guidance_path=None):
        self.engine.set_text(text_embeddings)
        return self.S._SDSLoss.apply(pred_rgb, self, bool(as_latent), guidance_scale, grad_scale)


def test_training_steps_all_shadings(device):
    opt = default_opt(h=64, w=64)
    guidance = SmallGuidance(device)
    tr = SDSTrainer(opt, device, guidance, seed=0)
    before = {n: p.detach().clone() for n, p in tr.model.named_parameters()}
    losses = []
    for sh in ["latent", "lambertian", "textureless", "albedo", "lambertian"]:
        losses.append(tr.train_step(shading=sh, read_loss=True))
        assert tr.last_M > 1000
    assert all(l == l and abs(l) < 1e12 for l in losses), losses
    for n, p in tr.model.named_parameters():
        assert torch.isfinite(p).all(), n
        assert (p.detach() - before[n]).abs().max().item() > 0, f"{n} did not move"
    # Adan's decoupled weight decay touches every table entry, sampled or not
    moved = ((tr.model.encoder.embeddings.detach() - before["encoder.embeddings"]).abs().sum(-1) > 0).float().mean().item()
    assert moved > 0.5


def test_schedule_default_path_and_occupancy_refresh(device):
    opt = default_opt(h=64, w=64, iters=10)         # iters=10: step 1-2 latent, then shaded
    guidance = SmallGuidance(device)
    tr = SDSTrainer(opt, device, guidance, seed=1)
    for _ in range(4):
        l = tr.train_step(read_loss=True)
        assert l == l
    assert tr.model.mean_density > 0
    assert int(tr.model.density_bitfield.sum()) > 0
    # the fp16 working copy of the table is maintained by the fused Adan step
    assert torch.equal(tr.model.table_half(), tr.model.encoder.embeddings.detach().half())
    # EMA shadow exists and follows torch_ema's per-epoch schedule (steps_per_epoch = 100 by default: untouched after 4 steps)
    emb = tr.model.encoder.embeddings
    assert id(emb) in tr.optimizer.ema_shadow and tr.optimizer.ema_num_updates == 0


def test_ema_rides_in_the_adan_pass_on_epoch_boundaries(device):
    opt = default_opt(h=64, w=64)
    guidance = SmallGuidance(device)
    tr = SDSTrainer(opt, device, guidance, seed=4, steps_per_epoch=2)
    emb = tr.model.encoder.embeddings
    s0 = tr.optimizer.ema_shadow[id(emb)].clone()
    tr.train_step(shading="albedo")
    assert torch.equal(tr.optimizer.ema_shadow[id(emb)], s0)
    tr.train_step(shading="albedo")                         # step 2 = end of epoch 1: shadow -= (1 - min(.95, 2/11)) (shadow - p)
    d = min(0.95, 2 / 11)
    exp = s0.double() - (1 - d) * (s0.double() - emb.detach().double())
    assert (tr.optimizer.ema_shadow[id(emb)].double() - exp).abs().max().item() < 1e-6
    assert tr.optimizer.ema_num_updates == 1


def test_no_host_synchronisation_inside_a_step(device):
    """the training step enqueues without blocking: with the GPU held busy by a long-running kernel, train_step() returns before that
    kernel has finished (a .item() / cudaStreamSynchronize anywhere in the step would wait for it)"""
    opt = default_opt(h=64, w=64)
    guidance = SmallGuidance(device)
    tr = SDSTrainer(opt, device, guidance, seed=5)
    for _ in range(3):
        tr.train_step(shading="lambertian")                # warm-up: allocations, graph capture, first occupancy refresh
    torch.cuda.synchronize()
    done = torch.cuda.Event()
    torch.cuda._sleep(int(2e9))                             # ~1 s of GPU time ahead of the step
    done.record()
    tr.train_step(shading="lambertian")
    assert not done.query(), "train_step blocked on the device"
    torch.cuda.synchronize()


def test_config_c3_render_size(device):
    """BASELINE.json config C3 renders 128x128 per view: same step at the larger render size (bilinear 128 -> 512 into the VAE)"""
    opt = default_opt(h=128, w=128)
    guidance = SmallGuidance(device, render_hw=128)
    tr = SDSTrainer(opt, device, guidance, seed=2)
    l1 = tr.train_step(shading="lambertian", read_loss=True)
    l2 = tr.train_step(shading="latent", read_loss=True)
    assert l1 == l1 and l2 == l2 and tr.last_M > 4000
    assert all(torch.isfinite(p).all() for p in tr.model.parameters())
