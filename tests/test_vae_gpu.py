"""AutoencoderKL on the engine against oracle/vae_oracle.py (fp32 torch restatement of diffusers 0.28
AutoencoderKL, parity unpinned): encoder moments, the encoder's INPUT GRADIENT (the SDS path), decoder images, and the
diffusers-style API the guidance / pipeline use.  Tolerance: fp16 activations with fp32 accumulation -> 1e-2 relative."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(seed=0):
    from animate3d_b200.vae import AutoencoderKL
    from oracle import vae_oracle as VO
    cfg = VO.VAEConfig()
    sd = VO.make_state_dict(cfg, seed)
    vae = AutoencoderKL()
    missing, unexpected = vae.load_state_dict(sd)
    assert not missing and not unexpected
    return vae, sd, cfg, VO


def _rel(a, b):
    return ((a.float().cpu() - b).norm() / b.norm()).item()


@pytest.mark.parametrize("n,size", [(2, 256), (3, 128), (1, 512)])
def test_encoder_moments_and_input_gradient(n, size):
    vae, sd, cfg, VO = _setup()
    g = torch.Generator().manual_seed(size + n)
    x = torch.rand(n, 3, size, size, generator=g) * 2 - 1
    dm = torch.randn(n, 8, size // 8, size // 8, generator=g)
    torch.set_num_threads(16)
    xo = x.clone().requires_grad_(True)
    mo = VO.encode_moments(sd, cfg, xo)
    (mo * dm).sum().backward()
    xg = x.cuda().requires_grad_(True)
    mg = vae.encode_moments(xg)
    assert mg.shape == mo.shape and mg.dtype == torch.float32
    (mg * dm.cuda()).sum().backward()
    r_m, r_g = _rel(mg.detach(), mo.detach()), _rel(xg.grad, xo.grad)
    print(f"encoder {n}x{size}^2: moments rel-l2 {r_m:.3e}, input gradient rel-l2 {r_g:.3e}")
    assert r_m < 1e-2 and r_g < 2e-2
    mx = ((xg.grad.cpu() - xo.grad).abs().max() / xo.grad.abs().max()).item()
    assert mx < 5e-2, mx


def test_decoder_matches_oracle_and_api():
    vae, sd, cfg, VO = _setup(seed=3)
    g = torch.Generator().manual_seed(9)
    z = torch.randn(2, 4, 32, 32, generator=g)
    torch.set_num_threads(16)
    with torch.no_grad():
        ref = VO.decode(sd, cfg, z)
    img = vae.decode(z.cuda()).sample
    assert img.shape == (2, 3, 256, 256)
    r = _rel(img, ref)
    print(f"decoder: rel-l2 {r:.3e}")
    assert r < 1e-2
    # diffusers-style surface used by the guidance (animatemv_guidance.py:365-373) and the pipeline (pipeline.py:540-567)
    x = (torch.rand(2, 3, 256, 256, generator=g) * 2 - 1).cuda()
    dist = vae.encode(x).latent_dist
    gen = torch.Generator(device="cuda").manual_seed(1)
    s1 = dist.sample(generator=gen)
    assert s1.shape == (2, 4, 32, 32) and torch.equal(dist.mode(), dist.mean)
    with torch.no_grad():
        mo = VO.encode_moments(sd, cfg, x.cpu())
    mean, logvar = mo.chunk(2, 1)
    assert _rel(dist.mean, mean) < 1e-2 and vae.config.scaling_factor == 0.18215


def test_guidance_sds_gradient_reaches_the_rendered_images():
    """`AnimateMVDiffusionGuidance.__call__` end to end on the engine: rgb -> VAE encoder (with grad) -> engine UNet (no grad) ->
    x0-reconstruction loss -> gradient w.r.t. rgb, against the same pipeline assembled from the oracles (VAE oracle + UNet oracle +
    the guidance arithmetic pinned in tests/test_guidance.py)."""
    from animate3d_b200.guidance import AnimateMVDiffusionGuidance, PrecomputedPromptUtils
    from animate3d_b200.unet import MVUNetMotionModel
    from animate3d_b200.unet_config import UNetConfig
    from oracle import unet_oracle as O
    vae, vsd, vcfg, VO = _setup(seed=5)
    nv, nf = 2, 3
    ocfg = O.UNetConfig(num_views=nv, num_frames=nf)
    usd = O.make_state_dict(ocfg, 2)
    unet = MVUNetMotionModel(UNetConfig(num_views=nv, num_frames=nf))
    unet.load_state_dict(usd)
    guide = AnimateMVDiffusionGuidance({"n_view": nv, "n_frame": nf, "guidance_scale": 5.0, "recon_std_rescale": 0.5}, unet=unet, vae=vae)
    g = torch.Generator().manual_seed(4)
    bnf = nv * nf
    rgb = torch.rand(bnf, 256, 256, 3, generator=g)
    c2w = torch.eye(4).repeat(bnf, 1, 1)
    c2w[:, :3, 3] = torch.randn(bnf, 3, generator=g)
    text, unc = torch.randn(77, 768, generator=g), torch.randn(77, 768, generator=g)
    img = torch.randn(nv, 1024, generator=g)
    t = torch.tensor([150])
    z = torch.zeros(bnf)
    noise_seed = 11
    # ---- engine
    rgb_g = rgb.cuda().requires_grad_(True)
    torch.manual_seed(noise_seed)                         # posterior.sample() and the forward-diffusion noise
    out = guide(rgb_g, PrecomputedPromptUtils(text.cuda(), unc.cuda()), z.cuda(), z.cuda(), z.cuda(), c2w.cuda(), image_embeds=img.cuda(),
                timestep=t.cuda())
    out["loss_sds"].backward()

    # ---- oracles: a guidance object whose UNet / VAE are the CPU oracles
    class OracleUNet:
        device = torch.device("cpu")

        def __call__(self, sample, timestep, encoder_hidden_states, camera=None, added_cond_kwargs=None, num_views=None,
                     i2v_cond_time_zero=False):
            from types import SimpleNamespace
            with torch.no_grad():
                return SimpleNamespace(sample=O.unet_forward(usd, ocfg, sample, timestep, encoder_hidden_states, camera,
                                                             added_cond_kwargs["image_embeds"], num_views, i2v_cond_time_zero))

    class OracleVAE:
        from types import SimpleNamespace as _NS
        config = _NS(scaling_factor=0.18215)

        def encode(self, x):
            from types import SimpleNamespace
            from animate3d_b200.vae import DiagonalGaussianDistribution
            return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(VO.encode_moments(vsd, vcfg, x)))

    ref = AnimateMVDiffusionGuidance({"n_view": nv, "n_frame": nf, "guidance_scale": 5.0, "recon_std_rescale": 0.5}, unet=OracleUNet(),
                                     vae=OracleVAE())
    torch.set_num_threads(16)
    rgb_o = rgb.clone().requires_grad_(True)
    # same random draws as the engine run: the CUDA generator was seeded, so replay its draws on the CPU explicitly
    torch.manual_seed(noise_seed)
    n1 = torch.randn(bnf, 4, 32, 32, device="cuda").cpu()
    # the forward-diffusion noise is `randn_like` of a permuted VIEW (frames 1.. of "b n c f h w"): same strides, same stream
    rest_like = torch.empty(bnf, 4, 32, 32, device="cuda").reshape(1, nv, nf, 4, 32, 32).permute(0, 1, 3, 2, 4, 5)[:, :, :, 1:]
    n2 = torch.randn_like(rest_like).cpu()
    x = torch.nn.functional.interpolate(rgb_o.permute(0, 3, 1, 2), (256, 256), mode="bilinear", align_corners=False)
    mo = VO.encode_moments(vsd, vcfg, x * 2 - 1)
    lat = VO.sample_latents(mo, n1) * 0.18215
    loss_o, _ = ref._recon_loss(lat, t, torch.cat([text[None].expand(nv, -1, -1), unc[None].expand(nv, -1, -1)]), c2w, img, noise=n2)
    loss_o.backward()
    rl = abs(float(out["loss_sds"].detach()) - float(loss_o.detach())) / abs(float(loss_o.detach()))
    rg = _rel(rgb_g.grad, rgb_o.grad)
    print(f"guidance through the engine: loss {float(out['loss_sds']):.5f} vs {float(loss_o):.5f} (rel {rl:.2e}), d loss / d rgb rel-l2 {rg:.3e}")
    assert rl < 2e-2 and rg < 5e-2
    fr0 = rgb_g.grad.reshape(nv, nf, 256, 256, 3)[:, 0].abs().max().item()
    assert fr0 == 0.0                                     # frame 0 of every view carries no SDS gradient
