"""SDS x0-reconstruction loss: `animate3d_b200/guidance.py` against the reference's own
`compute_mvdream_recon_loss` body (animatemv_guidance.py:391-513) executed with a deterministic stand-in UNet
(tests/golden/gen_reference_goldens.py::run_reference_recon_loss).  Host logic only: runs on the CPU."""
import os
from types import SimpleNamespace

import pytest
import torch


def fake_unet_eps(lat, ts, ehs, cam, img):   # same function as in gen_reference_goldens.py
    v = lambda x: x.reshape(-1, 1, 1, 1, 1)
    return (0.1 * lat * v(torch.cos(ts.float() / 1000.0)) + 0.01 * v(ehs.float().mean((1, 2))) + 0.02 * v(cam.float().sum(1))
            + 0.03 * v(img.float().mean(1)) + 0.05 * torch.sin(3.0 * lat))


class StubUNet:
    device = torch.device("cpu")

    def __call__(self, sample, timestep, encoder_hidden_states, camera=None, added_cond_kwargs=None, num_views=None,
                 i2v_cond_time_zero=False):
        return SimpleNamespace(sample=fake_unet_eps(sample, timestep, encoder_hidden_states, camera, added_cond_kwargs["image_embeds"]))


def test_recon_loss_matches_reference_method(golden_dir):
    from animate3d_b200.guidance import AnimateMVDiffusionGuidance
    d = torch.load(os.path.join(golden_dir, "ref_guidance.pt"), weights_only=False)
    bnf = d["latents"].shape[0]
    ele = azi = dist = torch.zeros(bnf)

    class Prompts:                            # PromptProcessorOutput stand-in: the golden's (cond, uncond) embeddings
        use_perp_neg = False

        def get_text_embeddings(self, elevation, azimuth, camera_distances, view_dependent_prompting):
            assert elevation.shape[0] == bnf // d["f"] and not view_dependent_prompting    # one entry per (b, view)
            return d["text"]

    for rescale, want in d["out"].items():
        g = AnimateMVDiffusionGuidance({"n_view": d["n"], "n_frame": d["f"], "guidance_scale": 5.0,
                                        "recon_std_rescale": rescale}, unet=StubUNet())
        lat = d["latents"].clone().requires_grad_(True)
        torch.manual_seed(d["seed"])          # same generator state as the reference run: same torch.randn_like noise
        loss, aux = g.compute_mvdream_recon_loss(lat, d["t"], Prompts(), ele, azi, dist, d["c2w"].clone(), d["img"].clone())
        loss.backward()
        torch.testing.assert_close(aux["latents_noisy"], want["latents_noisy"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(aux["noise_pred"], want["noise_pred"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(aux["latents_recon"], want["latents_recon"], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(loss.detach(), want["loss"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(lat.grad, want["grad"], rtol=1e-4, atol=1e-6)
        # frame 0 of every view is reconstructed by itself: no gradient there (animatemv_guidance.py:493-495)
        assert lat.grad.reshape(-1, d["f"], *lat.shape[1:])[:, 0].abs().max() == 0


def test_camera_cond_normalises_translation():
    from animate3d_b200.guidance import get_camera_cond
    c = torch.eye(4).repeat(3, 1, 1)
    c[:, :3, 3] = torch.tensor([[3.0, 0, 4.0], [0, 0, 2.0], [1.0, 1.0, 1.0]])
    out = get_camera_cond(c).reshape(3, 4, 4)
    torch.testing.assert_close(out[:, :3, 3].norm(dim=1), torch.ones(3), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out[:, :3, :3], torch.eye(3).repeat(3, 1, 1))


def test_plugin_call_signature_update_step_and_registry():
    """Reference surface (animatemv_guidance.py:54, 515-526, 767-793): registered name, `__call__(rgb, prompt_utils,
    elevation, azimuth, camera_distances, c2w, rgb_as_latents, guidance_eval)`, `update_step` schedules."""
    import inspect
    from animate3d_b200 import registry
    from animate3d_b200.guidance import AnimateMVDiffusionGuidance, PrecomputedPromptUtils
    assert registry.find("animatemv-diffusion-guidance") is AnimateMVDiffusionGuidance
    names = list(inspect.signature(AnimateMVDiffusionGuidance.__call__).parameters)
    assert names[:9] == ["self", "rgb", "prompt_utils", "elevation", "azimuth", "camera_distances", "c2w", "rgb_as_latents",
                         "guidance_eval"]
    n, f = 2, 3
    g = registry.find("animatemv-diffusion-guidance")({"n_view": n, "n_frame": f, "guidance_scale": 5.0,
                                                        "min_step_percent": [0, 0.5, 0.02, 100], "max_step_percent": 0.6,
                                                        "grad_clip": [0, 2.0, 8.0, 100]}, unet=StubUNet())
    assert (g.min_step, g.max_step) == (20, 980)                     # defaults until the first update_step (line 316)
    g.update_step(0, 50)
    assert g.min_step == int(1000 * (0.5 + (0.02 - 0.5) * 0.5)) and g.max_step == 600 and abs(g.grad_clip_val - 5.0) < 1e-6
    gen = torch.Generator().manual_seed(0)
    bnf = n * f
    rgb = torch.rand(bnf, 40, 40, 3, generator=gen).requires_grad_(True)
    c2w = torch.eye(4).repeat(bnf, 1, 1)
    c2w[:, :3, 3] = torch.randn(bnf, 3, generator=gen)
    pu = PrecomputedPromptUtils(torch.randn(77, 768, generator=gen), torch.zeros(77, 768))
    z = torch.zeros(bnf)
    out = g(rgb, pu, z, z, z, c2w, rgb_as_latents=True, image_embeds=torch.randn(n, 1024, generator=gen),
            timestep=torch.tensor([300]))
    assert set(out) == {"loss_sds", "min_step", "max_step"}
    out["loss_sds"].backward()
    grad = rgb.grad.reshape(n, f, 40, 40, 3)
    assert grad[:, 1:].abs().sum() > 0 and grad[:, 0].abs().max() == 0     # frame 0 carries no SDS gradient (493-495)
    with pytest.raises(ValueError):
        g(rgb, pu, z, z, z, c2w)                                      # no VAE attached and not rgb_as_latents
    with pytest.raises(KeyError):
        registry.find("animatemv-diffusion-guidance")({"no_such_option": 1}, unet=StubUNet())
    ev = g(rgb.detach(), pu, z, z, z, c2w, rgb_as_latents=True, guidance_eval=True, image_embeds=torch.randn(n, 1024, generator=gen),
           timestep=torch.tensor([120]))["eval"]
    assert ev["latents_final"].shape == (bnf, 3, 32, 32) and len(ev["texts"]) == len(ev["noise_levels"])
