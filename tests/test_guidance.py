"""SDS x0-reconstruction loss: `animate3d_b200/guidance.py` against the reference's own
`compute_mvdream_recon_loss` body (animatemv_guidance.py:391-513) executed with a deterministic stand-in UNet
(tests/golden/gen_reference_goldens.py::run_reference_recon_loss).  Host logic only: runs on the CPU."""
import os
from types import SimpleNamespace

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
    from animate3d_b200.guidance import AnimateMVDiffusionGuidance, GuidanceConfig
    d = torch.load(os.path.join(golden_dir, "ref_guidance.pt"), weights_only=False)
    for rescale, want in d["out"].items():
        g = AnimateMVDiffusionGuidance(StubUNet(), GuidanceConfig(n_view=d["n"], n_frame=d["f"], guidance_scale=5.0,
                                                                  recon_std_rescale=rescale))
        lat = d["latents"].clone().requires_grad_(True)
        torch.manual_seed(d["seed"])          # same generator state as the reference run: same torch.randn_like noise
        loss, aux = g.compute_mvdream_recon_loss(lat, d["t"], d["text"], d["c2w"].clone(), d["img"].clone())
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
