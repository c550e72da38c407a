"""Host-side mirror of the SDS guidance of the reference, `AnimateMVDiffusionGuidance`
(custom/threestudio-animate3d/guidance/animatemv_guidance.py:54-793), for the part that sits on the hot path:
`compute_mvdream_recon_loss` (391-513) = noise frames 1.., one classifier-free-guided UNet evaluation with NO grad
(422-459, note the (cond, uncond) order and `text + s*(text - uncond)` at 452-459), x0 via the DDIM scheduler (466),
std-rescale (468-487) and the x0-reconstruction MSE (497-501).

The UNet evaluation runs on the sm_100a engine; the remaining arithmetic is a handful of elementwise/reduction ops on
[B*Nv*F, 4, 32, 32] latents that must stay on the autograd tape (the loss's gradient flows to `latents` and from there
through the VAE encoder to the rasterizer), so they are expressed with torch.  The VAE encoder and CLIP image encoder
(`encode_images`, 365-373 / 546-555) are SURVEY section 8(f) "next" rows and are injected as callables."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import torch
import torch.nn.functional as F

from .scheduler import DDIMScheduler
from .unet import MVUNetMotionModel


def get_camera_cond(c2w: torch.Tensor) -> torch.Tensor:
    """animatemv_guidance.py:40-52, 355-357: translation of each 4x4 c2w normalised to the unit sphere, flattened to 16."""
    cam = c2w.clone().reshape(-1, 4, 4)
    t = cam[:, :3, 3]
    cam[:, :3, 3] = t / (torch.norm(t, dim=1, keepdim=True) + 1e-8)
    return cam.reshape(-1, 16)


@dataclass
class GuidanceConfig:
    n_view: int = 4
    n_frame: int = 16
    guidance_scale: float = 5.0
    recon_std_rescale: float = 0.5
    min_step_percent: float = 0.02
    max_step_percent: float = 0.2
    i2v_cond_time_zero: bool = False
    num_train_timesteps: int = 1000


class AnimateMVDiffusionGuidance:
    """Registered in the reference as "animatemv-diffusion-guidance" (animatemv_guidance.py:54)."""

    def __init__(self, unet: MVUNetMotionModel, cfg: Optional[GuidanceConfig] = None, scheduler: Optional[DDIMScheduler] = None,
                 encode_images: Optional[Callable] = None, encode_ip_image: Optional[Callable] = None):
        self.unet, self.cfg = unet, cfg or GuidanceConfig()
        self.scheduler = scheduler or DDIMScheduler()
        self.scheduler.set_timesteps(self.cfg.num_train_timesteps)         # animatemv_guidance.py:313 -> prev = t - 1
        self.alphas = torch.from_numpy(self.scheduler.alphas_cumprod).to(unet.device)
        self.encode_images, self.encode_ip_image = encode_images, encode_ip_image
        self.min_step = int(self.cfg.num_train_timesteps * self.cfg.min_step_percent)
        self.max_step = int(self.cfg.num_train_timesteps * self.cfg.max_step_percent)

    def forward_unet(self, latents, t, encoder_hidden_states, camera, image_embeds, i2v_cond_time_zero=False):
        """animatemv_guidance.py:328-346."""
        return self.unet(latents, t, encoder_hidden_states, camera=camera, added_cond_kwargs={"image_embeds": image_embeds},
                         num_views=self.cfg.n_view, i2v_cond_time_zero=i2v_cond_time_zero).sample

    def compute_mvdream_recon_loss(self, latents, t, text_embeddings, camera, image_embeds, noise=None):
        """latents [(b n f), 4, h, w] with grad; t [b] long; text_embeddings [2*b*n, 77, 768] in (cond, uncond) order;
        camera [b*n*f, 4, 4] c2w; image_embeds [b*n, 1024].  Returns (loss, aux) like animatemv_guidance.py:391-513."""
        cfg = self.cfg
        n, f = cfg.n_view, cfg.n_frame
        bnf, c, h, w = latents.shape
        b = bnf // (n * f)
        lat = latents.reshape(b, n, f, c, h, w).permute(0, 1, 3, 2, 4, 5)          # b n c f h w   (line 414)
        first = lat[:, :, :, 0:1]
        rest = lat[:, :, :, 1:]
        with torch.no_grad():
            if noise is None:
                noise = torch.randn_like(rest)
            a = self.alphas[t].reshape(b, 1, 1, 1, 1, 1)
            rest_noisy = a.sqrt() * rest + (1 - a).sqrt() * noise                   # scheduler.add_noise (line 429)
            noisy = torch.cat([first, rest_noisy], dim=3).reshape(b * n, c, f, h, w)
            cam = get_camera_cond(camera.reshape(b, n, f, 4, 4)[:, :, 0].reshape(b * n, 4, 4))
            ts = t[:, None].repeat(1, n).reshape(-1)
            eps2 = self.forward_unet(torch.cat([noisy, noisy]), torch.cat([ts, ts]).float(), text_embeddings, torch.cat([cam, cam]),
                                     torch.cat([image_embeds, torch.zeros_like(image_embeds)]), cfg.i2v_cond_time_zero)
            e_text, e_unc = eps2.chunk(2)                                           # (cond, uncond): line 452
            to_img = lambda x: x.permute(0, 2, 1, 3, 4).reshape(b * n * f, c, h, w)  # "b c f h w -> (b f) c h w"
            e_text, e_unc = to_img(e_text), to_img(e_unc)
            eps = e_text + cfg.guidance_scale * (e_text - e_unc)                    # line 457
            noisy_img = to_img(noisy)
            a_img = self.alphas[t].repeat_interleave(n * f).reshape(-1, 1, 1, 1)
            x0 = (noisy_img - (1 - a_img).sqrt() * eps) / a_img.sqrt()              # pred_original_sample (466)
            if cfg.recon_std_rescale > 0:
                x0_nocfg = (noisy_img - (1 - a_img).sqrt() * e_text) / a_img.sqrt()
                r = lambda x: x.reshape(b, n, f, c, h, w)[:, :, 1:]
                factor = (r(x0_nocfg).std([1, 2, 3, 4, 5], keepdim=True) + 1e-8) / (r(x0).std([1, 2, 3, 4, 5], keepdim=True) + 1e-8)
                adj = x0 * factor.reshape(b, 1, 1, 1).repeat_interleave(n * f, dim=0)
                x0 = cfg.recon_std_rescale * adj + (1 - cfg.recon_std_rescale) * x0
            x0 = x0.reshape(b * n, f, c, h, w)
            x0 = torch.cat([latents.detach().reshape(b * n, f, c, h, w)[:, 0:1], x0[:, 1:]], dim=1).reshape(bnf, c, h, w)
        loss = 0.5 * F.mse_loss(latents, x0, reduction="sum") / latents.shape[0] * f / (f - 1)     # 497-501
        return loss, {"latents_noisy": noisy_img, "noise_pred": eps, "latents_recon": x0, "t_orig": t}

    def __call__(self, rgb, text_embeddings, c2w, image_embeds=None, rgb_as_latents=False, timestep=None, **unused):
        """animatemv_guidance.py:515-600.  rgb [B,H,W,3] in [0,1] (or latents [B,4,32,32] when rgb_as_latents)."""
        cfg = self.cfg
        if rgb_as_latents:
            latents = rgb
        else:
            if self.encode_images is None:
                raise ValueError("inject a VAE encoder callable (SURVEY 8(f) next row) or pass rgb_as_latents=True")
            x = F.interpolate(rgb.permute(0, 3, 1, 2), (256, 256), mode="bilinear", align_corners=False)
            latents = self.encode_images(x)
        b = latents.shape[0] // (cfg.n_view * cfg.n_frame)
        if image_embeds is None:
            if self.encode_ip_image is None:
                raise ValueError("inject a CLIP image encoder callable or pass image_embeds")
            image_embeds = self.encode_ip_image(rgb.reshape(b, cfg.n_view, cfg.n_frame, *rgb.shape[1:])[:, :, 0])
        t = timestep if timestep is not None else torch.randint(self.min_step, self.max_step + 1, [b], dtype=torch.long,
                                                                device=latents.device)
        loss, aux = self.compute_mvdream_recon_loss(latents, t, text_embeddings, c2w, image_embeds)
        return {"loss_sds": loss, "min_step": self.min_step, "max_step": self.max_step, **aux}
