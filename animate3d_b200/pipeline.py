"""Host-side mirror of the reference's sampling pipeline `AnimateDiffMVI2VPipeline` (animatediff/pipelines/pipeline.py:274-1062)
for the denoising hot loop: FreeInit outer loop (987-999) x DDIM loop (1006-1047) with classifier-free guidance (1008,
1023-1025), scheduler step (1028) and first-frame re-injection (1031).

The UNet evaluation, CFG combine, DDIM update and frame-0 re-injection all run as CUDA kernels of liba3d.so.  The VAE and
the CLIP text / image encoders ("next" rows) are injected as opaque callables or bypassed with
pre-computed embeddings / latents, exactly as the reference's own `prompt_embeds`, `ip_adapter_image_embeds` and
`latents` arguments allow (pipeline.py:770-778)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional

import torch

from . import ops
from .scheduler import DDIMScheduler
from .unet import MVUNetMotionModel


@dataclass
class AnimateDiffMVI2VPipelineOutput:
    frames: object


def get_camera(num_views: int, elevation: float = 15.0, azimuth_start: float = 0.0, azimuth_span: float = 360.0) -> torch.Tensor:
    """Camera conditioning of pipeline.py:127-190: c2w of `num_views` cameras on a circle at `elevation`, translation
    normalised to the unit sphere, flattened to [num_views, 16]."""
    out = []
    gap = azimuth_span / num_views
    for i in range(num_views):
        e = math.radians(elevation)
        a = math.radians(azimuth_start + i * gap)
        pos = torch.tensor([math.cos(e) * math.cos(a), math.cos(e) * math.sin(a), math.sin(e)], dtype=torch.float32)
        look = -pos / pos.norm()
        right = torch.linalg.cross(look, torch.tensor([0.0, 0.0, 1.0]))
        right = right / right.norm()
        up = torch.linalg.cross(right, look)
        up = up / up.norm()
        c2w = torch.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2] = right, up, -look
        c2w[:3, 3] = pos / (pos.norm() + 1e-8)
        out.append(c2w.flatten())
    return torch.stack(out, 0)


def _butterworth_lpf(shape, order=4, d_s=0.25, d_t=0.25, device="cpu"):
    T, H, W = shape[-3], shape[-2], shape[-1]
    t = torch.arange(T, device=device)[:, None, None].float()
    h = torch.arange(H, device=device)[None, :, None].float()
    w = torch.arange(W, device=device)[None, None, :].float()
    d2 = ((d_s / d_t) * (2 * t / T - 1)) ** 2 + (2 * h / H - 1) ** 2 + (2 * w / W - 1) ** 2
    return 1.0 / (1.0 + (d2 / d_s ** 2) ** order)


class AnimateDiffMVI2VPipeline:
    """`AnimationPipeline` of north_star == this class (alias below).  Constructor keeps the reference's argument names
    (pipeline.py:308-325); everything except `unet` and `scheduler` is optional."""

    def __init__(self, vae=None, text_encoder=None, tokenizer=None, unet: MVUNetMotionModel = None, motion_adapter=None,
                 scheduler: DDIMScheduler = None, feature_extractor=None, image_encoder=None):
        if unet is None:
            raise ValueError("unet is required")
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.scheduler = unet, scheduler or DDIMScheduler()
        self.feature_extractor, self.image_encoder = feature_extractor, image_encoder
        self.free_init_enabled = False
        self._free_init_num_iters = 1
        self.device = unet.device

    def to(self, device):
        return self

    def enable_free_init(self, num_iters: int = 3, use_fast_sampling: bool = False, method: str = "butterworth", order: int = 4,
                         spatial_stop_frequency: float = 0.25, temporal_stop_frequency: float = 0.25):
        if method != "butterworth" or use_fast_sampling:
            raise NotImplementedError("released configuration only: butterworth, no fast sampling (inference.py:244-245)")
        self.free_init_enabled, self._free_init_num_iters = True, num_iters
        self._fi = (order, spatial_stop_frequency, temporal_stop_frequency)

    def disable_free_init(self):
        self.free_init_enabled = False

    def enable_vae_slicing(self):
        pass

    # -------------------------------------------------------------------------------------------- one denoise step
    def denoise_step(self, latents: torch.Tensor, t: int, prompt_embeds: torch.Tensor, camera: torch.Tensor,
                     image_embeds: torch.Tensor, first_frame_latents: torch.Tensor, guidance_scale: float,
                     i2v_cond_time_zero: bool = False, num_views: int = 4) -> torch.Tensor:
        """Body of the loop at pipeline.py:1006-1031, in place on `latents` [Nv, 4, F, h, w] (fp32, device).
        prompt_embeds / camera / image_embeds already carry the (uncond, cond) CFG duplication."""
        x2 = torch.cat([latents, latents], 0)
        noise_pred = self.unet(x2, t, prompt_embeds, camera=camera, added_cond_kwargs={"image_embeds": image_embeds},
                               num_views=num_views, i2v_cond_time_zero=i2v_cond_time_zero).sample
        a_t, a_p = self.scheduler.alphas_for(int(t))
        bn, c, f, h, w = latents.shape
        ops.ddim_cfg_step(latents, noise_pred.contiguous(), first_frame_latents.contiguous(), bn, c, f, h * w, guidance_scale,
                          a_t, a_p, uncond_first=True)
        return latents

    def denoise_step_host(self, latents_host: torch.Tensor, t: int, prompt_embeds_host, camera_host, image_embeds_host,
                          first_frame_host, guidance_scale: float, out_host: Optional[torch.Tensor] = None,
                          num_views: int = 4) -> torch.Tensor:
        """Same step through HOST (pinned) buffers: copies the step's inputs to the device, runs it, copies the updated
        latents back -- what an external caller holding numpy/CPU tensors pays (bench.py's e2e leg)."""
        dev = self.device
        lat = latents_host.to(dev, non_blocking=True)
        pe = prompt_embeds_host.to(dev, non_blocking=True)
        cam = camera_host.to(dev, non_blocking=True)
        ie = image_embeds_host.to(dev, non_blocking=True)
        ff = first_frame_host.to(dev, non_blocking=True)
        self.denoise_step(lat, t, pe, cam, ie, ff, guidance_scale, num_views=num_views)
        if out_host is None:
            out_host = torch.empty_like(latents_host).pin_memory()
        out_host.copy_(lat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def denoise_step_sharded(self, latents_local: torch.Tensor, t: int, prompt_embeds_local: torch.Tensor,
                             camera_local: torch.Tensor, image_embeds_local: torch.Tensor, first_frame_local: torch.Tensor,
                             guidance_scale: float, cfg_group=None, num_views_local: int = 4,
                             i2v_cond_time_zero: bool = False) -> torch.Tensor:
        """One denoise step of ONE prompt spread over ranks: this rank owns `num_views_local` views (all of them,
        or one when `self.unet.view_group` shards the views) and either both CFG branches (cfg_group None: inputs carry the
        (uncond, cond) duplication) or one of them (cfg_group of size 2: rank 0 = uncond, rank 1 = cond).  Collectives: the
        K|V all-gathers inside the UNet (view group) and ONE all-gather of the noise prediction over the CFG group
        (n_local x 4 x F x h x w fp32 = 0.25 MB per view).  The latents stay sharded by view across steps."""
        nloc = latents_local.shape[0]
        if cfg_group is None:
            x = torch.cat([latents_local, latents_local], 0)
        else:
            x = latents_local
        eps = self.unet(x, t, prompt_embeds_local, camera=camera_local, added_cond_kwargs={"image_embeds": image_embeds_local},
                        num_views=num_views_local, i2v_cond_time_zero=i2v_cond_time_zero).sample
        if cfg_group is not None:
            import torch.distributed as dist
            eps2 = self._eps2 if getattr(self, "_eps2", None) is not None and self._eps2.shape[0] == 2 * nloc else None
            if eps2 is None:
                eps2 = self._eps2 = torch.empty(2 * nloc, *eps.shape[1:], device=eps.device, dtype=eps.dtype)
            dist.all_gather_into_tensor(eps2.view(-1), eps.contiguous().view(-1), group=cfg_group)
            eps = eps2
        a_t, a_p = self.scheduler.alphas_for(int(t))
        _, c, f, h, w = latents_local.shape
        ops.ddim_cfg_step(latents_local, eps.contiguous(), first_frame_local.contiguous(), nloc, c, f, h * w, guidance_scale, a_t,
                          a_p, uncond_first=True)
        return latents_local

    # -------------------------------------------------------------------------------------------- full sampler
    def _apply_free_init(self, latents, it, num_inference_steps, generator):
        """FreeInitMixin._apply_free_init on frames 1.. (pipeline.py:990-992)."""
        if it == 0:
            self._free_init_initial_noise = latents.detach().clone()
        else:
            order, ds, dt = self._fi
            a = float(self.scheduler.alphas_cumprod[self.scheduler.num_train_timesteps - 1])
            z_T = math.sqrt(a) * latents + math.sqrt(1 - a) * self._free_init_initial_noise
            z_rand = torch.randn(latents.shape, generator=generator, device=latents.device, dtype=torch.float32)
            lpf = _butterworth_lpf(latents.shape, order, ds, dt, latents.device)
            dims = (-3, -2, -1)
            zf = torch.fft.fftshift(torch.fft.fftn(z_T, dim=dims), dim=dims)
            rf = torch.fft.fftshift(torch.fft.fftn(z_rand, dim=dims), dim=dims)
            latents = torch.fft.ifftn(torch.fft.ifftshift(zf * lpf + rf * (1 - lpf), dim=dims), dim=dims).real.contiguous()
        self.scheduler.set_timesteps(num_inference_steps)
        return latents, self.scheduler.timesteps

    @torch.no_grad()
    def __call__(self, prompt=None, num_frames: int = 16, height: int = 256, width: int = 256, num_inference_steps: int = 50,
                 guidance_scale: float = 7.5, negative_prompt=None, num_videos_per_prompt: int = 1, eta: float = 0.0,
                 generator=None, latents: Optional[torch.Tensor] = None, prompt_embeds: Optional[torch.Tensor] = None,
                 negative_prompt_embeds: Optional[torch.Tensor] = None, ip_adapter_image=None,
                 ip_adapter_image_embeds: Optional[torch.Tensor] = None, output_type: str = "latent", return_dict: bool = True,
                 cross_attention_kwargs=None, clip_skip=None, callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ("latents",), i2v_cond_time_zero: bool = False,
                 i2v_similarity_init=None, first_frame_latents: Optional[torch.Tensor] = None):
        """Argument names follow pipeline.py:760-786.  `num_videos_per_prompt` is the number of views.  Text / image
        encoders are only invoked if they were injected; otherwise prompt_embeds / negative_prompt_embeds [Nv,77,768],
        ip_adapter_image_embeds [Nv,1024] and first_frame_latents [Nv,4,1,h,w] must be given."""
        if eta != 0.0 or i2v_similarity_init is not None:
            raise NotImplementedError("eta != 0 / i2v_similarity_init are unused by the released configuration")
        dev = self.device
        nv = num_videos_per_prompt
        if prompt_embeds is None:
            if self.text_encoder is None:
                raise ValueError("pass prompt_embeds/negative_prompt_embeds or inject a text encoder callable")
            prompt_embeds = self.text_encoder(prompt, nv)
            negative_prompt_embeds = self.text_encoder(negative_prompt or "", nv)
        if ip_adapter_image_embeds is None:
            if self.image_encoder is None:
                raise ValueError("pass ip_adapter_image_embeds or inject an image encoder callable")
            ip_adapter_image_embeds = self.image_encoder(ip_adapter_image)
        if first_frame_latents is None:
            if self.vae is None:
                raise ValueError("pass first_frame_latents or inject a VAE")
            first_frame_latents = self.vae.encode_first_frames(ip_adapter_image, height, width)
        do_cfg = guidance_scale > 1.0
        if not do_cfg:
            raise NotImplementedError("the released sampler always runs with classifier-free guidance (guidance_scale 7.5)")
        pe = torch.cat([negative_prompt_embeds, prompt_embeds]).to(dev, torch.float32)            # (uncond, cond): line 932
        ie = torch.cat([torch.zeros_like(ip_adapter_image_embeds), ip_adapter_image_embeds]).to(dev, torch.float32)  # 537
        first = first_frame_latents.to(dev, torch.float32).reshape(nv, -1, 1, height // 8, width // 8).contiguous()
        c = self.unet.config.in_channels
        if latents is None:
            latents = torch.randn(nv, c, num_frames - 1, height // 8, width // 8, generator=generator, device=dev, dtype=torch.float32)
        rest = latents.to(dev, torch.float32)
        cam = get_camera(nv).to(dev)
        cam2 = torch.cat([cam, cam])
        iters = self._free_init_num_iters if self.free_init_enabled else 1
        lat = torch.cat([first, rest], dim=2).contiguous()
        for it in range(iters):
            if self.free_init_enabled:
                rest, timesteps = self._apply_free_init(lat[:, :, 1:].contiguous(), it, num_inference_steps, generator)
                lat = torch.cat([first, rest], dim=2).contiguous()
            else:
                timesteps = self.scheduler.set_timesteps(num_inference_steps)
            for i, t in enumerate(timesteps):
                self.denoise_step(lat, int(t), pe, cam2, ie, first, guidance_scale, i2v_cond_time_zero, nv)
                if callback_on_step_end is not None:
                    res = callback_on_step_end(self, i, int(t), {"latents": lat})
                    lat = res.pop("latents", lat)
        if output_type == "latent" or self.vae is None:
            video = lat
        else:
            video = self.vae.decode_latents(lat)
        return AnimateDiffMVI2VPipelineOutput(frames=video) if return_dict else (video,)


AnimationPipeline = AnimateDiffMVI2VPipeline   # name used by BASELINE.json's north_star
