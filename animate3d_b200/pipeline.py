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


def tensor2vid(video: torch.Tensor, processor=None, output_type: str = "np"):
    """pipeline.py:237-255: [B, C, F, H, W] in [-1, 1] -> per-batch frame lists.  `processor` (diffusers VaeImageProcessor) is
    optional: its postprocess for these three output types is (x / 2 + 0.5).clamp(0, 1) -> NHWC (-> uint8 PIL)."""
    import numpy as np
    outputs = []
    for b in range(video.shape[0]):
        frames = video[b].permute(1, 0, 2, 3)                                 # [F, C, H, W]
        if processor is not None:
            outputs.append(processor.postprocess(frames, output_type))
            continue
        img = (frames / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            outputs.append(img)
        else:
            arr = img.detach().cpu().permute(0, 2, 3, 1).float().numpy()
            if output_type == "np":
                outputs.append(arr)
            elif output_type == "pil":
                from PIL import Image
                outputs.append([Image.fromarray((a * 255).round().astype("uint8")) for a in arr])
            else:
                raise ValueError(f"{output_type} does not exist. Please choose one of ['np', 'pt', 'pil']")
    if output_type == "np":
        outputs = np.stack(outputs)
    elif output_type == "pt":
        outputs = torch.stack(outputs)
    return outputs


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

    # -------------------------------------------------------------------------------------------- conditioning encoders
    # The CLIP text / image towers and the VAE are the reference's own third-party modules (transformers CLIPTextModel /
    # CLIPVisionModelWithProjection, diffusers AutoencoderKL): they are injected, not re-implemented -- transformers is a
    # library here exactly as it is in the reference.  Callables are accepted too (round-1 behaviour).
    def encode_prompt(self, prompt, device=None, num_images_per_prompt: int = 1, do_classifier_free_guidance: bool = True,
                      negative_prompt=None, prompt_embeds=None, negative_prompt_embeds=None, lora_scale=None, clip_skip=None):
        """pipeline.py:345-512 without the LoRA / textual-inversion branches: tokenize (max_length, truncation), text encoder
        (optionally the clip_skip hidden state through final_layer_norm), repeat per view, same for the negative prompt."""
        device = device or self.device
        if prompt_embeds is None:
            if self.text_encoder is None:
                raise ValueError("pass prompt_embeds or construct the pipeline with tokenizer + text_encoder")
            if self.tokenizer is None:                                     # plain callable: (prompt, n) -> embeddings
                return self.text_encoder(prompt, num_images_per_prompt), self.text_encoder(negative_prompt or "", num_images_per_prompt)
            prompt_embeds = self._encode_text([prompt] if isinstance(prompt, str) else list(prompt), device, clip_skip)
        bs, seq, _ = prompt_embeds.shape
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(bs * num_images_per_prompt, seq, -1)
        if do_classifier_free_guidance and negative_prompt_embeds is None:
            neg = [""] * bs if negative_prompt is None else ([negative_prompt] * bs if isinstance(negative_prompt, str) else list(negative_prompt))
            if len(neg) != bs:
                raise ValueError(f"`negative_prompt` has batch size {len(neg)}, but `prompt` has batch size {bs}")
            negative_prompt_embeds = self._encode_text(neg, device, None)
        if do_classifier_free_guidance:
            negative_prompt_embeds = negative_prompt_embeds.to(prompt_embeds.dtype).repeat(1, num_images_per_prompt, 1).view(
                bs * num_images_per_prompt, seq, -1)
        return prompt_embeds, negative_prompt_embeds

    def _encode_text(self, texts, device, clip_skip):
        tok = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length, truncation=True, return_tensors="pt")
        ids = tok.input_ids.to(device)
        if clip_skip is None:
            return self.text_encoder(ids)[0]
        out = self.text_encoder(ids, output_hidden_states=True)
        return self.text_encoder.text_model.final_layer_norm(out[-1][-(clip_skip + 1)])

    def encode_image(self, image, device=None):
        """pipeline.py:514-526: CLIP image embeds of the condition view(s) + their all-zero unconditional twin."""
        device = device or self.device
        if self.image_encoder is None:
            raise ValueError("construct the pipeline with image_encoder (+ feature_extractor) or pass ip_adapter_image_embeds")
        if not hasattr(self.image_encoder, "parameters"):                 # plain callable
            emb = self.image_encoder(image)
            return emb, torch.zeros_like(emb)
        dtype = next(self.image_encoder.parameters()).dtype
        if not isinstance(image, torch.Tensor):
            image = self.feature_extractor(image, return_tensors="pt").pixel_values
        emb = self.image_encoder(image.to(device=device, dtype=dtype)).image_embeds
        return emb, torch.zeros_like(emb)

    def encode_latents(self, image_size, image_list):
        """pipeline.py:528-551: PIL condition images -> VAE latents * scaling_factor (resize to image_size[0], normalise to [-1,1])."""
        import numpy as np
        if self.vae is None:
            raise ValueError("construct the pipeline with a VAE or pass first_frame_latents")
        x = torch.stack([torch.from_numpy(np.array(im)) for im in image_list], 0).float() / 255
        x = x.permute(0, 3, 1, 2)
        x = torch.nn.functional.interpolate(x, size=(image_size[0], image_size[0] * x.shape[3] // x.shape[2]) if x.shape[2] <= x.shape[3]
                                            else (image_size[0] * x.shape[2] // x.shape[3], image_size[0]), mode="bilinear",
                                            antialias=True, align_corners=False)
        x = (x - 0.5) / 0.5
        with torch.no_grad():
            lat = self.vae.encode(x.to(self.device)).latent_dist.sample()
        return lat * self.vae.config.scaling_factor

    def decode_latents(self, latents: torch.Tensor) -> torch.Tensor:
        """pipeline.py:554-567: [B, 4, F, h, w] -> video [B, 3, F, 8h, 8w] in [-1, 1] (fp32)."""
        if self.vae is None:
            raise ValueError("construct the pipeline with a VAE to decode latents")
        latents = latents / self.vae.config.scaling_factor
        b, c, f, h, w = latents.shape
        image = self.vae.decode(latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)).sample
        return image[None].reshape((b, f, -1) + image.shape[2:]).permute(0, 2, 1, 3, 4).float()

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
                 ip_adapter_image_embeds: Optional[torch.Tensor] = None, output_type: str = "pil", return_dict: bool = True,
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
        if prompt_embeds is None or negative_prompt_embeds is None:
            prompt_embeds, negative_prompt_embeds = self.encode_prompt(prompt, dev, nv, True, negative_prompt, prompt_embeds=prompt_embeds,
                                                                       negative_prompt_embeds=negative_prompt_embeds, clip_skip=clip_skip)
        if ip_adapter_image_embeds is None:
            ip_adapter_image_embeds, _ = self.encode_image(ip_adapter_image, dev)
        if first_frame_latents is None:
            first_frame_latents = self.encode_latents((height, width), ip_adapter_image)
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
        if output_type == "latent":
            video = lat
        else:
            if self.vae is None:
                raise ValueError('output_type "pil" / "np" / "pt" needs a VAE; pass output_type="latent" to get the latents '
                                 "(pipeline.py:1049-1056)")
            video = tensor2vid(self.decode_latents(lat), getattr(self, "image_processor", None), output_type=output_type)
        return AnimateDiffMVI2VPipelineOutput(frames=video) if return_dict else (video,)


AnimationPipeline = AnimateDiffMVI2VPipeline   # name used by BASELINE.json's north_star
