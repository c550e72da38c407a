"""Host-side mirrors of the sampler around the UNet (no GPU): DDIM constants / time steps, camera conditioning, FreeInit
low-pass mixing -- product code (`animate3d_b200/scheduler.py`, `pipeline.py`) against the oracle and the reference goldens."""
import math
import os

import numpy as np
import pytest
import torch


def test_ddim_constants_and_timesteps_match_oracle():
    from animate3d_b200.scheduler import DDIMScheduler
    from oracle.scheduler_oracle import DDIMOracle
    s, o = DDIMScheduler(), DDIMOracle()
    np.testing.assert_allclose(s.alphas_cumprod, o.alphas_cumprod.numpy(), rtol=2e-6)
    for n in (25, 50, 4):
        ts, to = s.set_timesteps(n), o.set_timesteps(n)
        assert ts.tolist() == to.tolist() and ts[0] == (n - 1) * (1000 // n) + 1 and ts[-1] == 1
        for t in ts.tolist():
            a_t, a_p = s.alphas_for(t)
            prev = t - 1000 // n
            assert a_t == pytest.approx(float(o.alphas_cumprod[t]), rel=2e-6)
            assert a_p == pytest.approx(float(o.alphas_cumprod[prev]) if prev >= 0 else 1.0, rel=2e-6)
    with pytest.raises(NotImplementedError):
        DDIMScheduler(beta_schedule="scaled_linear")


def test_ddim_update_formula_is_the_oracle_step():
    """The fused kernel's formula (x0 prediction, eta = 0) written out on the host with the scheduler's constants."""
    from animate3d_b200.scheduler import DDIMScheduler
    from oracle.scheduler_oracle import DDIMOracle, denoise_step
    s, o = DDIMScheduler(), DDIMOracle()
    s.set_timesteps(25); o.set_timesteps(25)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(4, 4, 6, 8, 8, generator=g)
    eps2 = torch.randn(8, 4, 6, 8, 8, generator=g)
    first = torch.randn(4, 4, 1, 8, 8, generator=g)
    t = int(s.timesteps[3])
    a_t, a_p = s.alphas_for(t)
    eps = eps2[:4] + 7.5 * (eps2[4:] - eps2[:4])                       # (uncond, cond) order, pipeline.py:1023-1025
    x0 = (lat - math.sqrt(1 - a_t) * eps) / math.sqrt(a_t)
    mine = math.sqrt(a_p) * x0 + math.sqrt(1 - a_p) * eps
    mine[:, :, :1] = first                                             # frame-0 re-injection, pipeline.py:1029-1031
    torch.testing.assert_close(mine, denoise_step(lat, eps2, first, 7.5, o, t), rtol=1e-5, atol=1e-5)


def test_get_camera_matches_reference(golden_dir):
    from animate3d_b200.pipeline import get_camera
    ref = torch.load(os.path.join(golden_dir, "ref_camera.pt"), weights_only=False)
    for n, want in ref.items():
        torch.testing.assert_close(get_camera(n), want.float().reshape(n, 16), rtol=1e-5, atol=1e-6)


def test_freeinit_mix_matches_oracle():
    from animate3d_b200.pipeline import AnimateDiffMVI2VPipeline, _butterworth_lpf
    from animate3d_b200.scheduler import DDIMScheduler
    from oracle.scheduler_oracle import butterworth_lpf, freeinit_mix
    shape = (2, 4, 8, 6, 6)
    torch.testing.assert_close(_butterworth_lpf(shape).expand(shape), butterworth_lpf(shape))
    pipe = AnimateDiffMVI2VPipeline.__new__(AnimateDiffMVI2VPipeline)   # host logic only: no UNet, no device
    pipe.scheduler = DDIMScheduler()
    pipe._fi = (4, 0.25, 0.25)
    g = torch.Generator().manual_seed(3)
    x0 = torch.randn(shape, generator=g)
    lat0, ts = pipe._apply_free_init(x0, 0, 25, g)
    assert lat0 is x0 and len(ts) == 25
    x1 = torch.randn(shape, generator=g)
    g2 = torch.Generator().manual_seed(11)
    lat1, _ = pipe._apply_free_init(x1, 1, 25, g2)
    a = float(pipe.scheduler.alphas_cumprod[999])
    z_T = math.sqrt(a) * x1 + math.sqrt(1 - a) * x0
    z_rand = torch.randn(shape, generator=torch.Generator().manual_seed(11))
    torch.testing.assert_close(lat1, freeinit_mix(z_T, z_rand, butterworth_lpf(shape)), rtol=1e-5, atol=1e-5)


class _StubUNet:
    device = torch.device("cpu")
    config = type("C", (), {"in_channels": 4})()


def test_pipeline_conditioning_encoders_with_transformers_clip():
    """encode_prompt / encode_image (pipeline.py:345-526) over the reference's own encoder classes (transformers CLIP, random
    init, tiny geometry): shapes, per-view repetition, CFG negatives, clip_skip, zero unconditional image embeds."""
    from types import SimpleNamespace
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPVisionConfig, CLIPVisionModelWithProjection
    from animate3d_b200.pipeline import AnimateDiffMVI2VPipeline
    torch.manual_seed(0)
    text = CLIPTextModel(CLIPTextConfig(vocab_size=100, hidden_size=32, intermediate_size=64, num_hidden_layers=3, num_attention_heads=4,
                                        max_position_embeddings=77)).eval()
    vision = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=4,
                                                            image_size=32, patch_size=8, projection_dim=24)).eval()

    class Tok:
        model_max_length = 77

        def __call__(self, texts, padding=None, max_length=None, truncation=None, return_tensors=None):
            ids = torch.zeros(len(texts), max_length, dtype=torch.long)
            for i, t in enumerate(texts):
                w = [1] + [2 + (hash(x) % 90) for x in t.split()][: max_length - 2] + [99]
                ids[i, : len(w)] = torch.tensor(w)
            return SimpleNamespace(input_ids=ids)

    pipe = AnimateDiffMVI2VPipeline(unet=_StubUNet(), text_encoder=text, tokenizer=Tok(), image_encoder=vision)
    with torch.no_grad():
        pe, ne = pipe.encode_prompt("a dragon head", "cpu", num_images_per_prompt=4, negative_prompt="blurry")
        assert pe.shape == (4, 77, 32) and ne.shape == (4, 77, 32)
        assert torch.equal(pe[0], pe[3]) and not torch.allclose(pe[0], ne[0])
        pe2, _ = pipe.encode_prompt("a dragon head", "cpu", num_images_per_prompt=1, clip_skip=1)
        assert pe2.shape == (1, 77, 32) and not torch.allclose(pe2[0], pe[0])
        emb, unc = pipe.encode_image(torch.randn(4, 3, 32, 32), "cpu")
        assert emb.shape == (4, 24) and unc.abs().sum() == 0
    with pytest.raises(ValueError):
        AnimateDiffMVI2VPipeline(unet=_StubUNet()).encode_prompt("x", "cpu")


def test_decode_latents_and_tensor2vid():
    """pipeline.py:554-567 and 237-255 with a stand-in VAE: frame/batch regrouping and the three output types."""
    from types import SimpleNamespace
    from animate3d_b200.pipeline import AnimateDiffMVI2VPipeline, tensor2vid

    class VAE:
        config = SimpleNamespace(scaling_factor=0.5)

        def decode(self, z):                       # [N,4,h,w] -> [N,3,8h,8w]: channel means upsampled, tagged by the input sum
            up = torch.nn.functional.interpolate(z[:, :3], scale_factor=8.0, mode="nearest")
            return SimpleNamespace(sample=up)

    pipe = AnimateDiffMVI2VPipeline(unet=_StubUNet(), vae=VAE())
    lat = torch.arange(2 * 4 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 4, 3, 2, 2) / 100
    vid = pipe.decode_latents(lat)
    assert vid.shape == (2, 3, 3, 16, 16)
    torch.testing.assert_close(vid[1, 2, 1, 0, 0], lat[1, 2, 1, 0, 0] / 0.5)         # (batch 1, channel 2, frame 1) stays in place
    v = torch.rand(2, 3, 5, 8, 8) * 2 - 1
    assert tensor2vid(v, output_type="np").shape == (2, 5, 8, 8, 3)
    assert tensor2vid(v, output_type="pt").shape == (2, 5, 3, 8, 8)
    pil = tensor2vid(v, output_type="pil")
    assert len(pil) == 2 and len(pil[0]) == 5 and pil[0][0].size == (8, 8)
    with pytest.raises(ValueError):
        tensor2vid(v, output_type="gif")
