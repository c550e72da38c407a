"""ORACLE (test infrastructure) -- fp32 torch restatement of the SD-1.5 `AutoencoderKL` the reference loads with
`AutoencoderKL.from_pretrained(..., subfolder="vae")` (custom/threestudio-animate3d/guidance/animatemv_guidance.py:123) and calls
at 365-389 (`encode_images` / `decode_latents`) and animatediff/pipelines/pipeline.py:528-567.

PARITY UNPINNED: diffusers 0.28.0 is not installed and not vendored; restated from its published
`models/autoencoders/{autoencoder_kl.py, vae.py}`, `unets/unet_2d_blocks.py` ({Down,Up}EncoderBlock2D, UNetMidBlock2D),
`resnet.py` (ResnetBlock2D with temb None, Downsample2D(padding=0) = F.pad (0,1,0,1) + stride-2 conv, Upsample2D = nearest x2 +
conv) and `attention_processor.py` (single-head `Attention` with group_norm, residual_connection=True).  Key names are
diffusers' so a released `vae/diffusion_pytorch_model.safetensors` state dict loads unchanged."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class VAEConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-6
    scaling_factor: float = 0.18215


def _res_keys(p, cin, cout):
    ks = {f"{p}.norm1.weight": (cin,), f"{p}.norm1.bias": (cin,), f"{p}.conv1.weight": (cout, cin, 3, 3), f"{p}.conv1.bias": (cout,),
          f"{p}.norm2.weight": (cout,), f"{p}.norm2.bias": (cout,), f"{p}.conv2.weight": (cout, cout, 3, 3), f"{p}.conv2.bias": (cout,)}
    if cin != cout:
        ks[f"{p}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        ks[f"{p}.conv_shortcut.bias"] = (cout,)
    return ks


def _mid_keys(p, c):
    ks = {}
    ks.update(_res_keys(f"{p}.resnets.0", c, c))
    a = f"{p}.attentions.0"
    ks[f"{a}.group_norm.weight"] = (c,)
    ks[f"{a}.group_norm.bias"] = (c,)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        ks[f"{a}.{n}.weight"] = (c, c)
        ks[f"{a}.{n}.bias"] = (c,)
    ks.update(_res_keys(f"{p}.resnets.1", c, c))
    return ks


def key_plan(cfg: VAEConfig) -> Dict[str, Tuple[int, ...]]:
    ch = cfg.block_out_channels
    ks: Dict[str, Tuple[int, ...]] = {"encoder.conv_in.weight": (ch[0], cfg.in_channels, 3, 3), "encoder.conv_in.bias": (ch[0],)}
    cout = ch[0]
    for i, c in enumerate(ch):
        cin, cout = cout, c
        for j in range(cfg.layers_per_block):
            ks.update(_res_keys(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout))
        if i != len(ch) - 1:
            ks[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            ks[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
    ks.update(_mid_keys("encoder.mid_block", ch[-1]))
    ks.update({"encoder.conv_norm_out.weight": (ch[-1],), "encoder.conv_norm_out.bias": (ch[-1],),
               "encoder.conv_out.weight": (2 * cfg.latent_channels, ch[-1], 3, 3), "encoder.conv_out.bias": (2 * cfg.latent_channels,),
               "quant_conv.weight": (2 * cfg.latent_channels, 2 * cfg.latent_channels, 1, 1), "quant_conv.bias": (2 * cfg.latent_channels,),
               "post_quant_conv.weight": (cfg.latent_channels, cfg.latent_channels, 1, 1), "post_quant_conv.bias": (cfg.latent_channels,),
               "decoder.conv_in.weight": (ch[-1], cfg.latent_channels, 3, 3), "decoder.conv_in.bias": (ch[-1],)})
    ks.update(_mid_keys("decoder.mid_block", ch[-1]))
    rev = list(reversed(ch))
    cout = rev[0]
    for i, c in enumerate(rev):
        cin, cout = cout, c
        for j in range(cfg.layers_per_block + 1):
            ks.update(_res_keys(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout))
        if i != len(ch) - 1:
            ks[f"decoder.up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            ks[f"decoder.up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    ks.update({"decoder.conv_norm_out.weight": (ch[0],), "decoder.conv_norm_out.bias": (ch[0],),
               "decoder.conv_out.weight": (cfg.out_channels, ch[0], 3, 3), "decoder.conv_out.bias": (cfg.out_channels,)})
    return ks


def make_state_dict(cfg: VAEConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Seeded random weights with activations of O(1) through the depth: conv / linear weights N(0, 1/fan_in), norm affine
    (1 + 0.1 n, 0.1 n), small biases."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in key_plan(cfg).items():
        if "norm" in k.split(".")[-2]:
            n = torch.randn(shape, generator=g) * 0.1
            sd[k] = 1.0 + n if k.endswith("weight") else n
        elif k.endswith("bias"):
            sd[k] = torch.randn(shape, generator=g) * 0.02
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[k] = torch.randn(shape, generator=g) * fan_in ** -0.5
    return sd


def resnet(sd, p, x, cfg: VAEConfig):
    h = F.silu(F.group_norm(x, cfg.norm_num_groups, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], cfg.norm_eps))
    h = F.conv2d(h, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    h = F.silu(F.group_norm(h, cfg.norm_num_groups, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], cfg.norm_eps))
    h = F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h                                                  # output_scale_factor 1


def mid_attention(sd, p, x, cfg: VAEConfig):
    """Attention(heads = 1, dim_head = C, norm_num_groups, residual_connection=True, bias=True)."""
    n, c, h, w = x.shape
    t = F.group_norm(x, cfg.norm_num_groups, sd[f"{p}.group_norm.weight"], sd[f"{p}.group_norm.bias"], cfg.norm_eps)
    t = t.reshape(n, c, h * w).transpose(1, 2)
    q = F.linear(t, sd[f"{p}.to_q.weight"], sd[f"{p}.to_q.bias"])
    k = F.linear(t, sd[f"{p}.to_k.weight"], sd[f"{p}.to_k.bias"])
    v = F.linear(t, sd[f"{p}.to_v.weight"], sd[f"{p}.to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) * c ** -0.5, dim=-1) @ v
    a = F.linear(a, sd[f"{p}.to_out.0.weight"], sd[f"{p}.to_out.0.bias"])
    return x + a.transpose(1, 2).reshape(n, c, h, w)


def mid_block(sd, p, x, cfg):
    x = resnet(sd, f"{p}.resnets.0", x, cfg)
    x = mid_attention(sd, f"{p}.attentions.0", x, cfg)
    return resnet(sd, f"{p}.resnets.1", x, cfg)


def encode_moments(sd, cfg: VAEConfig, x: Tensor) -> Tensor:
    """x [N,3,H,W] in [-1,1] -> moments [N, 8, H/8, W/8] (mean | logvar) = quant_conv(encoder(x))."""
    ch = cfg.block_out_channels
    h = F.conv2d(x, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block):
            h = resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", h, cfg)
        if i != len(ch) - 1:
            p = f"encoder.down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"{p}.weight"], sd[f"{p}.bias"], stride=2)
    h = mid_block(sd, "encoder.mid_block", h, cfg)
    h = F.silu(F.group_norm(h, cfg.norm_num_groups, sd["encoder.conv_norm_out.weight"], sd["encoder.conv_norm_out.bias"], cfg.norm_eps))
    h = F.conv2d(h, sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    return F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])


def sample_latents(moments: Tensor, noise: Tensor) -> Tensor:
    """DiagonalGaussianDistribution.sample: mean + exp(0.5 * clamp(logvar, -30, 20)) * noise."""
    mean, logvar = moments.chunk(2, dim=1)
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


def decode(sd, cfg: VAEConfig, z: Tensor) -> Tensor:
    """z [N,4,h,w] (already divided by scaling_factor) -> image [N,3,8h,8w]."""
    ch = cfg.block_out_channels
    h = F.conv2d(z, sd["post_quant_conv.weight"], sd["post_quant_conv.bias"])
    h = F.conv2d(h, sd["decoder.conv_in.weight"], sd["decoder.conv_in.bias"], padding=1)
    h = mid_block(sd, "decoder.mid_block", h, cfg)
    for i in range(len(ch)):
        for j in range(cfg.layers_per_block + 1):
            h = resnet(sd, f"decoder.up_blocks.{i}.resnets.{j}", h, cfg)
        if i != len(ch) - 1:
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = F.conv2d(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd[f"{p}.weight"], sd[f"{p}.bias"], padding=1)
    h = F.silu(F.group_norm(h, cfg.norm_num_groups, sd["decoder.conv_norm_out.weight"], sd["decoder.conv_norm_out.bias"], cfg.norm_eps))
    return F.conv2d(h, sd["decoder.conv_out.weight"], sd["decoder.conv_out.bias"], padding=1)
