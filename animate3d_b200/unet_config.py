"""Geometry and state-dict key plan of the MV motion UNet (SD1.5 + AnimateDiff motion modules + MVDream camera embedding +
IP-Adapter), mirroring MVUNetMotionModel.__init__ (animatediff/models/unet_motion_mv_model.py:67-273 of the reference) and
the processor wiring of inference.py:107-192.  Key names are the reference's (diffusers 0.28.0 naming) so released
checkpoints load unchanged: the released motion-module checkpoint holds exactly the keys containing "motion_modules." or
"i2v." and leaves 726 others missing (inference.py:219-223)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch

Tensor = torch.Tensor


@dataclass
class UNetConfig:
    """Geometry of the SD1.5 MV motion UNet (unet_motion_mv_model.py:67-102 defaults + mvdream-sd1.5 config)."""
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    cross_attention_dim: int = 768
    num_attention_heads: int = 8
    motion_num_attention_heads: int = 8
    motion_max_seq_length: int = 32
    camera_embedding_dim: int = 16
    ip_image_embed_dim: int = 1024
    ip_num_tokens: int = 4
    ip_scale: float = 1.0
    sample_size: int = 32          # latent side for a 256^2 video (inference.py:93)
    num_views: int = 4
    num_frames: int = 16
    # which down blocks carry spatial transformers (CrossAttnDownBlockMotion x3 + DownBlockMotion)
    down_has_attn: Tuple[bool, ...] = (True, True, True, False)

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    def feature_size(self, level: int) -> int:
        return self.sample_size >> level


# --------------------------------------------------------------------------------------------------------------------
# state-dict key plan
# --------------------------------------------------------------------------------------------------------------------

def _resnet_keys(prefix, cin, cout, temb):
    ks = {
        f"{prefix}.norm1.weight": (cin,), f"{prefix}.norm1.bias": (cin,),
        f"{prefix}.conv1.weight": (cout, cin, 3, 3), f"{prefix}.conv1.bias": (cout,),
        f"{prefix}.time_emb_proj.weight": (cout, temb), f"{prefix}.time_emb_proj.bias": (cout,),
        f"{prefix}.norm2.weight": (cout,), f"{prefix}.norm2.bias": (cout,),
        f"{prefix}.conv2.weight": (cout, cout, 3, 3), f"{prefix}.conv2.bias": (cout,),
    }
    if cin != cout:
        ks[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1)
        ks[f"{prefix}.conv_shortcut.bias"] = (cout,)
    return ks


def _attn_keys(prefix, c, kv_dim):
    return {
        f"{prefix}.to_q.weight": (c, c), f"{prefix}.to_k.weight": (c, kv_dim), f"{prefix}.to_v.weight": (c, kv_dim),
        f"{prefix}.to_out.0.weight": (c, c), f"{prefix}.to_out.0.bias": (c,),
    }


def _tblock_keys(prefix, c, cross_dim):
    ks = {}
    for n in ("norm1", "norm2", "norm3"):
        ks[f"{prefix}.{n}.weight"] = (c,)
        ks[f"{prefix}.{n}.bias"] = (c,)
    ks.update(_attn_keys(f"{prefix}.attn1", c, c))
    ks.update(_attn_keys(f"{prefix}.attn2", c, cross_dim))
    ks[f"{prefix}.ff.net.0.proj.weight"] = (8 * c, c)
    ks[f"{prefix}.ff.net.0.proj.bias"] = (8 * c,)
    ks[f"{prefix}.ff.net.2.weight"] = (c, 4 * c)
    ks[f"{prefix}.ff.net.2.bias"] = (c,)
    return ks


def _transformer2d_keys(prefix, c, cfg: UNetConfig):
    ks = {f"{prefix}.norm.weight": (c,), f"{prefix}.norm.bias": (c,),
          f"{prefix}.proj_in.weight": (c, c, 1, 1), f"{prefix}.proj_in.bias": (c,),
          f"{prefix}.proj_out.weight": (c, c, 1, 1), f"{prefix}.proj_out.bias": (c,)}
    tb = f"{prefix}.transformer_blocks.0"
    ks.update(_tblock_keys(tb, c, cfg.cross_attention_dim))
    # MVDreamI2V processor params (attention_processor.py:322-323) live under <attn>.processor.*
    ks[f"{tb}.attn1.processor.to_q_i2v.weight"] = (c, c)
    ks[f"{tb}.attn1.processor.to_out_i2v.weight"] = (c, c)
    ks[f"{tb}.attn1.processor.to_out_i2v.bias"] = (c,)
    # IPAdapter processor params (attention_processor.py:162-167)
    ks[f"{tb}.attn2.processor.to_k_ip.0.weight"] = (c, cfg.cross_attention_dim)
    ks[f"{tb}.attn2.processor.to_v_ip.0.weight"] = (c, cfg.cross_attention_dim)
    return ks


def _motion_keys(prefix, c, cfg: UNetConfig):
    ks = {f"{prefix}.norm.weight": (c,), f"{prefix}.norm.bias": (c,),
          f"{prefix}.proj_in.weight": (c, c), f"{prefix}.proj_in.bias": (c,),
          f"{prefix}.proj_out.weight": (c, c), f"{prefix}.proj_out.bias": (c,)}
    tb = f"{prefix}.transformer_blocks.0"
    ks.update(_tblock_keys(tb, c, c))  # double_self_attention: attn2 has cross_attention_dim=None
    for a in ("attn1", "attn2"):
        p = f"{tb}.{a}.processor"
        for n in ("to_q_sp", "to_k_sp", "to_v_sp", "to_out_sp"):
            ks[f"{p}.{n}.weight"] = (c, c)
        ks[f"{p}.to_out_sp.bias"] = (c,)
        ks[f"{p}.time_pos_embed.pe"] = (1, cfg.motion_max_seq_length, c)
        ks[f"{p}.alpha_blender.mix_factor"] = (1,)
    return ks


def key_plan(cfg: UNetConfig) -> Dict[str, Tuple[int, ...]]:
    """Every tensor of the released model's state dict -> shape.  Construction order follows
    unet_motion_mv_model.py:123-273 and SURVEY.md Appendix B.8."""
    ch = cfg.block_out_channels
    temb = cfg.time_embed_dim
    ks: Dict[str, Tuple[int, ...]] = {}
    ks["conv_in.weight"] = (ch[0], cfg.in_channels, 3, 3)
    ks["conv_in.bias"] = (ch[0],)
    for name, din in (("time_embedding", ch[0]), ("camera_embedding", cfg.camera_embedding_dim)):
        ks[f"{name}.linear_1.weight"] = (temb, din)
        ks[f"{name}.linear_1.bias"] = (temb,)
        ks[f"{name}.linear_2.weight"] = (temb, temb)
        ks[f"{name}.linear_2.bias"] = (temb,)
    ip = "encoder_hid_proj.image_projection_layers.0"
    ks[f"{ip}.image_embeds.weight"] = (cfg.ip_num_tokens * cfg.cross_attention_dim, cfg.ip_image_embed_dim)
    ks[f"{ip}.image_embeds.bias"] = (cfg.ip_num_tokens * cfg.cross_attention_dim,)
    ks[f"{ip}.norm.weight"] = (cfg.cross_attention_dim,)
    ks[f"{ip}.norm.bias"] = (cfg.cross_attention_dim,)
    # down
    cout = ch[0]
    for i, c in enumerate(ch):
        cin, cout = cout, c
        for j in range(cfg.layers_per_block):
            ks.update(_resnet_keys(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, temb))
            if cfg.down_has_attn[i]:
                ks.update(_transformer2d_keys(f"down_blocks.{i}.attentions.{j}", cout, cfg))
            ks.update(_motion_keys(f"down_blocks.{i}.motion_modules.{j}", cout, cfg))
        if i != len(ch) - 1:
            ks[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            ks[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (cout,)
    # mid
    c = ch[-1]
    ks.update(_resnet_keys("mid_block.resnets.0", c, c, temb))
    ks.update(_transformer2d_keys("mid_block.attentions.0", c, cfg))
    ks.update(_motion_keys("mid_block.motion_modules.0", c, cfg))
    ks.update(_resnet_keys("mid_block.resnets.1", c, c, temb))
    # up
    for i, (cin_list, cout, has_attn, has_up) in enumerate(up_plan(cfg)):
        for j, cin in enumerate(cin_list):
            ks.update(_resnet_keys(f"up_blocks.{i}.resnets.{j}", cin, cout, temb))
            if has_attn:
                ks.update(_transformer2d_keys(f"up_blocks.{i}.attentions.{j}", cout, cfg))
            ks.update(_motion_keys(f"up_blocks.{i}.motion_modules.{j}", cout, cfg))
        if has_up:
            ks[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            ks[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (cout,)
    ks["conv_norm_out.weight"] = (ch[0],)
    ks["conv_norm_out.bias"] = (ch[0],)
    ks["conv_out.weight"] = (cfg.out_channels, ch[0], 3, 3)
    ks["conv_out.bias"] = (cfg.out_channels,)
    return ks


def skip_channels(cfg: UNetConfig) -> List[int]:
    """Channel count of every tensor pushed on the skip stack, in push order."""
    ch = cfg.block_out_channels
    out = [ch[0]]
    for i, c in enumerate(ch):
        out += [c] * cfg.layers_per_block
        if i != len(ch) - 1:
            out.append(c)
    return out


def up_plan(cfg: UNetConfig):
    """[(resnet input channels per layer, out channels, has spatial transformer, has upsampler)] for the 4 up blocks.
    Mirrors the prev_output_channel/input_channel arithmetic at unet_motion_mv_model.py:220-258 together with diffusers'
    `res_skip_channels = in_channels if i == num_layers-1 else out_channels` rule."""
    ch = cfg.block_out_channels
    rev = list(reversed(ch))
    n = len(ch)
    plan = []
    prev = rev[0]
    for i in range(n):
        out = rev[i]
        inp = rev[min(i + 1, n - 1)]
        layers = cfg.layers_per_block + 1
        cins = []
        for j in range(layers):
            res_skip = inp if j == layers - 1 else out
            res_in = prev if j == 0 else out
            cins.append(res_in + res_skip)
        has_attn = tuple(reversed(cfg.down_has_attn))[i]
        plan.append((cins, out, has_attn, i != n - 1))
        prev = out
    return plan


def sinusoidal_pe(c: int, max_len: int) -> Tensor:
    """diffusers SinusoidalPositionalEmbedding buffer `pe` [1,max_len,c]."""
    position = torch.arange(max_len, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, c, 2, dtype=torch.float32) * (-math.log(10000.0) / c))
    pe = torch.zeros(1, max_len, c)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe
