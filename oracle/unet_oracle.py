"""ORACLE (test infrastructure, NOT product code) -- fp32 CPU restatement of the MV-VDM UNet forward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

What it restates (reference = /root/reference, yanqinJiang/Animate3D @ 033a1be):
  * MVUNetMotionModel.forward                 animatediff/models/unet_motion_mv_model.py:633-867
  * block construction / channel plan         animatediff/models/unet_motion_mv_model.py:123-273
  * MVDreamI2VXFormersAttnProcessor           animatediff/models/attention_processor.py:302-445
  * IPAdapterXFormersAttnProcessor            animatediff/models/attention_processor.py:129-298
  * SpatioTemporalI2VXFormersAttnProcessor    animatediff/models/attention_processor.py:448-723
    (released config: spatial attn on, sinusoid 2-D enc, camera enc off, image attn off, alpha blender on;
     configs/inference/inference.yaml:9-24)
  * SinePositionalEncoding2D                  animatediff/models/embeddings.py:8-96
  * processor wiring / pos_embed=None         inference.py:107-192
The UNet body the reference imports from diffusers==0.28.0 (requirements.txt:2; NOT installed here, NOT vendored in
/root/reference) is restated from the published semantics recorded in SURVEY.md Appendix B:
  ResnetBlock2D, Downsample2D, Upsample2D, Attention, BasicTransformerBlock, FeedForward/GEGLU, Transformer2DModel,
  TransformerTemporalModel, Timesteps, TimestepEmbedding, SinusoidalPositionalEmbedding, ImageProjection, AlphaBlender,
  CrossAttn{Down,Up}BlockMotion / {Down,Up}BlockMotion / UNetMidBlockCrossAttnMotion.
xformers.ops.memory_efficient_attention (xformers==0.0.16, CUDA only) is restated as softmax(q k^T * scale) v.

PARITY PINNING: the reference ships no tests/goldens.  The four processors and the 2-D positional
encoding of this file ARE pinned against the reference's own source run in the build container with the absent third
party packages stubbed (tests/golden/gen_reference_goldens.py -> tests/golden/ref_*.pt).  The diffusers-0.28.0 body is
"parity unpinned" (restated from published semantics only) -- stated again in DESIGN.md.

All state-dict keys follow the reference/diffusers naming so that the 726-missing-keys invariant (inference.py:222)
can be checked (tests/test_state_dict_layout.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

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


def make_state_dict(cfg: UNetConfig, seed: int = 0, std: float = 0.02) -> Dict[str, Tensor]:
    """Seeded random weights per SURVEY section 8(d) config 1: N(0,std^2) for Linear/Conv weights and biases, norm affine
    (1,0)+small noise so that the affine path is exercised, zero-init branches overridden to N(0,std^2), mix_factor 0."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in key_plan(cfg).items():
        if k.endswith("time_pos_embed.pe"):
            sd[k] = sinusoidal_pe(shape[2], shape[1])
        elif k.endswith("mix_factor"):
            sd[k] = torch.zeros(shape)
        elif ".norm" in k or k.startswith("conv_norm_out") or k.endswith("norm.weight") or k.endswith("norm.bias"):
            noise = torch.randn(shape, generator=g) * 0.05
            sd[k] = (1.0 + noise) if k.endswith("weight") else noise
        else:
            sd[k] = torch.randn(shape, generator=g) * std
    return sd


# --------------------------------------------------------------------------------------------------------------------
# diffusers-0.28.0 pieces
# --------------------------------------------------------------------------------------------------------------------

def timesteps_proj(t: Tensor, dim: int = 320) -> Tensor:
    """Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) -- unet_motion_mv_model.py:133, Appendix B.1."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def timestep_embedding(sd, name, x):
    x = F.linear(x, sd[f"{name}.linear_1.weight"], sd[f"{name}.linear_1.bias"])
    x = F.silu(x)
    return F.linear(x, sd[f"{name}.linear_2.weight"], sd[f"{name}.linear_2.bias"])


def resnet_block(sd, p, x, emb, groups, eps):
    """ResnetBlock2D (Appendix B.2)."""
    h = F.group_norm(x, groups, sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], eps)
    h = F.conv2d(F.silu(h), sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1)
    t = F.linear(F.silu(emb), sd[f"{p}.time_emb_proj.weight"], sd[f"{p}.time_emb_proj.bias"])
    h = h + t[:, :, None, None]
    h = F.group_norm(h, groups, sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], eps)
    h = F.conv2d(F.silu(h), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
    if f"{p}.conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[f"{p}.conv_shortcut.weight"], sd[f"{p}.conv_shortcut.bias"])
    return x + h


def head_to_batch(x: Tensor, heads: int) -> Tensor:
    b, l, c = x.shape
    return x.reshape(b, l, heads, c // heads).permute(0, 2, 1, 3).reshape(b * heads, l, c // heads)


def batch_to_head(x: Tensor, heads: int) -> Tensor:
    bh, l, d = x.shape
    return x.reshape(bh // heads, heads, l, d).permute(0, 2, 1, 3).reshape(bh // heads, l, heads * d)


def mea(q: Tensor, k: Tensor, v: Tensor, scale: float) -> Tensor:
    """xformers.ops.memory_efficient_attention(q,k,v,attn_bias=None,scale) on [B*heads, L, d] inputs."""
    # batches are independent: evaluate them in chunks so the [chunk, L, L] score tensor stays below ~1 GiB (same result;
    # the headline 4-view x 16-frame config would otherwise materialise 2 x 8.6 GB per cross-view attention)
    step = max(1, (1 << 28) // max(1, q.shape[1] * k.shape[1]))
    if step >= q.shape[0]:
        s = torch.bmm(q, k.transpose(1, 2)) * scale
        return torch.bmm(torch.softmax(s, dim=-1), v)
    out = torch.empty(q.shape[0], q.shape[1], v.shape[2], dtype=q.dtype)
    for i in range(0, q.shape[0], step):
        s = torch.bmm(q[i:i + step], k[i:i + step].transpose(1, 2)) * scale
        out[i:i + step] = torch.bmm(torch.softmax(s, dim=-1), v[i:i + step])
    return out


def feed_forward(sd, p, x):
    """FeedForward(activation_fn='geglu') (Appendix B.5): proj -> (u, g) chunk -> u * gelu_exact(g) -> Linear."""
    h = F.linear(x, sd[f"{p}.net.0.proj.weight"], sd[f"{p}.net.0.proj.bias"])
    u, g = h.chunk(2, dim=-1)
    return F.linear(u * F.gelu(g), sd[f"{p}.net.2.weight"], sd[f"{p}.net.2.bias"])


# --------------------------------------------------------------------------------------------------------------------
# the reference's processors
# --------------------------------------------------------------------------------------------------------------------

def sine_pos_enc_2d(num_feats: int, h: int, w: int, temperature=10000, scale=2 * math.pi, eps=1e-6) -> Tensor:
    """SinePositionalEncoding2D(num_feats, normalize=True)._forward on an all-valid mask -> [2*num_feats, h, w]
    (animatediff/models/embeddings.py:58-96)."""
    y_embed = torch.arange(1, h + 1, dtype=torch.float32)[:, None].expand(h, w)
    x_embed = torch.arange(1, w + 1, dtype=torch.float32)[None, :].expand(h, w)
    y_embed = y_embed / (y_embed[-1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x = x_embed[:, :, None] / dim_t
    pos_y = y_embed[:, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, 0::2].sin(), pos_x[:, :, 1::2].cos()), dim=3).reshape(h, w, -1)
    pos_y = torch.stack((pos_y[:, :, 0::2].sin(), pos_y[:, :, 1::2].cos()), dim=3).reshape(h, w, -1)
    return torch.cat((pos_y, pos_x), dim=2).permute(2, 0, 1)


def proc_mv_i2v(sd, ap, x, heads, nv, nf):
    """MVDreamI2VXFormersAttnProcessor.__call__ (attention_processor.py:325-445), self-attention call.
    x: [(b n f), l, c]; `ap` = '<...>.attn1'."""
    bnf, l, c = x.shape
    b = bnf // (nv * nf)
    scale = (c // heads) ** -0.5
    # "(b n f) l c -> (b f) (n l) c"   (line 340)
    h = x.reshape(b, nv, nf, l, c).permute(0, 2, 1, 3, 4).reshape(b * nf, nv * l, c)
    q = F.linear(h, sd[f"{ap}.to_q.weight"])
    k = F.linear(h, sd[f"{ap}.to_k.weight"])
    v = F.linear(h, sd[f"{ap}.to_v.weight"])
    # frame-0 K/V broadcast to every frame (389-397)
    k0 = k.reshape(b, nf, nv * l, c)[:, 0:1].expand(b, nf, nv * l, c).reshape(b * nf, nv * l, c)
    v0 = v.reshape(b, nf, nv * l, c)[:, 0:1].expand(b, nf, nv * l, c).reshape(b * nf, nv * l, c)
    o = batch_to_head(mea(head_to_batch(q, heads), head_to_batch(k, heads), head_to_batch(v, heads), scale), heads)
    qi = F.linear(h, sd[f"{ap}.processor.to_q_i2v.weight"])                                     # 413
    oi = batch_to_head(mea(head_to_batch(qi, heads), head_to_batch(k0, heads), head_to_batch(v0, heads), scale), heads)
    oi = F.linear(oi, sd[f"{ap}.processor.to_out_i2v.weight"], sd[f"{ap}.processor.to_out_i2v.bias"])  # 423
    o = o + oi                                                                                   # 426
    o = F.linear(o, sd[f"{ap}.to_out.0.weight"], sd[f"{ap}.to_out.0.bias"])                      # 429
    # "(b f) (n l) c -> (b n f) l c"   (line 443)
    return o.reshape(b, nf, nv, l, c).permute(0, 2, 1, 3, 4).reshape(bnf, l, c)


def proc_ip_adapter(sd, ap, x, text, ip_tokens, heads, ip_scale):
    """IPAdapterXFormersAttnProcessor.__call__ (attention_processor.py:169-298).
    x [N, l, c]; text [N, 77, 768]; ip_tokens [N, 4, 768] (the reference holds [N,1,4,768]; head_to_batch_dim folds the
    extra dim into the sequence, Appendix B.1)."""
    c = x.shape[-1]
    scale = (c // heads) ** -0.5
    q = head_to_batch(F.linear(x, sd[f"{ap}.to_q.weight"]), heads)
    k = head_to_batch(F.linear(text, sd[f"{ap}.to_k.weight"]), heads)
    v = head_to_batch(F.linear(text, sd[f"{ap}.to_v.weight"]), heads)
    o = batch_to_head(mea(q, k, v, scale), heads)
    ik = head_to_batch(F.linear(ip_tokens, sd[f"{ap}.processor.to_k_ip.0.weight"]), heads)
    iv = head_to_batch(F.linear(ip_tokens, sd[f"{ap}.processor.to_v_ip.0.weight"]), heads)
    o = o + ip_scale * batch_to_head(mea(q, ik, iv, scale), heads)                               # 283
    return F.linear(o, sd[f"{ap}.to_out.0.weight"], sd[f"{ap}.to_out.0.bias"])


def proc_spatiotemporal(sd, ap, x, heads, nv, nf, fs):
    """SpatioTemporalI2VXFormersAttnProcessor.__call__ (attention_processor.py:541-723), released config.
    x: [(b n h w), f, c]; `ap` = '<...>.attn1' or '.attn2'."""
    bl, f, c = x.shape
    assert f == nf
    L = nv * fs * fs
    b = bl // L
    scale = (c // heads) ** -0.5
    pp = f"{ap}.processor"
    # ---- spatial branch input (555-563): "(b l) f c -> (b f) l c", + 2-D sinusoid per view
    xs = x.reshape(b, L, f, c).permute(0, 2, 1, 3).reshape(b * f, L, c)
    pos = sine_pos_enc_2d(c // 2, fs, fs)                       # [c, fs, fs]
    pos_tok = pos.permute(1, 2, 0).reshape(1, 1, fs * fs, c)    # per view "(h w) c"
    xs = (xs.reshape(b * f, nv, fs * fs, c) + pos_tok.to(xs.dtype)).reshape(b * f, L, c)
    # ---- restore temporal encoding (583-584): hidden_states = time_pos_embed(hidden_states)
    xt = x + sd[f"{pp}.time_pos_embed.pe"][:, :f].to(x.dtype)
    # ---- temporal attention (620-641), materialised probabilities
    q = head_to_batch(F.linear(xt, sd[f"{ap}.to_q.weight"]), heads)
    k = head_to_batch(F.linear(xt, sd[f"{ap}.to_k.weight"]), heads)
    v = head_to_batch(F.linear(xt, sd[f"{ap}.to_v.weight"]), heads)
    t_out = batch_to_head(mea(q, k, v, scale), heads)
    t_out = F.linear(t_out, sd[f"{ap}.to_out.0.weight"], sd[f"{ap}.to_out.0.bias"])
    # ---- spatial attention (644-669)
    qs = head_to_batch(F.linear(xs, sd[f"{pp}.to_q_sp.weight"]), heads)
    ks = head_to_batch(F.linear(xs, sd[f"{pp}.to_k_sp.weight"]), heads)
    vs = head_to_batch(F.linear(xs, sd[f"{pp}.to_v_sp.weight"]), heads)
    s_out = batch_to_head(mea(qs, ks, vs, scale), heads)
    s_out = F.linear(s_out, sd[f"{pp}.to_out_sp.weight"], sd[f"{pp}.to_out_sp.bias"])
    s_out = s_out.reshape(b, f, L, c).permute(0, 2, 1, 3).reshape(bl, f, c)   # "(b f) l c -> (b l) f c"
    # ---- AlphaBlender(alpha=0,'learned')(x_spatial=s_out, x_temporal=t_out) (709; Appendix B.9)
    alpha = torch.sigmoid(sd[f"{pp}.alpha_blender.mix_factor"]).to(x.dtype)
    return alpha * s_out + (1.0 - alpha) * t_out


# --------------------------------------------------------------------------------------------------------------------
# transformer wrappers
# --------------------------------------------------------------------------------------------------------------------

def transformer2d(sd, p, x, text, ip_tokens, cfg: UNetConfig, nv, nf):
    """Transformer2DModel + BasicTransformerBlock (Appendix B.5/B.6) with the reference processors."""
    n, c, h, w = x.shape
    heads = cfg.num_attention_heads
    res = x
    y = F.group_norm(x, cfg.norm_num_groups, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    y = F.conv2d(y, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    y = y.permute(0, 2, 3, 1).reshape(n, h * w, c)
    tb = f"{p}.transformer_blocks.0"
    yn = F.layer_norm(y, (c,), sd[f"{tb}.norm1.weight"], sd[f"{tb}.norm1.bias"], 1e-5)
    y = proc_mv_i2v(sd, f"{tb}.attn1", yn, heads, nv, nf) + y
    yn = F.layer_norm(y, (c,), sd[f"{tb}.norm2.weight"], sd[f"{tb}.norm2.bias"], 1e-5)
    y = proc_ip_adapter(sd, f"{tb}.attn2", yn, text, ip_tokens, heads, cfg.ip_scale) + y
    yn = F.layer_norm(y, (c,), sd[f"{tb}.norm3.weight"], sd[f"{tb}.norm3.bias"], 1e-5)
    y = feed_forward(sd, f"{tb}.ff", yn) + y
    y = y.reshape(n, h, w, c).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    return y + res


def transformer_temporal(sd, p, x, cfg: UNetConfig, nv, nf, fs):
    """TransformerTemporalModel (Appendix B.7): GroupNorm over (C/32, F, h, w), tokens [(B' h w), F, C]."""
    n, c, h, w = x.shape
    heads = cfg.motion_num_attention_heads
    bp = n // nf
    res = x
    y = x.reshape(bp, nf, c, h, w).permute(0, 2, 1, 3, 4)
    y = F.group_norm(y, cfg.norm_num_groups, sd[f"{p}.norm.weight"], sd[f"{p}.norm.bias"], 1e-6)
    y = y.permute(0, 3, 4, 2, 1).reshape(bp * h * w, nf, c)
    y = F.linear(y, sd[f"{p}.proj_in.weight"], sd[f"{p}.proj_in.bias"])
    tb = f"{p}.transformer_blocks.0"
    # pos_embed is None (inference.py:177-192); the processor re-adds the temporal encoding itself
    yn = F.layer_norm(y, (c,), sd[f"{tb}.norm1.weight"], sd[f"{tb}.norm1.bias"], 1e-5)
    y = proc_spatiotemporal(sd, f"{tb}.attn1", yn, heads, nv, nf, fs) + y
    yn = F.layer_norm(y, (c,), sd[f"{tb}.norm2.weight"], sd[f"{tb}.norm2.bias"], 1e-5)
    y = proc_spatiotemporal(sd, f"{tb}.attn2", yn, heads, nv, nf, fs) + y
    yn = F.layer_norm(y, (c,), sd[f"{tb}.norm3.weight"], sd[f"{tb}.norm3.bias"], 1e-5)
    y = feed_forward(sd, f"{tb}.ff", yn) + y
    y = F.linear(y, sd[f"{p}.proj_out.weight"], sd[f"{p}.proj_out.bias"])
    y = y.reshape(bp, h, w, nf, c).permute(0, 3, 4, 1, 2).reshape(n, c, h, w)
    return y + res


# --------------------------------------------------------------------------------------------------------------------
# full forward
# --------------------------------------------------------------------------------------------------------------------

def unet_forward(sd: Dict[str, Tensor], cfg: UNetConfig, sample: Tensor, timestep, encoder_hidden_states: Tensor,
                 camera: Optional[Tensor], image_embeds: Tensor, num_views: int, i2v_cond_time_zero: bool = False,
                 taps: Optional[dict] = None) -> Tensor:
    """MVUNetMotionModel.forward (unet_motion_mv_model.py:633-867).
    sample [B*Nv, 4, F, h, w]; timestep scalar or [B*Nv]; encoder_hidden_states [B*Nv,77,768]; camera [B*Nv,16];
    image_embeds [B*Nv,1024].  Returns [B*Nv, 4, F, h, w].  `taps`, if given, collects intermediate activations."""
    bn, _, nf, h0, w0 = sample.shape
    assert bn % num_views == 0                                                     # 684
    groups, eps = cfg.norm_num_groups, cfg.norm_eps
    t = torch.as_tensor(timestep)
    if t.ndim == 0:
        t = t[None]
    t = t.expand(bn)                                                               # 721
    emb = timestep_embedding(sd, "time_embedding", timesteps_proj(t, cfg.block_out_channels[0]))
    if i2v_cond_time_zero:
        cond_emb = timestep_embedding(sd, "time_embedding", timesteps_proj(torch.zeros_like(t), cfg.block_out_channels[0]))
    if camera is not None:
        cam = timestep_embedding(sd, "camera_embedding", camera)
        emb = emb + cam
        if i2v_cond_time_zero:
            cond_emb = cond_emb + cam
    emb = emb.repeat_interleave(nf, dim=0)                                         # 747
    if i2v_cond_time_zero:                                                         # 748-752
        emb = emb.reshape(bn, nf, -1)
        emb = torch.cat([cond_emb[:, None], emb[:, 1:]], dim=1).reshape(bn * nf, -1)
    text = encoder_hidden_states.repeat_interleave(nf, dim=0)                      # 754
    ipp = "encoder_hid_proj.image_projection_layers.0"
    ip = F.linear(image_embeds, sd[f"{ipp}.image_embeds.weight"], sd[f"{ipp}.image_embeds.bias"])
    ip = ip.reshape(bn, cfg.ip_num_tokens, cfg.cross_attention_dim)
    ip = F.layer_norm(ip, (cfg.cross_attention_dim,), sd[f"{ipp}.norm.weight"], sd[f"{ipp}.norm.bias"], 1e-5)
    ip = ip.repeat_interleave(nf, dim=0)                                           # 763

    x = sample.permute(0, 2, 1, 3, 4).reshape(bn * nf, -1, h0, w0)                 # 767
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)           # 768
    if taps is not None:
        taps["conv_in"] = x
    skips = [x]
    nlev = len(cfg.block_out_channels)
    for i in range(nlev):
        fs = cfg.feature_size(i)
        for j in range(cfg.layers_per_block):
            x = resnet_block(sd, f"down_blocks.{i}.resnets.{j}", x, emb, groups, eps)
            if cfg.down_has_attn[i]:
                x = transformer2d(sd, f"down_blocks.{i}.attentions.{j}", x, text, ip, cfg, num_views, nf)
            x = transformer_temporal(sd, f"down_blocks.{i}.motion_modules.{j}", x, cfg, num_views, nf, fs)
            skips.append(x)
            if taps is not None:
                taps[f"down_blocks.{i}.{j}"] = x
        if i != nlev - 1:
            p = f"down_blocks.{i}.downsamplers.0.conv"
            x = F.conv2d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], stride=2, padding=1)
            skips.append(x)
    fs = cfg.feature_size(nlev - 1)
    x = resnet_block(sd, "mid_block.resnets.0", x, emb, groups, eps)
    x = transformer2d(sd, "mid_block.attentions.0", x, text, ip, cfg, num_views, nf)
    x = transformer_temporal(sd, "mid_block.motion_modules.0", x, cfg, num_views, nf, fs)
    x = resnet_block(sd, "mid_block.resnets.1", x, emb, groups, eps)
    if taps is not None:
        taps["mid"] = x
    for i, (cins, cout, has_attn, has_up) in enumerate(up_plan(cfg)):
        fs = cfg.feature_size(nlev - 1 - i)
        for j in range(len(cins)):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet_block(sd, f"up_blocks.{i}.resnets.{j}", x, emb, groups, eps)
            if has_attn:
                x = transformer2d(sd, f"up_blocks.{i}.attentions.{j}", x, text, ip, cfg, num_views, nf)
            x = transformer_temporal(sd, f"up_blocks.{i}.motion_modules.{j}", x, cfg, num_views, nf, fs)
            if taps is not None:
                taps[f"up_blocks.{i}.{j}"] = x
        if has_up:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            p = f"up_blocks.{i}.upsamplers.0.conv"
            x = F.conv2d(x, sd[f"{p}.weight"], sd[f"{p}.bias"], padding=1)
    x = F.group_norm(x, groups, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], eps)   # 856
    x = F.conv2d(F.silu(x), sd["conv_out.weight"], sd["conv_out.bias"], padding=1)           # 859
    return x.reshape(bn, nf, -1, h0, w0).permute(0, 2, 1, 3, 4)                              # 862


# --------------------------------------------------------------------------------------------------------------------
# camera helper (pipeline.py:127-190) and synthetic inputs
# --------------------------------------------------------------------------------------------------------------------

def get_camera(num_views: int, elevation: float = 15.0, azimuth_start: float = 0.0, azimuth_span: float = 360.0):
    """pipeline.py:127-190 get_camera/generate_c2w/normalize_camera -> [num_views, 16]."""
    cams = []
    gap = azimuth_span / num_views
    az = azimuth_start
    for _ in range(num_views):
        e = torch.tensor([elevation * math.pi / 180])
        a = torch.tensor([az * math.pi / 180])
        pos = torch.stack([torch.cos(e) * torch.cos(a), torch.cos(e) * torch.sin(a), torch.sin(e)], dim=-1)
        up = torch.tensor([[0.0, 0.0, 1.0]])
        lookat = F.normalize(-pos, dim=-1)
        right = F.normalize(torch.linalg.cross(lookat, up), dim=-1)
        up = F.normalize(torch.linalg.cross(right, lookat), dim=-1)
        c2w = torch.eye(4)
        c2w[:3, :3] = torch.stack([right, up, -lookat], dim=-1)[0]
        tr = pos[0]
        c2w[:3, 3] = tr / (torch.norm(tr) + 1e-8)
        cams.append(c2w.flatten())
        az += gap
    return torch.stack(cams, 0).float()


def synthetic_inputs(cfg: UNetConfig, groups: int, num_views: int, num_frames: int, seed: int = 0):
    """Seeded synthetic inputs of SURVEY 8(d): sample seed, text seed+1, image embeds seed+2; camera = get_camera."""
    bn = groups * num_views
    g0 = torch.Generator().manual_seed(seed)
    g1 = torch.Generator().manual_seed(seed + 1)
    g2 = torch.Generator().manual_seed(seed + 2)
    sample = torch.randn(bn, cfg.in_channels, num_frames, cfg.sample_size, cfg.sample_size, generator=g0)
    text = torch.randn(bn, 77, cfg.cross_attention_dim, generator=g1)
    img = torch.randn(bn, cfg.ip_image_embed_dim, generator=g2)
    camera = get_camera(num_views).repeat(groups, 1)
    return sample, text, camera, img
