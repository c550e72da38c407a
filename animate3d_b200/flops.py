"""Algorithmic FLOP count (2*MAC on unpadded dims) of one MVUNetMotionModel forward, enumerated from the layer list of
SURVEY.md section 3.2 / Appendix B.8.  bench.py's roofline numbers and the cpu_baseline extrapolation use these figures;
tests/test_flops.py asserts them against SURVEY section 8(d) (25.97 TFLOP per CFG branch at 4 views x 16 frames)."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict

from .unet_config import UNetConfig, up_plan


def unet_forward_flops(cfg: UNetConfig, groups: int, num_views: int, num_frames: int, n_text: int = 77) -> Dict[str, float]:
    """FLOPs by category for `groups` (prompt x CFG-branch) groups of num_views x num_frames latent images."""
    f = defaultdict(float)
    N = groups * num_views * num_frames
    ch = cfg.block_out_channels
    fs0 = cfg.sample_size

    def conv(cin, cout, h, k=3):
        return 2.0 * N * h * h * cin * cout * k * k

    def lin(tokens, cin, cout):
        return 2.0 * tokens * cin * cout

    def resnet(cin, cout, h):
        f["conv"] += conv(cin, cout, h) + conv(cout, cout, h)
        if cin != cout:
            f["conv"] += conv(cin, cout, h, 1)
        f["temb"] += lin(N, cfg.time_embed_dim, cout)

    def transformer2d(c, h):
        T = N * h * h
        f["conv"] += 2 * conv(c, c, h, 1)
        f["attn1_proj"] += 4 * lin(T, c, c) + 2 * lin(T, c, c)            # q,k,v,out + q_i2v,out_i2v
        L = num_views * h * h
        f["mv_qkpv"] += groups * num_frames * 4.0 * L * L * c
        f["i2v_qkpv"] += groups * num_frames * 4.0 * L * L * c
        nk = n_text + cfg.ip_num_tokens
        f["attn2_proj"] += 2 * lin(T, c, c) + 2 * lin(N * n_text, cfg.cross_attention_dim, c) \
            + 2 * lin(N * cfg.ip_num_tokens, cfg.cross_attention_dim, c)
        f["attn2_qkpv"] += 4.0 * T * nk * c
        f["spatial_ff"] += lin(T, c, 8 * c) + lin(T, 4 * c, c)

    def motion(c, h):
        T = N * h * h
        f["motion_inout"] += 2 * lin(T, c, c)
        L = num_views * h * h
        for _ in range(2):
            f["temporal_proj"] += 4 * lin(T, c, c)
            f["temporal_qkpv"] += (T / num_frames) * 4.0 * num_frames * num_frames * c
            f["spatial_proj"] += 4 * lin(T, c, c)
            f["spatial_qkpv"] += groups * num_frames * 4.0 * L * L * c
        f["motion_ff"] += lin(T, c, 8 * c) + lin(T, 4 * c, c)

    f["conv"] += conv(cfg.in_channels, ch[0], fs0)
    cout = ch[0]
    for i, c in enumerate(ch):
        cin, cout = cout, c
        h = fs0 >> i
        for j in range(cfg.layers_per_block):
            resnet(cin if j == 0 else cout, cout, h)
            if cfg.down_has_attn[i]:
                transformer2d(cout, h)
            motion(cout, h)
        if i != len(ch) - 1:
            f["conv"] += 2.0 * N * (h // 2) * (h // 2) * cout * cout * 9
    h = fs0 >> (len(ch) - 1)
    resnet(ch[-1], ch[-1], h)
    transformer2d(ch[-1], h)
    motion(ch[-1], h)
    resnet(ch[-1], ch[-1], h)
    for i, (cins, cout, has_attn, has_up) in enumerate(up_plan(cfg)):
        h = fs0 >> (len(ch) - 1 - i)
        for cin in cins:
            resnet(cin, cout, h)
            if has_attn:
                transformer2d(cout, h)
            motion(cout, h)
        if has_up:
            f["conv"] += conv(cout, cout, 2 * h)
    f["conv"] += conv(ch[0], cfg.out_channels, fs0)
    # time / camera embedding MLPs
    bn = groups * num_views
    f["temb"] += lin(bn, ch[0], cfg.time_embed_dim) + 2 * lin(bn, cfg.time_embed_dim, cfg.time_embed_dim) \
        + lin(bn, cfg.camera_embedding_dim, cfg.time_embed_dim)
    out = dict(f)
    out["total"] = sum(f.values())
    return out


def attention_call_flops(batches: int, L: int, c: int) -> float:
    """QK^T + PV of one fused cross-view attention launch: batches * 4 * L^2 * C (C = heads*d)."""
    return batches * 4.0 * L * L * c
