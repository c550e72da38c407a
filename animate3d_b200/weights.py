"""Seeded random-init weights of the released geometry (there is no network for checkpoints: bench and tests use these).
Distribution per SURVEY section 8(d): N(0, 0.02^2) Linear/Conv weights and biases, norm affine (1, 0) + small noise, the
reference's zero-initialised branches (to_out_i2v, to_out_sp) overridden with N(0, 0.02^2) so they are exercised,
alpha-blender mix_factor 0, time_pos_embed.pe = the sinusoidal buffer."""
from __future__ import annotations

from typing import Dict

import torch

from .unet_config import UNetConfig, key_plan, sinusoidal_pe


def random_state_dict(cfg: UNetConfig, seed: int = 0, device="cuda", std: float = 0.02) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    for k, shape in key_plan(cfg).items():
        if k.endswith("time_pos_embed.pe"):
            sd[k] = sinusoidal_pe(shape[2], shape[1]).to(device)
        elif k.endswith("mix_factor"):
            sd[k] = torch.zeros(shape, device=device)
        elif ".norm" in k or k.startswith("conv_norm_out"):
            noise = torch.randn(shape, generator=g, device=device) * 0.05
            sd[k] = (1.0 + noise) if k.endswith("weight") else noise
        else:
            sd[k] = torch.randn(shape, generator=g, device=device) * std
    return sd


def load_unet_checkpoint(unet, path: str, map_location="cpu"):
    """inference.py:213-223: `torch.load`, unwrap an optional "state_dict" entry, `load_state_dict(strict=False)` and the
    reference's two invariants -- nothing unexpected, and either the full model (0 missing) or the released motion-module
    checkpoint (exactly the 726 non-motion keys missing).  Returns (missing, unexpected)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=True)
    sd = ckpt["state_dict"] if isinstance(ckpt, dict) and "state_dict" in ckpt else ckpt
    missing, unexpected = unet.load_state_dict(sd, strict=False)
    if unexpected:
        raise ValueError(f"{path}: {len(unexpected)} unexpected keys (file is broken), e.g. {unexpected[:3]}")
    if len(missing) not in (0, 726):
        raise ValueError(f"{path}: {len(missing)} missing keys; expected 0 (full model) or 726 (motion modules only)")
    return missing, unexpected
