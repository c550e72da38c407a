"""ORACLE (test infrastructure) -- per-frame deformation field of the 4D gaussians:
k-planes (HexPlane) multi-scale feature lookup + three bias-free MLPs, as in
custom/threestudio-animate3d/geometry/gaussian_4d.py:39-64 (grid_sample_wrapper), 101-147 (construction, zero-init of the
last layers), 151-174 (init_grid_param), 450-484 (interpolate_ms_features), 486-548 (get_scaling/get_rotation/get_xyz without
the optional global rot/trans) and threestudio/models/networks.py:214-251 (VanillaMLP: Linear(no bias)-ReLU-Linear(no bias)).

PINNED: `interpolate_ms_features` is checked against golden vectors produced by the reference's own source
(tests/golden/gen_reference_goldens.py executes grid_sample_wrapper + interpolate_ms_features straight from the file)."""
from __future__ import annotations

import itertools
from typing import List, Sequence

import torch
import torch.nn.functional as F


def init_grids(grid_size: Sequence[Sequence[int]], n_grid_dims: int = 16, seed: int = 0, a: float = 0.1, b: float = 0.5):
    """gaussian_4d.py:101-117, 151-174: for each scale, 6 planes over the coordinate pairs of (x,y,z,t); planes touching t
    initialised to 1, others U(a,b).  Returns list (per scale) of lists (per plane) of [1, C, reso[j], reso[i]] tensors."""
    g = torch.Generator().manual_seed(seed)
    grids = []
    for reso in grid_size:
        planes = []
        for comb in itertools.combinations(range(4), 2):
            shape = [1, n_grid_dims] + [reso[cc] for cc in comb[::-1]]
            if 3 in comb:
                planes.append(torch.ones(shape))
            else:
                planes.append(torch.rand(shape, generator=g) * (b - a) + a)
        grids.append(planes)
    return grids


def interpolate_ms_features(pts: torch.Tensor, grids) -> torch.Tensor:
    """pts [P,4] in [-1,1] -> [P, C * num_scales]: product over the 6 planes of bilinear samples (align_corners=True,
    border padding), concatenated over scales."""
    out = []
    for planes in grids:
        feat = 1.0
        for ci, comb in enumerate(itertools.combinations(range(4), 2)):
            coords = pts[:, list(comb)].reshape(1, 1, -1, 2)
            s = F.grid_sample(planes[ci], coords, align_corners=True, mode="bilinear", padding_mode="border")
            feat = feat * s.reshape(planes[ci].shape[1], -1).t()
        out.append(feat)
    return torch.cat(out, dim=-1)


def mlp(x, w1, w2):
    return F.linear(F.relu(F.linear(x, w1)), w2)


def deform(xyz, scaling_raw, rotation_raw, t: float, grids, mlps, deform_scale: bool = True):
    """diff_gaussian_rasterizer_advanced_4d.py:77-83, 119-135: returns (means3D, scales, rotations) of the frame at time t.
    mlps = dict(xyz=(w1,w2), rot=(w1,w2), scale=(w1,w2)).  Activations: exp for scales, normalize for rotations."""
    pts = torch.cat([xyz, torch.full_like(xyz[:, :1], t)], dim=-1)
    h = interpolate_ms_features(pts, grids)
    means = xyz + mlp(h, *mlps["xyz"])
    sc = scaling_raw + (mlp(h, *mlps["scale"]) if deform_scale else 0.0)
    rot = rotation_raw + mlp(h, *mlps["rot"])
    return means, torch.exp(sc), F.normalize(rot, dim=-1)
