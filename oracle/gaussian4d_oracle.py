"""ORACLE (test infrastructure) -- per-frame deformation field of the 4D gaussians:
k-planes (HexPlane) multi-scale feature lookup + three bias-free MLPs, as in
custom/threestudio-animate3d/geometry/gaussian_4d.py:39-64 (grid_sample_wrapper), 101-147 (construction, zero-init of the
last layers), 151-174 (init_grid_param), 450-484 (interpolate_ms_features), 486-548 (get_scaling/get_rotation/get_xyz incl.
the `use_global_trans` branch, on in configs/refine_frame_16.yaml:56), geometry/utils.py:33-62 (build_rotation), 73-133
(extract_rotation_torch), 135-167 (euler_angles_to_rotation_matrix) and threestudio/models/networks.py:214-251 (VanillaMLP:
Linear(no bias)-ReLU-Linear(no bias)).

PINNED: `interpolate_ms_features` and the three rotation helpers are checked against golden vectors produced by the
reference's own source (tests/golden/gen_reference_goldens.py executes the functions straight from the files)."""
from __future__ import annotations

import itertools
import math
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


def quat_to_matrix(q):
    """geometry/utils.py:33-62: [N,4] (r,x,y,z), normalised first -> [N,3,3]."""
    q = q / q.norm(dim=1, keepdim=True)
    r, x, y, z = q.unbind(1)
    rows = [1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
            2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
            2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)]
    return torch.stack(rows, dim=1).reshape(-1, 3, 3)


def matrix_to_quat(m):
    """geometry/utils.py:73-133: branch on the trace / largest diagonal entry (strict '>' comparisons, in that order), then
    normalise.  Written with torch.where so that autograd sees the selected branch only through its own values."""
    m00, m01, m02 = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    m10, m11, m12 = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    m20, m21, m22 = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    tr = m00 + m11 + m22
    c1 = tr > 0
    c2 = (~c1) & (m00 > m11) & (m00 > m22)
    c3 = (~c1) & (~c2) & (m11 > m22)
    one = torch.ones_like(tr)
    # arguments of the square roots, made safe (=1) where the branch is not taken so that no NaN gradient leaks
    a1 = torch.where(c1, tr + 1.0, one)
    a2 = torch.where(c2, 1.0 + m00 - m11 - m22, one)
    a3 = torch.where(c3, 1.0 + m11 - m00 - m22, one)
    c4 = (~c1) & (~c2) & (~c3)
    a4 = torch.where(c4, 1.0 + m22 - m00 - m11, one)
    t1, t2, t3, t4 = [torch.sqrt(a) * 2 for a in (a1, a2, a3, a4)]
    cand = [
        (0.25 * t1, (m21 - m12) / t1, (m02 - m20) / t1, (m10 - m01) / t1),
        ((m21 - m12) / t2, 0.25 * t2, (m01 + m10) / t2, (m02 + m20) / t2),
        ((m02 - m20) / t3, (m01 + m10) / t3, 0.25 * t3, (m12 + m21) / t3),
        ((m10 - m01) / t4, (m02 + m20) / t4, (m12 + m21) / t4, 0.25 * t4),
    ]
    comps = []
    for k in range(4):
        comps.append(torch.where(c1, cand[0][k], torch.where(c2, cand[1][k], torch.where(c3, cand[2][k], cand[3][k]))))
    q = torch.stack(comps, dim=1)
    return q / q.norm(p=2, dim=1, keepdim=True)


def euler_to_matrix(angles):
    """geometry/utils.py:135-167: angles [3] = (roll, pitch, yaw) -> Rz(yaw) @ Ry(pitch) @ Rx(roll)."""
    cr, sr = torch.cos(angles[0]), torch.sin(angles[0])
    cp, sp = torch.cos(angles[1]), torch.sin(angles[1])
    cy, sy = torch.cos(angles[2]), torch.sin(angles[2])
    z, o = torch.zeros_like(cr), torch.ones_like(cr)
    rx = torch.stack([o, z, z, z, cr, -sr, z, sr, cr]).reshape(3, 3)
    ry = torch.stack([cp, z, sp, z, o, z, -sp, z, cp]).reshape(3, 3)
    rz = torch.stack([cy, -sy, z, sy, cy, z, z, z, o]).reshape(3, 3)
    return rz @ (ry @ rx)


def deform(xyz, scaling_raw, rotation_raw, t: float, grids, mlps, deform_scale: bool = True):
    """diff_gaussian_rasterizer_advanced_4d.py:77-83, 119-135: returns (means3D, scales, rotations) of the frame at time t.
    mlps = dict(xyz=(w1,w2), rot=(w1,w2), scale=(w1,w2)[, global_rot=(w1,w2), global_trans=(w1,w2)]).  Activations: exp for
    scales, normalize for rotations.  With the two global MLPs present (use_global_trans, gaussian_4d.py:499-511, 525-539):
    h_g = mean_p h; angles = sigmoid(MLP(h_g)) 2 pi - pi; trans = sigmoid(MLP(h_g)) 2 - 1; xyz <- R xyz + trans;
    q <- matrix_to_quat(R @ quat_to_matrix(q)) BEFORE the per-point deltas are added."""
    pts = torch.cat([xyz, torch.full_like(xyz[:, :1], t)], dim=-1)
    h = interpolate_ms_features(pts, grids)
    base_xyz, base_rot = xyz, rotation_raw
    if "global_rot" in mlps:
        hg = h.mean(0, keepdim=True)
        ang = torch.sigmoid(mlp(hg, *mlps["global_rot"])) * 2 * math.pi - math.pi
        rmat = euler_to_matrix(ang.squeeze(0))
        trans = torch.sigmoid(mlp(hg, *mlps["global_trans"])) * 2 - 1
        base_rot = matrix_to_quat(rmat.to(xyz) @ quat_to_matrix(rotation_raw))
        base_xyz = (rmat.to(xyz) @ xyz.T).T + trans
    means = base_xyz + mlp(h, *mlps["xyz"])
    sc = scaling_raw + (mlp(h, *mlps["scale"]) if deform_scale else 0.0)
    rot = base_rot + mlp(h, *mlps["rot"])
    return means, torch.exp(sc), F.normalize(rot, dim=-1)
