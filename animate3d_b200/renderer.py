"""Host-side mirror of the reference's 4D batch renderer: `DiffGaussian4D.forward` (custom/threestudio-animate3d/renderer/
diff_gaussian_rasterizer_advanced_4d.py:50-192, registered "diff-gaussian-rasterizer-advanced-4d") and
`Gaussian4DBatchRenderer.batch_forward` (renderer/gaussian_batch_renderer_4d.py:11-111).

The reference loops over the `bs` cameras of a batch in Python (line 27): per camera 2 x torch.inverse, a k-planes lookup +
3 MLPs, ~8 rasterizer kernels, 2 CUB calls and a blocking D2H read.  Here one call does: camera matrices for all cameras
(threestudio/utils/ops.py:305-359), ONE deformation launch for the distinct timestamps, ONE launch per rasterizer stage for
all cameras, one host sync per batch."""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch

from .gaussian4d import Gaussian4DModel
from .rasterizer import GaussianRasterizationSettings, _RasterizeBatch, rasterize_batch


def get_cam_info_gaussian(c2w: torch.Tensor, fovx, fovy, znear: float = 0.1, zfar: float = 100.0):
    """threestudio/utils/ops.py:344-359, batched over a leading camera dimension (fovx / fovy: floats or [bs] tensors):
    returns (world_view_transform, full_proj_transform, camera_center) in the row-vector convention of the rasterizer."""
    c2w = c2w.float().reshape(-1, 4, 4)
    bs, dev = c2w.shape[0], c2w.device
    flip = torch.eye(4, device=dev)
    flip[1, 1] = -1
    flip[2, 2] = -1
    wv = torch.linalg.inv(c2w @ flip).transpose(1, 2).contiguous()
    tx = torch.tan(torch.as_tensor(fovx, dtype=torch.float32, device=dev).reshape(-1) * 0.5).expand(bs)
    ty = torch.tan(torch.as_tensor(fovy, dtype=torch.float32, device=dev).reshape(-1) * 0.5).expand(bs)
    P = torch.zeros(bs, 4, 4, device=dev)
    P[:, 0, 0] = 1.0 / tx
    P[:, 1, 1] = 1.0 / ty
    P[:, 3, 2] = 1.0
    P[:, 2, 2] = zfar / (zfar - znear)
    P[:, 2, 3] = -(zfar * znear) / (zfar - znear)
    full = wv @ P.transpose(1, 2)
    cam = torch.linalg.inv(wv)[:, 3, :3]
    return wv, full, cam, tx, ty


class Gaussian4DBatchRenderer:
    def __init__(self, geometry: Gaussian4DModel, back_ground_color=(0.5, 0.5, 0.5), first_frame_trainable: bool = True):
        self.geometry = geometry
        self.background_tensor = torch.tensor(back_ground_color, dtype=torch.float32, device=geometry._xyz.device)
        self.first_frame_trainable = first_frame_trainable
        self.training = False

    def batch_forward(self, batch: Dict) -> Dict:
        """batch: c2w [bs,4,4], fovy [bs], width, height, timestamps [bs], do_guidance, do_reconstruction
        (gaussian_batch_renderer_4d.py:11-50).  Returns the reference's output dict (72-109)."""
        pc = self.geometry
        c2w, fovy = batch["c2w"], batch["fovy"]
        bs = c2w.shape[0]
        H, W = int(batch["height"]), int(batch["width"])
        ts = batch["timestamps"].reshape(bs).float()
        # distinct timestamps -> one deformation evaluation each (the 4 views of a frame share it)
        uniq, inverse = torch.unique(ts, return_inverse=True)
        means_t, scales_t, rots_t = pc.deform_all(uniq, deform_scale=bool(batch.get("do_guidance", True)))
        if not self.first_frame_trainable:
            first = (uniq == -1)
            if first.any():     # static gaussians for the condition frame (diff_gaussian_rasterizer_advanced_4d.py:79-83)
                keep = first[:, None, None]
                means_t = torch.where(keep, pc._xyz[None], means_t)
                scales_t = torch.where(keep, torch.exp(pc._scaling)[None], scales_t)
                rots_t = torch.where(keep, torch.nn.functional.normalize(pc._rotation, dim=-1)[None], rots_t)
        means, scales, rots = means_t[inverse], scales_t[inverse], rots_t[inverse]
        if not batch.get("do_reconstruction", True):
            means = means.detach()
        wv, full, cam, tx, ty = get_cam_info_gaussian(c2w, fovy, fovy)
        cams_t = torch.cat([wv.reshape(bs, 16), full.reshape(bs, 16), cam, tx[:, None], ty[:, None]], dim=1).contiguous()
        m2 = torch.zeros(bs, means.shape[1], 3, device=means.device, requires_grad=True)
        meta = (H, W, int(pc.active_sh_degree), True, 1.0, [float(x) for x in self.background_tensor.tolist()])
        color, radii, depth, alpha = _RasterizeBatch.apply(means, m2, scales, rots, pc.get_opacity, pc._features_dc.reshape(-1, 1, 3),
                                                           None, cams_t, meta)
        return {"comp_rgb": color.clamp(0, 1).permute(0, 2, 3, 1), "comp_depth": depth.permute(0, 2, 3, 1),
                "comp_mask": alpha.permute(0, 2, 3, 1), "viewspace_points": m2, "visibility_filter": [r > 0 for r in radii],
                "radii": list(radii), "means3D": list(means), "scales": list(scales), "rotations": list(rots),
                "opacities": [pc.get_opacity] * bs}
