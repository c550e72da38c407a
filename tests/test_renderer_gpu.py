"""Batch renderer (deformation field + batched rasterizer, the mirror of gaussian_batch_renderer_4d.py:11-111) against the
oracles composed per camera the way the reference's Python loop does it."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_batch_forward_matches_per_camera_oracle():
    from animate3d_b200.gaussian4d import Gaussian4DModel
    from animate3d_b200.renderer import Gaussian4DBatchRenderer
    from oracle import gaussian4d_oracle as G
    from oracle import raster_oracle as R
    P, H, W = 1200, 64, 64
    xyz, s, q, o, sh = R.synthetic_scene(P, 21)
    s = s * 3
    op_raw = torch.logit(o.clamp(1e-4, 1 - 1e-4))
    model = Gaussian4DModel(xyz, torch.log(s), q, op_raw, sh[:, 0], grid_size=((12, 12, 12, 4), (24, 24, 24, 8)), seed=4)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for net in (model.delta_xyz_network, model.delta_rot_network, model.delta_scaling_network):
            net[1].copy_((torch.randn(net[1].shape, generator=g) * 0.02).cuda())
    cams = R.random_cameras(2, 21)
    times = [-1.0, 0.2, 1.0]
    c2w = torch.stack([c for c, _ in cams for _ in times])
    fovy = torch.tensor([f for _, f in cams for _ in times])
    ts = torch.tensor(times * len(cams))
    r = Gaussian4DBatchRenderer(model, back_ground_color=(0.5, 0.5, 0.5))
    out = r.batch_forward({"c2w": c2w.cuda(), "fovy": fovy.cuda(), "width": W, "height": H, "timestamps": ts.cuda(),
                           "do_guidance": True, "do_reconstruction": True})
    assert out["comp_rgb"].shape == (6, H, W, 3) and out["comp_mask"].shape == (6, H, W, 1)
    grids = [[p.detach().cpu() for p in pl] for pl in model.grids]
    mlps = {"xyz": [w.detach().cpu() for w in model.delta_xyz_network], "rot": [w.detach().cpu() for w in model.delta_rot_network],
            "scale": [w.detach().cpu() for w in model.delta_scaling_network]}
    bg = torch.tensor([0.5, 0.5, 0.5])
    for i in range(6):
        m_, s_, q_ = G.deform(xyz, torch.log(s), q, float(ts[i]), grids, mlps, True)
        wv, full, cp = R.get_cam_info_gaussian(c2w[i], float(fovy[i]), float(fovy[i]))
        tf = math.tan(float(fovy[i]) / 2)
        ora = R.rasterize(m_, s_, q_, torch.sigmoid(op_raw), sh[:, :1], None, 0, wv, full, cp, tf, tf, H, W, bg)
        got = out["comp_rgb"][i].permute(2, 0, 1).cpu()
        bad = ((got - ora["color"].clamp(0, 1)).abs() > 2e-3).sum().item()
        assert bad <= 12, f"camera {i}: {bad} pixels differ"
        assert (out["comp_mask"][i, ..., 0].cpu() - ora["alpha"][0]).abs().max() < 5e-3
    loss = out["comp_rgb"].sum() + out["comp_mask"].sum()
    loss.backward()
    assert model.grids[0][2].grad.abs().sum() > 0 and model.delta_xyz_network[1].grad.abs().sum() > 0
