"""Batch renderer (deformation field + batched rasterizer, the mirror of gaussian_batch_renderer_4d.py:11-111) against the
oracles composed per camera the way the reference's Python loop does it."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_batch_forward_matches_per_camera_oracle():
    from animate3d_b200.gaussian4d import Gaussian4DModel
    from animate3d_b200.renderer import make_renderer
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
    r = make_renderer(model, back_ground_color=(0.5, 0.5, 0.5), first_frame_trainable=True)
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


def _toy(P=900, sh_degree=0, seed=31, **kw):
    from animate3d_b200.gaussian4d import Gaussian4DModel
    from oracle import raster_oracle as R
    xyz, s, q, o, sh = R.synthetic_scene(P, seed, sh_degree)
    op_raw = torch.logit(o.clamp(1e-4, 1 - 1e-4))
    model = Gaussian4DModel(xyz, torch.log(s * 3), q, op_raw, sh[:, :1], grid_size=((12, 12, 12, 4), (24, 24, 24, 8)), seed=4,
                            features_rest=sh[:, 1:] if sh_degree else None, sh_degree=sh_degree, **kw)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for net in (model.delta_xyz_network, model.delta_rot_network, model.delta_scaling_network):
            net[1].copy_((torch.randn(net[1].shape, generator=g) * 0.05).cuda())
    return model, (xyz, s * 3, q, o, sh)


def _batch(n_views, times, H=48, W=48, seed=31, **flags):
    from oracle import raster_oracle as R
    cams = R.random_cameras(n_views, seed)
    c2w = torch.stack([c for c, _ in cams for _ in times]).cuda()
    fovy = torch.tensor([f for _, f in cams for _ in times]).cuda()
    ts = torch.tensor(list(times) * n_views).cuda()
    return {"c2w": c2w, "fovy": fovy, "width": W, "height": H, "timestamps": ts, **flags}


def test_reconstruction_stage_gradient_mask_and_undeformed_scales():
    """do_guidance=False (diff_gaussian_rasterizer_advanced_4d.py:132-135, 147-160): scales are the static exp(_scaling) and
    only ~10 % of the gaussians of each camera receive gradient through means / scales / rotations."""
    from animate3d_b200.renderer import make_renderer
    model, _ = _toy(P=4000)
    r = make_renderer(model, first_frame_trainable=True).train()
    batch = _batch(2, (-0.5, 0.5), do_guidance=False, do_reconstruction=True)
    torch.manual_seed(0)
    out = r.batch_forward(batch)
    static = torch.exp(model._scaling)
    for sc in out["scales"]:
        assert torch.equal(sc, static), "reconstruction stage must render the un-deformed scales"
    for m in out["means3D"]:
        m.retain_grad()
    (out["comp_rgb"].sum() + out["comp_mask"].sum()).backward()
    vis = torch.stack(out["visibility_filter"])
    for i, m in enumerate(out["means3D"]):
        g = m.grad
        touched = (g.abs().sum(-1) > 0)
        frac = touched.float().sum() / vis[i].float().sum().clamp(min=1)
        assert 0.03 < frac < 0.2, f"camera {i}: {frac:.3f} of the visible gaussians got gradient (reference: ~0.10)"
    # guidance stage: every visible gaussian gets gradient
    out = r.batch_forward(_batch(2, (-0.5, 0.5), do_guidance=True, do_reconstruction=True))
    for m in out["means3D"]:
        m.retain_grad()
    out["comp_rgb"].sum().backward()
    frac = (out["means3D"][0].grad.abs().sum(-1) > 0).float().sum() / out["visibility_filter"][0].float().sum()
    assert frac > 0.3       # every gaussian that contributes to a pixel (occluded ones have radius > 0 but no gradient)
    # do_reconstruction=False: means are detached at the rasterizer input (line 161), scales / rotations are not
    model.zero_grad()
    out = r.batch_forward(_batch(2, (-0.5, 0.5), do_guidance=True, do_reconstruction=False))
    for m in out["means3D"]:
        m.retain_grad()
    out["comp_rgb"].sum().backward()
    assert all(m.grad is None or m.grad.abs().sum() == 0 for m in out["means3D"])
    assert model.delta_scaling_network[1].grad.abs().sum() > 0


def test_viewspace_points_are_per_camera_tensors_with_grad():
    from animate3d_b200.renderer import make_renderer
    model, _ = _toy()
    r = make_renderer(model).train()
    out = r.batch_forward(_batch(2, (-1.0, 0.3), do_guidance=True, do_reconstruction=True))
    assert isinstance(out["viewspace_points"], list) and len(out["viewspace_points"]) == 4
    out["comp_rgb"].sum().backward()
    for i, v in enumerate(out["viewspace_points"]):
        assert v.shape == (model._xyz.shape[0], 3) and v.grad is not None, f"camera {i}: no screen-space gradient"
        assert v.grad[out["visibility_filter"][i]].abs().sum() > 0


def test_single_camera_forward_matches_batch_and_first_frame_is_static():
    from types import SimpleNamespace
    from animate3d_b200.renderer import get_cam_info_gaussian, make_renderer
    model, _ = _toy()
    r = make_renderer(model, first_frame_trainable=False)
    batch = _batch(1, (-1.0, 0.4), do_guidance=True, do_reconstruction=True)
    out = r.batch_forward(batch)
    # frame -1 with first_frame_trainable=False: static gaussians (lines 77-83)
    assert torch.equal(out["means3D"][0], model._xyz) and not torch.equal(out["means3D"][1], model._xyz)
    wv, full, cam, tx, ty = get_cam_info_gaussian(batch["c2w"], batch["fovy"], batch["fovy"])
    for i in range(2):
        vc = SimpleNamespace(FoVx=float(batch["fovy"][i]), FoVy=float(batch["fovy"][i]), image_width=48, image_height=48,
                             world_view_transform=wv[i], full_proj_transform=full[i], camera_center=cam[i])
        one = r.forward(vc, r.background_tensor, timestamps=batch["timestamps"], batch_idx=i, do_guidance=True, do_reconstruction=True)
        assert set(one) == {"render", "depth", "mask", "viewspace_points", "visibility_filter", "radii", "means3D", "scales",
                            "rotations", "opacities"}
        assert one["render"].shape == (3, 48, 48)
        torch.testing.assert_close(one["render"], out["comp_rgb"][i].permute(2, 0, 1), rtol=1e-5, atol=1e-6)
        assert torch.equal(one["radii"], out["radii"][i])


def test_background_inversion_and_sh_features():
    """invert_bg_prob = 0 -> every training render uses 1 - bg (65-70); SH degree 1 features go through pc.get_features."""
    import math
    from animate3d_b200.renderer import make_renderer
    from oracle import raster_oracle as R
    model, (xyz, s, q, o, sh) = _toy(P=700, sh_degree=1)
    assert model.get_features.shape == (700, 4, 3) and model.active_sh_degree == 1
    batch = _batch(1, (0.0,), do_guidance=True, do_reconstruction=True)
    batch.pop("timestamps")                                           # static render: no deformation
    bgc = (0.2, 0.4, 0.9)
    eval_img = make_renderer(model, back_ground_color=bgc, invert_bg_prob=0.0).eval().batch_forward(batch)["comp_rgb"][0]
    train_img = make_renderer(model, back_ground_color=bgc, invert_bg_prob=0.0).train().batch_forward(batch)["comp_rgb"][0]
    mixed = make_renderer(model, back_ground_color=bgc, invert_bg_prob=0.5).train()
    c2w, fovy = batch["c2w"][0].cpu(), float(batch["fovy"][0])
    wv, full, cp = R.get_cam_info_gaussian(c2w, fovy, fovy)
    tf = math.tan(fovy / 2)
    for img, bg in ((eval_img, torch.tensor(bgc)), (train_img, 1 - torch.tensor(bgc))):
        ora = R.rasterize(xyz, s, q, o, sh, None, 1, wv, full, cp, tf, tf, 48, 48, bg)
        bad = ((img.permute(2, 0, 1).cpu() - ora["color"].clamp(0, 1)).abs() > 2e-3).sum().item()
        assert bad <= 8, f"{bad} pixels differ"
    # mixed batch: some cameras inverted, some not -- every image equals one of the two single-background renders
    b8 = {**batch, "c2w": batch["c2w"].repeat(8, 1, 1), "fovy": batch["fovy"].repeat(8)}
    import numpy as np
    np.random.seed(3)
    imgs = mixed.batch_forward(b8)["comp_rgb"]
    kinds = [bool(torch.allclose(im, train_img, atol=1e-6)) for im in imgs]
    assert all(k or torch.allclose(im, eval_img, atol=1e-6) for k, im in zip(kinds, imgs)) and 0 < sum(kinds) < 8
