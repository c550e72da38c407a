"""4D deformation field: the oracle against the reference's own k-planes source (CPU), and the CUDA kernels (forward +
backward) against autograd through the oracle (GPU)."""
import os

import pytest
import torch


def test_oracle_kplanes_matches_reference_source(golden_dir):
    from oracle import gaussian4d_oracle as G
    d = torch.load(os.path.join(golden_dir, "ref_kplanes.pt"), weights_only=False)
    torch.testing.assert_close(G.interpolate_ms_features(d["pts"], d["grids"]), d["feats"], rtol=1e-6, atol=1e-6)


def test_oracle_rotation_helpers_match_reference_source(golden_dir):
    """build_rotation / extract_rotation_torch / euler_angles_to_rotation_matrix (geometry/utils.py) as restated by the
    oracle AND by the product's host module, against outputs of the reference functions themselves."""
    from animate3d_b200 import gaussian4d as H
    from oracle import gaussian4d_oracle as G
    d = torch.load(os.path.join(golden_dir, "ref_rotation.pt"), weights_only=False)
    for impl in (G, H):
        torch.testing.assert_close(impl.quat_to_matrix(d["quats"]), d["mats"], rtol=1e-6, atol=1e-6)
        for a, r in zip(d["angles"], d["rmats"]):
            torch.testing.assert_close(impl.euler_to_matrix(a), r, rtol=1e-6, atol=1e-6)
        for r, q in zip(d["rmats"], d["rotated"]):
            torch.testing.assert_close(impl.matrix_to_quat(r @ d["mats"]), q, rtol=1e-5, atol=1e-6)
    # batched host version: all frames at once
    torch.testing.assert_close(H.matrix_to_quat(d["rmats"][:, None] @ d["mats"][None]), d["rotated"], rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_deform_global_trans_matches_oracle():
    """use_global_trans (refine_frame_16.yaml:56): mean-feature kernel + global MLPs + rotated base quaternions, forward and
    all gradients (planes, the three delta MLPs, the two global MLPs) against autograd through the oracle."""
    from animate3d_b200.gaussian4d import Gaussian4DModel
    from oracle import gaussian4d_oracle as G
    from oracle import raster_oracle as R
    P, T = 2500, 4
    xyz, s, q, o, sh = R.synthetic_scene(P, 13)
    model = Gaussian4DModel(xyz, torch.log(s), q, o, sh[:, 0], grid_size=((20, 18, 22, 6), (40, 36, 44, 12)), seed=3,
                            use_global_trans=True)
    g = torch.Generator().manual_seed(6)
    all_nets = (model.delta_xyz_network, model.delta_rot_network, model.delta_scaling_network, model.global_rot_network,
                model.global_trans_network)
    with torch.no_grad():
        for net in all_nets:
            net[1].copy_((torch.randn(net[1].shape, generator=g) * 0.08).cuda())
        for pl in model.grids:
            for p in pl:
                p.copy_((torch.rand(p.shape, generator=g) * 0.8 + 0.3).cuda())
    times = torch.linspace(-1, 1, T)
    means, scales, rots = model.deform_all(times.cuda(), True)
    grids = [[p.detach().cpu().clone().requires_grad_(True) for p in pl] for pl in model.grids]
    nets = [[w.detach().cpu().clone().requires_grad_(True) for w in n] for n in all_nets]
    mlps = {"xyz": nets[0], "rot": nets[1], "scale": nets[2], "global_rot": nets[3], "global_trans": nets[4]}
    om, osc, orot = [], [], []
    for t in times.tolist():
        m_, s_, r_ = G.deform(xyz, torch.log(s), q, t, grids, mlps, True)
        om.append(m_); osc.append(s_); orot.append(r_)
    om, osc, orot = torch.stack(om), torch.stack(osc), torch.stack(orot)
    torch.testing.assert_close(means.cpu(), om, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(scales.cpu(), osc, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rots.cpu(), orot, rtol=1e-4, atol=2e-5)
    gm, gs, gr = torch.randn(om.shape, generator=g), torch.randn(osc.shape, generator=g), torch.randn(orot.shape, generator=g)
    ((om * gm).sum() + (osc * gs).sum() + (orot * gr).sum()).backward()
    ((means * gm.cuda()).sum() + (scales * gs.cuda()).sum() + (rots * gr.cuda()).sum()).backward()

    def cmp(name, mine, ref):
        err = (mine.cpu() - ref).abs().max() / (ref.abs().max() + 1e-12)
        assert err < 3e-3, f"{name}: {err:.3e}"
    for si, pl in enumerate(model.grids):
        for pi, p in enumerate(pl):
            cmp(f"grid {si}/{pi}", p.grad, grids[si][pi].grad)
    for ni, (net, ref) in enumerate(zip(all_nets, nets)):
        cmp(f"mlp{ni}.w1", net[0].grad, ref[0].grad)
        cmp(f"mlp{ni}.w2", net[1].grad, ref[1].grad)


@pytest.mark.gpu
@pytest.mark.parametrize("deform_scale", [True, False])
def test_deform_forward_backward_match_oracle(deform_scale):
    from animate3d_b200.gaussian4d import Gaussian4DModel
    from oracle import gaussian4d_oracle as G
    from oracle import raster_oracle as R
    P, T = 3000, 5
    xyz, s, q, o, sh = R.synthetic_scene(P, 11)
    model = Gaussian4DModel(xyz, torch.log(s), q, o, sh[:, 0], grid_size=((20, 18, 22, 6), (40, 36, 44, 12)), seed=3)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():   # non-trivial last layers and time planes
        for net in (model.delta_xyz_network, model.delta_rot_network, model.delta_scaling_network):
            net[1].copy_((torch.randn(net[1].shape, generator=g) * 0.05).cuda())
        for pl in model.grids:
            for p in pl:
                p.copy_((torch.rand(p.shape, generator=g) * 0.8 + 0.3).cuda())
    times = torch.linspace(-1, 1, T)
    means, scales, rots = model.deform_all(times.cuda(), deform_scale)
    # oracle with autograd
    grids = [[p.detach().cpu().clone().requires_grad_(True) for p in pl] for pl in model.grids]
    nets = [[w.detach().cpu().clone().requires_grad_(True) for w in n] for n in
            (model.delta_xyz_network, model.delta_rot_network, model.delta_scaling_network)]
    mlps = {"xyz": nets[0], "rot": nets[1], "scale": nets[2]}
    om, osc, orot = [], [], []
    for t in times.tolist():
        m_, s_, r_ = G.deform(xyz, torch.log(s), q, t, grids, mlps, deform_scale)
        om.append(m_); osc.append(s_); orot.append(r_)
    om, osc, orot = torch.stack(om), torch.stack(osc), torch.stack(orot)
    torch.testing.assert_close(means.cpu(), om, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(scales.cpu(), osc, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(rots.cpu(), orot, rtol=1e-4, atol=1e-5)
    gm, gs, gr = torch.randn(om.shape, generator=g), torch.randn(osc.shape, generator=g), torch.randn(orot.shape, generator=g)
    ((om * gm).sum() + (osc * gs).sum() + (orot * gr).sum()).backward()
    ((means * gm.cuda()).sum() + (scales * gs.cuda()).sum() + (rots * gr.cuda()).sum()).backward()

    def cmp(name, mine, ref):
        if ref is None:
            assert mine is None or mine.abs().max() == 0
            return
        err = (mine.cpu() - ref).abs().max() / (ref.abs().max() + 1e-12)
        assert err < 2e-3, f"{name}: {err:.3e}"
    for si, pl in enumerate(model.grids):
        for pi, p in enumerate(pl):
            cmp(f"grid {si}/{pi}", p.grad, grids[si][pi].grad)
    for ni, (net, ref) in enumerate(zip((model.delta_xyz_network, model.delta_rot_network, model.delta_scaling_network), nets)):
        if ni == 2 and not deform_scale:
            continue
        cmp(f"mlp{ni}.w1", net[0].grad, ref[0].grad)
        cmp(f"mlp{ni}.w2", net[1].grad, ref[1].grad)
