"""Splat Mpix/s (BASELINE config 3): 50k synthetic gaussians, 4 views x 16 timestamps = 64 cameras at 512^2, forward and
forward+backward through the batched rasterizer (+ the deformation field), CUDA-event timed."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synthetic_model(P=50000, seed=0, device="cuda"):
    from animate3d_b200.gaussian4d import Gaussian4DModel
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(P, 3, generator=g)
    xyz = d / d.norm(dim=-1, keepdim=True) * 0.5 * torch.rand(P, 1, generator=g) ** (1 / 3)
    log_s = math.log(0.004) + (math.log(0.02) - math.log(0.004)) * torch.rand(P, 3, generator=g)
    q = torch.nn.functional.normalize(torch.randn(P, 4, generator=g), dim=-1)
    op = 2.0 + torch.randn(P, 1, generator=g)
    fdc = torch.rand(P, 3, generator=g) * 2 - 1
    m = Gaussian4DModel(xyz, log_s, q, op, fdc, seed=seed, device=device)
    with torch.no_grad():
        for net in (m.delta_xyz_network, m.delta_rot_network, m.delta_scaling_network):
            net[1].copy_((torch.randn(net[1].shape, generator=g) * 1e-3).to(device))
    return m


def cameras(n_views=4, n_frames=16, seed=0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    elev = math.radians(30 * torch.rand(1, generator=g).item())
    az0 = -180 + 90 * torch.rand(1, generator=g).item()
    fovy = math.radians(15 + 45 * torch.rand(1, generator=g).item())
    dist = (0.8 + 0.2 * torch.rand(1, generator=g).item()) / math.tan(fovy / 2)
    c2ws = []
    for v in range(n_views):
        az = math.radians(az0 + 90 * v)
        pos = torch.tensor([dist * math.cos(elev) * math.cos(az), dist * math.cos(elev) * math.sin(az), dist * math.sin(elev)])
        look = -pos / pos.norm()
        right = torch.linalg.cross(look, torch.tensor([0.0, 0.0, 1.0])); right = right / right.norm()
        up = torch.linalg.cross(right, look)
        c = torch.eye(4); c[:3, 0], c[:3, 1], c[:3, 2], c[:3, 3] = right, up, -look, pos
        c2ws += [c] * n_frames
    ts = torch.linspace(-1, 1, n_frames).repeat(n_views)
    return torch.stack(c2ws).to(device), torch.full((n_views * n_frames,), fovy, device=device), ts.to(device)


def run(P=50000, H=512, W=512, iters=5):
    from animate3d_b200.renderer import make_renderer
    model = synthetic_model(P)
    r = make_renderer(model)
    c2w, fovy, ts = cameras()
    batch = {"c2w": c2w, "fovy": fovy, "width": W, "height": H, "timestamps": ts, "do_guidance": True, "do_reconstruction": True}
    target = torch.rand(c2w.shape[0], H, W, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    res = {}
    for mode in ("fwd", "fwd_bwd"):
        for _ in range(2):
            out = r.batch_forward(batch)
            if mode == "fwd_bwd":
                (0.5 * ((out["comp_rgb"] - target) ** 2).sum()).backward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            out = r.batch_forward(batch)
            if mode == "fwd_bwd":
                (0.5 * ((out["comp_rgb"] - target) ** 2).sum()).backward()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res[mode] = {"ms_per_batch": ms, "mpix_per_s": c2w.shape[0] * H * W / ms / 1e3}
    res["config"] = {"gaussians": P, "cameras": int(c2w.shape[0]), "H": H, "W": W}
    return res


def profile_one():
    """One forward+backward between cudaProfilerStart/Stop (ncu --profile-from-start off): the launch list of a whole SDS-style
    render step -- deformation field, rasterizer stages, loss, every autograd kernel in between."""
    from animate3d_b200.renderer import make_renderer
    model = synthetic_model(50000)
    r = make_renderer(model)
    c2w, fovy, ts = cameras()
    batch = {"c2w": c2w, "fovy": fovy, "width": 512, "height": 512, "timestamps": ts, "do_guidance": True, "do_reconstruction": True}
    target = torch.rand(c2w.shape[0], 512, 512, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        out = r.batch_forward(batch)
        (0.5 * ((out["comp_rgb"] - target) ** 2).sum()).backward()
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    if "--profile" in sys.argv:
        profile_one()
    else:
        print(json.dumps(run()))
