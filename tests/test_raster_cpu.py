"""The rasterizer arithmetic shared by the CUDA kernels (animate3d_b200/csrc/a3d_raster_math.h), driven serially on the
CPU through tests/cpu_harness/raster_cpu.cpp, against the oracle (oracle/raster_oracle.py): tile/bin indices bit-exact,
images and all gradients (torch autograd through the oracle) within fp32 tolerance.  CPU-only."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import raster_oracle as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("h") / "raster_cpu.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC",
                           os.path.join(ROOT, "tests", "cpu_harness", "raster_cpu.cpp"), "-o", out])
    lib = C.CDLL(out)
    lib.raster_cpu_forward.restype = C.c_long
    return lib


def fp(a):
    return a.ctypes.data_as(C.c_void_p)


def run_harness(lib, xyz, s, q, o, col, wv, full, tf, H, W, bg, grads=None):
    P = xyz.shape[0]
    f = lambda t: np.ascontiguousarray(t.detach().numpy().astype(np.float32))
    xyz_, s_, q_, o_, col_, wv_, full_ = f(xyz), f(s), f(q), f(o), f(col), f(wv), f(full)
    bg_ = f(bg)
    gx, gy = (W + 15) // 16, (H + 15) // 16
    color = np.zeros((3, H, W), np.float32); depth = np.zeros((H, W), np.float32); alpha = np.zeros((H, W), np.float32)
    radii = np.zeros(P, np.int32); ncon = np.zeros((H, W), np.int32); fT = np.zeros((H, W), np.float32)
    cap = P * gx * gy
    keys = np.zeros(cap, np.uint64); vals = np.zeros(cap, np.uint32); ranges = np.zeros((gx * gy, 2), np.int64)
    pre = np.zeros((P, 8), np.float32)
    n = lib.raster_cpu_forward(P, fp(xyz_), fp(s_), fp(q_), fp(o_), fp(col_), fp(wv_), fp(full_), C.c_float(tf), C.c_float(tf), H, W,
                               fp(bg_), C.c_float(1.0), fp(color), fp(depth), fp(alpha), fp(radii), fp(ncon), fp(fT), fp(keys),
                               fp(vals), C.c_long(cap), fp(ranges), fp(pre))
    res = dict(color=color, depth=depth, alpha=alpha, radii=radii, ncon=ncon, fT=fT, keys=keys[:n], vals=vals[:n], ranges=ranges, pre=pre)
    if grads is not None:
        dC, dD, dA = [np.ascontiguousarray(g.numpy().astype(np.float32)) for g in grads]
        gm = np.zeros((P, 3), np.float32); gs = np.zeros((P, 3), np.float32); gr = np.zeros((P, 4), np.float32)
        go = np.zeros(P, np.float32); gc = np.zeros((P, 3), np.float32)
        lib.raster_cpu_backward(P, fp(xyz_), fp(s_), fp(q_), fp(o_), fp(col_), fp(wv_), fp(full_), C.c_float(tf), C.c_float(tf), H, W,
                                fp(bg_), C.c_float(1.0), fp(vals), fp(ranges), fp(ncon), fp(fT), fp(dC), fp(dD), fp(dA),
                                fp(gm), fp(gs), fp(gr), fp(go), fp(gc))
        res.update(gm=gm, gs=gs, gr=gr, go=go, gc=gc)
    return res


@pytest.mark.parametrize("P,H,W,seed", [(300, 48, 64, 0), (1500, 64, 64, 1), (800, 40, 72, 2)])
def test_forward_indices_bit_exact_and_image_close(harness, P, H, W, seed):
    xyz, s, q, o, sh = R.synthetic_scene(P, seed)
    s = s * 3.0                                   # larger splats -> several tiles per gaussian
    (c2w, fovy) = R.random_cameras(4, seed)[seed % 4]
    wv, full, cp = R.get_cam_info_gaussian(c2w, fovy, fovy)
    tf = math.tan(fovy / 2)
    bg = torch.tensor([0.5, 0.2, 0.8])
    col = torch.clamp(R.SH_C0 * sh[:, 0] + 0.5, min=0)
    ora = R.rasterize(xyz, s, q, o, None, col, 0, wv, full, cp, tf, tf, H, W, bg)
    got = run_harness(harness, xyz, s, q, o, col, wv, full, tf, H, W, bg)
    assert np.array_equal(got["radii"], ora["radii"].numpy())
    assert np.array_equal(got["keys"], ora["keys"]), "sorted (tile | depth) keys must be bit-exact"
    assert np.array_equal(got["vals"].astype(np.int64), ora["point_list"])
    assert np.array_equal(got["ranges"], ora["ranges"])
    assert len(got["keys"]) > P // 2
    np.testing.assert_allclose(got["color"], ora["color"].numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(got["depth"], ora["depth"][0].numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(got["alpha"], ora["alpha"][0].numpy(), rtol=1e-4, atol=2e-5)
    assert np.array_equal(got["ncon"], ora["n_contrib"].numpy().astype(np.int32))


@pytest.mark.parametrize("P,H,W,seed", [(200, 32, 48, 3), (600, 48, 48, 4)])
def test_backward_matches_autograd_through_oracle(harness, P, H, W, seed):
    xyz, s, q, o, sh = R.synthetic_scene(P, seed)
    s = s * 3.0
    (c2w, fovy) = R.random_cameras(4, seed)[1]
    wv, full, cp = R.get_cam_info_gaussian(c2w, fovy, fovy)
    tf = math.tan(fovy / 2)
    bg = torch.tensor([0.5, 0.5, 0.5])
    col = torch.clamp(R.SH_C0 * sh[:, 0] + 0.5, min=0)
    leaves = [t.clone().requires_grad_(True) for t in (xyz, s, q, o, col)]
    ora = R.rasterize(leaves[0], leaves[1], leaves[2], leaves[3], None, leaves[4], 0, wv, full, cp, tf, tf, H, W, bg)
    g = torch.Generator().manual_seed(seed)
    dC = torch.randn(3, H, W, generator=g); dD = torch.randn(H, W, generator=g) * 0.3; dA = torch.randn(H, W, generator=g)
    loss = (ora["color"] * dC).sum() + (ora["depth"][0] * dD).sum() + (ora["alpha"][0] * dA).sum()
    loss.backward()
    got = run_harness(harness, xyz, s, q, o, col, wv, full, tf, H, W, bg, grads=(dC, dD, dA))
    for name, ref, mine in (("means3D", leaves[0].grad, got["gm"]), ("scales", leaves[1].grad, got["gs"]),
                            ("rotations", leaves[2].grad, got["gr"]), ("opacity", leaves[3].grad[:, 0], got["go"]),
                            ("colors", leaves[4].grad, got["gc"])):
        ref = ref.numpy()
        scale = np.abs(ref).max() + 1e-12
        err = np.abs(mine - ref).max() / scale
        assert err < 2e-3, f"grad {name}: max err / max ref = {err:.3e}"


def test_cam_info_matches_reference_source(golden_dir):
    """Camera -> (world_view, full_proj, centre) of threestudio/utils/ops.py:305-359: the oracle and the product's batched
    version against outputs of the reference functions themselves."""
    import os

    import torch

    from animate3d_b200.renderer import get_cam_info_gaussian as batched
    from oracle import raster_oracle as R
    cams = torch.load(os.path.join(golden_dir, "ref_cam_info.pt"), weights_only=False)
    for c in cams:
        wv, full, centre = R.get_cam_info_gaussian(c["c2w"], c["fovx"], c["fovy"], 0.1, 100.0)
        torch.testing.assert_close(wv, c["wv"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(full, c["full"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(centre, c["center"], rtol=1e-5, atol=1e-6)
    c2w = torch.stack([c["c2w"] for c in cams])
    wv, full, centre, tx, ty = batched(c2w, torch.tensor([c["fovx"] for c in cams]), torch.tensor([c["fovy"] for c in cams]))
    torch.testing.assert_close(wv, torch.stack([c["wv"] for c in cams]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(full, torch.stack([c["full"] for c in cams]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(centre, torch.stack([c["center"] for c in cams]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(tx, torch.tan(torch.tensor([c["fovx"] for c in cams]) * 0.5))
