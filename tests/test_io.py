"""On-disk formats: 3DGS PLY parsing and the reference loader's rotate / scale semantics, mesh trajectories."""
import os
import struct

import numpy as np
import pytest
import torch

from animate3d_b200 import io as A


def test_ply_bytes_roundtrip(tmp_path):
    """Writer output parsed independently with struct.unpack; reader returns the same columns (binary and ascii)."""
    rng = np.random.default_rng(1)
    n = 17
    cols = dict(xyz=rng.normal(size=(n, 3)), f_dc=rng.normal(size=(n, 3)), opacity=rng.normal(size=(n, 1)),
                scale=rng.normal(size=(n, 3)), rotation=rng.normal(size=(n, 4)), f_rest=rng.normal(size=(n, 9)))
    p = str(tmp_path / "g.ply")
    A.write_gaussian_ply(p, **cols)
    raw = open(p, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    names = [l.split()[2] for l in head.decode().splitlines() if l.startswith("property")]
    assert names == A.gaussian_attribute_names(9) and b"format binary_little_endian 1.0" in head and b"element vertex 17" in head
    vals = np.array(struct.unpack("<%df" % (n * len(names)), body)).reshape(n, len(names))
    v = A.read_ply_vertices(p)
    for i, nm in enumerate(names):
        np.testing.assert_array_equal(v[nm], vals[:, i].astype(np.float32))
    np.testing.assert_array_equal(np.stack([v["x"], v["y"], v["z"]], 1), cols["xyz"].astype(np.float32))
    np.testing.assert_array_equal(v["nx"], np.zeros(n, np.float32))
    # ascii flavour of the same table
    pa = str(tmp_path / "a.ply")
    with open(pa, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment test\nelement vertex %d\n" % n + "".join(f"property float {a}\n" for a in names)
                + "element face 0\nproperty list uchar int vertex_indices\nend_header\n")
        for r in vals:
            f.write(" ".join(repr(float(x)) for x in r) + "\n")
    va = A.read_ply_vertices(pa)
    for nm in names:
        np.testing.assert_allclose(va[nm], v[nm], rtol=0, atol=0)


def test_ply_errors(tmp_path):
    p = str(tmp_path / "bad.ply")
    open(p, "wb").write(b"plx\n")
    with pytest.raises(ValueError):
        A.read_ply_vertices(p)
    open(p, "wb").write(b"ply\nformat binary_little_endian 1.0\nelement vertex 4\nproperty float x\nend_header\n\x00\x00")
    with pytest.raises(ValueError, match="truncated"):
        A.read_ply_vertices(p)


def test_load_ply_matches_reference_loader(golden_dir):
    """`load_gaussian_ply` against the reference's own `Gaussian4DModel.load_ply` body executed on the committed fixture
    (tests/golden/gen_reference_goldens.py::run_reference_load_ply), for three rotate / scale settings."""
    ref = torch.load(os.path.join(golden_dir, "ref_ply.pt"), weights_only=False)
    ply = os.path.join(golden_dir, "ref_gaussians.ply")
    assert len(ref) == 3
    for (rx, rz, sf), want in ref.items():
        got = A.load_gaussian_ply(ply, rx, rz, sf)
        for k in ("_xyz", "_features_dc", "_opacity", "_scaling", "_rotation", "_features_rest"):
            assert tuple(got[k].shape) == tuple(want[k].shape), (k, got[k].shape, want[k].shape)
            torch.testing.assert_close(torch.from_numpy(got[k]), want[k].float(), rtol=1e-6, atol=1e-6)


def test_mesh_trajectory(tmp_path):
    for i in range(3):
        np.save(tmp_path / f"{i}.npy", np.full((5, 3), i, np.float64))
    t = A.load_mesh_trajectory(str(tmp_path), 3)
    assert t.shape == (3, 5, 3) and t.dtype == np.float32 and t[2, 0, 0] == 2.0
    with pytest.raises(FileNotFoundError):
        A.load_mesh_trajectory(str(tmp_path), 4)
