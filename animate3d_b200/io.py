"""On-disk formats either side of the hot path, host code only (one-time loads, numpy):

* 3DGS PLY (`x y z nx ny nz f_dc_0..2 [f_rest_*] opacity scale_0..2 rot_0..3`, float32, binary little endian or ascii) as
  written by `tools/mesh_animation/mesh2gaussian.py:93-139` and read by `Gaussian4DModel.load_ply`
  (`custom/threestudio-animate3d/geometry/gaussian_4d.py:177-306`), including that loader's rotate / scale step
  (`load_ply_cfg.rot_x_degree / rot_z_degree / scale_factor`): positions are rotated by Rz Rx and scaled, log-scales get
  + log(scale_factor), and every gaussian's orientation is left-multiplied by the same matrix (quaternion -> matrix ->
  quaternion through scipy, exactly like the reference's `extract_rotation_scipy`, geometry/utils.py:63-71).
* per-frame mesh trajectories `mesh_trajectory/{i}.npy` (animate3d.py:465-471): plain `np.load`, see `load_mesh_trajectory`.

The reference uses the `plyfile` package (not installed here); the reader below parses the header itself."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
              "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
              "float32": "f4", "float64": "f8"}


def gaussian_attribute_names(n_rest: int = 0) -> List[str]:
    """mesh2gaussian.py:93-106 `construct_list_of_attributes`."""
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(n_rest)]
    return names + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def read_ply_vertices(path: str) -> Dict[str, np.ndarray]:
    """First element (`vertex`) of a PLY file as {property name: 1-D array}.  Scalar properties only (what 3DGS writes)."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, count, props, in_first = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                if count is None:
                    count, in_first = int(tok[2]), True
                else:
                    in_first = False
            elif tok[0] == "property" and in_first:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not supported in the vertex element")
                if tok[1] not in _PLY_TYPES:
                    raise ValueError(f"{path}: unknown property type {tok[1]}")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt is None or count is None or not props:
            raise ValueError(f"{path}: incomplete PLY header")
        if fmt == "ascii":
            data = np.loadtxt(f, max_rows=count, ndmin=2)
            if data.shape != (count, len(props)):
                raise ValueError(f"{path}: expected {count} x {len(props)} values, got {data.shape}")
            return {name: data[:, i].astype(t) for i, (name, t) in enumerate(props)}
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unsupported PLY format {fmt}")
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(name, order + t) for name, t in props])
        raw = f.read(count * dt.itemsize)
        if len(raw) != count * dt.itemsize:
            raise ValueError(f"{path}: vertex data truncated ({len(raw)} of {count * dt.itemsize} bytes)")
        arr = np.frombuffer(raw, dtype=dt, count=count)
        return {name: np.ascontiguousarray(arr[name]) for name, _ in props}


def write_gaussian_ply(path: str, xyz, f_dc, opacity, scale, rotation, f_rest=None) -> None:
    """Binary little-endian 3DGS PLY with the reference's attribute order (mesh2gaussian.py:112-139); normals are zero."""
    xyz = np.asarray(xyz, np.float32)
    n = xyz.shape[0]
    f_rest = np.zeros((n, 0), np.float32) if f_rest is None else np.asarray(f_rest, np.float32).reshape(n, -1)
    cols = np.concatenate([xyz, np.zeros_like(xyz), np.asarray(f_dc, np.float32).reshape(n, 3), f_rest,
                           np.asarray(opacity, np.float32).reshape(n, 1), np.asarray(scale, np.float32).reshape(n, 3),
                           np.asarray(rotation, np.float32).reshape(n, 4)], axis=1).astype("<f4")
    names = gaussian_attribute_names(f_rest.shape[1])
    assert cols.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n" + "".join(f"property float {a}\n" for a in names)
    with open(path, "wb") as f:
        f.write((header + "end_header\n").encode("ascii"))
        f.write(np.ascontiguousarray(cols).tobytes())


def _sorted_props(v: Dict[str, np.ndarray], prefix: str) -> List[str]:
    names = [k for k in v if k.startswith(prefix)]
    return sorted(names, key=lambda s: int(s.split("_")[-1]))


def load_gaussian_ply(path: str, rot_x_degree: float = 0.0, rot_z_degree: float = 0.0, scale_factor: float = 1.0,
                      max_sh_degree: int = 0) -> Dict[str, np.ndarray]:
    """What `Gaussian4DModel.load_ply` (gaussian_4d.py:177-306) puts into `_xyz`, `_features_dc` [P,1,3], `_features_rest`,
    `_opacity` [P,1], `_scaling` [P,3] (log) and `_rotation` [P,4] (w,x,y,z), as float32 numpy arrays."""
    from scipy.spatial.transform import Rotation   # the reference's own choice for matrix -> quaternion (utils.py:63-71)
    v = read_ply_vertices(path)
    tx, tz = np.deg2rad(rot_x_degree), np.deg2rad(rot_z_degree)
    rx = np.array([[1, 0, 0], [0, np.cos(tx), -np.sin(tx)], [0, np.sin(tx), np.cos(tx)]])
    rz = np.array([[np.cos(tz), -np.sin(tz), 0], [np.sin(tz), np.cos(tz), 0], [0, 0, 1]])
    rmat = rz @ rx
    xyz = np.stack([v["x"], v["y"], v["z"]], axis=1).astype(np.float64)
    xyz = (rmat @ xyz.T).T * scale_factor
    opacity = np.asarray(v["opacity"], np.float64)[:, None]
    f_dc = np.stack([v["f_dc_0"], v["f_dc_1"], v["f_dc_2"]], axis=1).astype(np.float64)[:, :, None]       # [P,3,1]
    out: Dict[str, np.ndarray] = {}
    if max_sh_degree > 0:
        rest = _sorted_props(v, "f_rest_")
        if len(rest) != 3 * (max_sh_degree + 1) ** 2 - 3:
            raise ValueError(f"{path}: {len(rest)} f_rest_* properties, expected {3 * (max_sh_degree + 1) ** 2 - 3}")
        fr = np.stack([v[nm] for nm in rest], axis=1).astype(np.float64).reshape(xyz.shape[0], 3, (max_sh_degree + 1) ** 2 - 1)
        out["_features_rest"] = np.ascontiguousarray(fr.transpose(0, 2, 1)).astype(np.float32)
    else:
        out["_features_rest"] = np.zeros((xyz.shape[0], 0, 3), np.float32)    # features_dc[:, :, 1:] transposed: empty
    scales = np.stack([v[nm] for nm in _sorted_props(v, "scale_")], axis=1).astype(np.float64)
    scales = np.log(np.exp(scales) * scale_factor)
    rots = np.stack([v[nm] for nm in _sorted_props(v, "rot")], axis=1).astype(np.float64)
    q = rots / np.linalg.norm(rots, axis=1, keepdims=True)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    m = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], axis=1).reshape(-1, 3, 3)
    quat_xyzw = Rotation.from_matrix(rmat @ m).as_quat()
    out.update({"_xyz": xyz.astype(np.float32), "_features_dc": np.ascontiguousarray(f_dc.transpose(0, 2, 1)).astype(np.float32),
                "_opacity": opacity.astype(np.float32), "_scaling": scales.astype(np.float32),
                "_rotation": quat_xyzw[:, [3, 0, 1, 2]].astype(np.float32)})
    return out


def load_mesh_trajectory(folder: str, n_frames: int) -> np.ndarray:
    """`mesh_trajectory/{i}.npy` for i in range(n_frames) (animate3d.py:465-471) -> [n_frames, V, 3] float32."""
    frames = []
    for i in range(n_frames):
        p = os.path.join(folder, f"{i}.npy")
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        frames.append(np.load(p).astype(np.float32))
    return np.stack(frames)
