"""ORACLE (test infrastructure) -- ARAP regulariser of the 4D-SDS / reconstruction step, as in
custom/threestudio-animate3d/systems/util.py: 36-44 (produce_edge_matrix_nfmt), 58-117 (cal_connectivity_from_points, 'nn'
mode), 138-173 (estimate_rotation), 183-215 (cal_arap_error); called from systems/animate3d.py:215-244.

KNN: the reference calls pytorch3d.ops.knn_points (pinned nowhere, not installed here); restated as its published
contract -- squared distances, K nearest sorted ascending, ties by index order (brute force below).

PINNED: every function is checked against the reference's own source executed on the CPU (`.cuda()` dropped, the KNN call
served by the same brute-force contract) in tests/golden/gen_reference_goldens.py::run_reference_arap -> ref_arap.pt."""
from __future__ import annotations

import torch


def knn_points(p: torch.Tensor, K: int):
    """[N,3] -> (squared distances [N,K], indices [N,K]) of the K nearest points of p to each point of p (self included)."""
    d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
    dist, idx = torch.sort(d2, dim=1, stable=True)
    return dist[:, :K], idx[:, :K]


def connectivity_from_points(points: torch.Tensor, radius: float = 0.1, K: int = 10, least_edge_num: int = 3):
    """util.py:58-117, mode 'nn', adaptive weighting.  points [Nt, Nv, 3] (frame 0 defines the graph; an edge survives only
    if its neighbour stays within `radius` of the frame-0 position of the node in EVERY later frame).  Returns
    (ii, jj, nn, weight [Nv, K]).  NOTE: as in the reference, a dropped edge (distance set to inf) makes `nn_dist.mean()`
    infinite and the adaptive weights NaN; every shipped config has K == least_edge_num == 3 and passes frame 0 only
    (systems/animate3d.py:236-239), where no edge is ever dropped."""
    nv = points.shape[1]
    nn_dist, nn_idx = knn_points(points[0], K + 1)
    nn_dist, nn_idx = nn_dist[:, 1:].clone(), nn_idx[:, 1:].clone()              # drop self
    if points.shape[0] > 1:
        rest = points[1:][:, nn_idx]                                             # [Nt-1, Nv, K, 3]
        rest_d = ((rest - points[0:1][:, :, None]) ** 2).sum(-1)
        nn_dist = torch.where((rest_d < radius ** 2).all(0), nn_dist, torch.full_like(nn_dist, float("inf")))
    far = nn_dist[:, least_edge_num:] >= radius ** 2
    nn_idx[:, least_edge_num:] = torch.where(far, torch.full_like(nn_idx[:, least_edge_num:], -1), nn_idx[:, least_edge_num:])
    nn_dist[:, least_edge_num:] = torch.where(far, torch.full_like(nn_dist[:, least_edge_num:], float("inf")),
                                              nn_dist[:, least_edge_num:])
    weight = torch.exp(-nn_dist / nn_dist.mean())
    weight = weight / weight.sum(dim=-1, keepdim=True)
    ii = torch.arange(nv)[:, None].expand(nv, K).reshape(-1)
    jj = nn_idx.reshape(-1)
    nn = torch.arange(K)[None].expand(nv, K).reshape(-1)
    keep = jj != -1
    return ii[keep], jj[keep], nn[keep], weight


def edge_matrix(verts: torch.Tensor, nv: int, K: int, ii, jj, nn) -> torch.Tensor:
    """util.py:36-44: E[i, n] = p_i - p_(J[n]); absent edges stay zero."""
    e = torch.zeros(nv, K, 3, dtype=verts.dtype)
    e[ii, nn] = verts[ii] - verts[jj]
    return e


def estimate_rotation(source, target, ii, jj, nn, K, weight, sample_idx=None):
    """util.py:138-173: per-node weighted Procrustes rotation source edges -> target edges (torch.svd, reflection fixed by
    flipping the column of U with the smallest singular value); nodes whose edges did not move get R = I via S = 0."""
    nv = source.shape[0]
    se, te = edge_matrix(source, nv, K, ii, jj, nn), edge_matrix(target, nv, K, ii, jj, nn)
    if sample_idx is not None:
        se, te = se[sample_idx], te[sample_idx]
    s = torch.bmm(se.permute(0, 2, 1), torch.bmm(torch.diag_embed(weight), te))
    still = torch.unique(torch.where((se == te).all(dim=1))[0])
    s[still] = 0
    u, sig, w = torch.svd(s)
    r = torch.bmm(w, u.permute(0, 2, 1))
    flip = torch.nonzero(torch.det(r) <= 0, as_tuple=False).flatten()
    if len(flip) > 0:
        um = u.clone()
        cols = torch.argmin(sig[flip], dim=1)
        um[flip, :, cols] *= -1
        r[flip] = torch.bmm(w[flip], um[flip].permute(0, 2, 1))
    return r


def arap_error(nodes_sequence, ii, jj, nn, K, weight=None, sample_idx=None):
    """util.py:183-215 with the random node subset passed in (`sample_idx`; the reference draws it with
    np.random.choice(Nv, sample_num) when Nv > sample_num): sum over frames t >= 1 and sampled nodes of
    w_in |e_in(t) - R_i(t) e_in(0)|^2, R from `estimate_rotation` without gradient."""
    nt, nv, _ = nodes_sequence.shape
    if weight is None:                     # how systems/animate3d.py:241 calls it: 1 on existing edges (util.py:190-192)
        weight = torch.zeros(nv, K, dtype=nodes_sequence.dtype)
        weight[ii, nn] = 1
    if sample_idx is None:
        sample_idx = torch.arange(nv)
    src = edge_matrix(nodes_sequence[0], nv, K, ii, jj, nn)
    w = weight[sample_idx]
    err = nodes_sequence.new_zeros(())
    for t in range(1, nt):
        with torch.no_grad():
            rot = estimate_rotation(nodes_sequence[0], nodes_sequence[t], ii, jj, nn, K, w, sample_idx)
        tgt = edge_matrix(nodes_sequence[t], nv, K, ii, jj, nn)[sample_idx]
        rigid = torch.bmm(rot, src[sample_idx].permute(0, 2, 1)).permute(0, 2, 1)
        err = err + (w * ((tgt - rigid).norm(dim=2) ** 2)).sum()
    return err
