"""ARAP regulariser (a "next" row): the oracle restatement against the reference's own functions
(systems/util.py) executed on the CPU -- tests/golden/gen_reference_goldens.py::run_reference_arap.  No product kernel yet."""
import os

import torch


def test_arap_oracle_matches_reference_source(golden_dir):
    from oracle import arap_oracle as A
    d = torch.load(os.path.join(golden_dir, "ref_arap.pt"), weights_only=False)
    nodes, K = d["nodes"], d["K"]
    ii, jj, nn, w = A.connectivity_from_points(nodes[:1], radius=d["radius"], K=K)
    assert torch.equal(ii, d["ii"]) and torch.equal(jj, d["jj"]) and torch.equal(nn, d["nn"])      # indices: bit-exact
    torch.testing.assert_close(w, d["weight"], rtol=1e-6, atol=1e-7)
    rot = A.estimate_rotation(nodes[0], nodes[2], ii, jj, nn, K, w)
    torch.testing.assert_close(rot, d["rot_0_2"], rtol=1e-5, atol=1e-6)
    assert (torch.det(rot) > 0).all()
    for key, idx in (("all", None), ("sub", d["sample_idx"])):
        x = nodes.clone().requires_grad_(True)
        err = A.arap_error(x, ii, jj, nn, K, None, idx)
        err.backward()
        torch.testing.assert_close(err.detach(), d[f"err_{key}"], rtol=1e-5, atol=1e-7)
        torch.testing.assert_close(x.grad, d[f"grad_{key}"], rtol=1e-4, atol=1e-6)


def test_arap_is_zero_for_rigid_motion():
    from oracle import arap_oracle as A
    g = torch.Generator().manual_seed(2)
    p = torch.rand(60, 3, generator=g)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 0] = -q[:, 0]
    seq = torch.stack([p, p @ q.T + 0.3, p @ (q @ q).T - 0.1])
    ii, jj, nn, w = A.connectivity_from_points(seq[:1], radius=0.01, K=3)
    assert float(A.arap_error(seq, ii, jj, nn, 3)) < 1e-9
