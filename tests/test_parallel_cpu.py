"""Host-side logic of the N>1 path on CPU with the gloo backend, world_size 2 (no GPU needed)."""
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["A3D_ROOT"])
from animate3d_b200.parallel import shard_units, gather_latents, max_over_ranks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_units(5, rank, world)
local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1).expand(hi - lo, 2, 3).contiguous()
parts = gather_latents(local, world)
full = torch.cat(parts)
assert full.shape[0] == 5 and torch.equal(full[:, 0, 0], torch.arange(5.0)), full
m = max_over_ranks(10.0 + rank, "cpu", world)
assert m == 10.0 + world - 1
dist.barrier()
if rank == 0:
    print("OK", lo, hi)
dist.destroy_process_group()
'''


def test_shard_units_partition():
    from animate3d_b200.parallel import shard_units
    for n in (1, 5, 8, 13):
        for w in (1, 2, 4, 8):
            spans = [shard_units(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_gather_and_max(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, A3D_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK 0 3" in r.stdout


def test_bench_reference_arm_under_torchrun_rank0_only(tmp_path):
    """bench.py --impl reference with 2 ranks: rank 0 prints the JSON line, the other rank exits 0 without work."""
    env = dict(os.environ, A3D_BENCH_TINY="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--impl", "reference"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    import json
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


CAM_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["A3D_ROOT"])
from animate3d_b200.parallel import shard_cameras, allreduce_gradients, gather_renders, scatter_render_grads
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
bs = 7
batch = {"c2w": torch.arange(bs * 16, dtype=torch.float32).reshape(bs, 4, 4), "fovy": torch.arange(bs, dtype=torch.float32),
         "timestamps": torch.linspace(-1, 1, bs), "width": 8, "height": 8, "do_guidance": True}
sub = shard_cameras(batch, rank, world)
assert sub["camera_index"].tolist() == list(range(rank, bs, world)) and sub["width"] == 8
assert torch.equal(sub["fovy"], batch["fovy"][rank::world]) and torch.equal(sub["c2w"], batch["c2w"][rank::world])
# a toy "renderer": image_i = w * fovy_i ; loss = sum_i image_i * (i + 1)  ->  dL/dw = sum_i fovy_i * (i + 1)
w = torch.nn.Parameter(torch.tensor([2.0, -1.0]))
unused = torch.nn.Parameter(torch.zeros(3))
local = (w[None, :] * sub["fovy"][:, None])                                   # [n_local, 2]
full = gather_renders(local, rank, world, bs)
assert torch.allclose(full, torch.tensor([2.0, -1.0])[None] * batch["fovy"][:, None])
full_grad = (torch.arange(bs, dtype=torch.float32) + 1)[:, None].expand(bs, 2).contiguous()
local.backward(scatter_render_grads(full_grad, rank, world))
nbytes = allreduce_gradients([w, unused], world)
want = (batch["fovy"] * (torch.arange(bs) + 1)).sum()
assert nbytes == 5 * 4 and torch.allclose(w.grad, torch.stack([want, want])), (w.grad, want)
assert unused.grad is not None and unused.grad.abs().sum() == 0
dist.barrier()
if rank == 0:
    print("CAM_OK")
dist.destroy_process_group()
'''


def test_gloo_world2_camera_sharding_and_grad_allreduce(tmp_path):
    """Rasterizer sharding of SURVEY 8e: cameras r::N per rank, forward all-gather of the renders, ONE flat all-reduce of
    the deformation-field gradients -- equals the single-process gradient."""
    script = tmp_path / "cam_worker.py"
    script.write_text(CAM_WORKER)
    env = dict(os.environ, A3D_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29535", str(script)], env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "CAM_OK" in r.stdout, r.stdout + r.stderr
