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
