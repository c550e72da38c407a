"""Per-step clock64 timeline of the four softmax warps that share SM sub-partition 0 in CTA (0,0,0) of the head-dim-40
cross-view attention kernel: when each warp starts waiting for S, gets it, has it in registers, finishes the exponentials and
has published P.  Shows whether the warps of a sub-partition overlap their compute phases or idle together."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animate3d_b200 import _lib as L
from tools import kernel_bench as kb

lib = L.load()
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 15       # bit 0: early barrier tests, bit 1: TS-mode P V, bit 2: one tile per CTA, bit 3: no explicit P-buffer wait
early = flags
L.check(lib.a3d_debug_set_attn_poly(flags))
buf = torch.zeros(8 + 4 * 32 * 8 + 32, dtype=torch.int64, device="cuda")
kb.attn_case("warm", 2, 4, 16, 1024, 40)
lib.a3d_debug_set_attn_trace(C.c_void_p(buf.data_ptr()))
kb.attn_case(f"l0 cross-view flags={early} (traced)", 2, 4, 16, 1024, 40)
lib.a3d_debug_set_attn_trace(C.c_void_p(None))
t = buf[8:8 + 4 * 32 * 8].cpu().view(4, 32, 8)
cta = buf[8 + 4 * 32 * 8:].cpu().view(4, 8)
t0 = int(t[:, 0, 0].min())
names = ["top", "s_in_regs", "p_free", "exp_done", "published"]
print("warp = (tile g, key half h); columns: cycles since the first warp's step 0")
for j in list(range(4, 12)) + [20, 21, 30, 31]:
    for w in range(4):
        r = [int(x) - t0 for x in t[w, j, :5]]
        print(f"step {j:2d} warp(g={w >> 1},h={w & 1}) " + " ".join(f"{n}={v:7d}" for n, v in zip(names, r)) +
              f" | wait+ldtm {r[1] - r[0]:5d} p_free {r[2] - r[1]:5d} exp {r[3] - r[2]:5d} publish {r[4] - r[3]:5d}")
for w in range(4):
    d = t[w, 31, 4] - t[w, 8, 4]
    print(f"warp {w}: {int(d) / 23:.0f} cycles per step (steps 8..31)")
ph = {"wait+ldtm": (0, 1), "p_free": (1, 2), "exp": (2, 3), "publish": (3, 4)}
for k, (a, b) in ph.items():
    v = (t[:, 8:32, b] - t[:, 8:32, a]).float()
    print(f"{k:14s} mean {v.mean():7.0f}  min {v.min():7.0f}  max {v.max():7.0f}")
print("CTA life cycles (cycles since entry): set-up done | first scores in registers | step loop done | accumulators complete | rows stored | exit")
for k in range(4):
    c = [int(x) for x in cta[k]]
    print(f"CTA at {k}/4 of the grid on SM {c[7]:3d}: " + " | ".join(f"{c[i] - c[0]:7d}" for i in range(1, 7)))
L.check(lib.a3d_debug_set_attn_poly(15))
