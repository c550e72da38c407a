"""Dump the per-step timeline (SM clock cycles) of one CTA of the head-dim-40 attention kernel (the v3 kernel carries the
clock64 hooks; v4/v5 share its pipeline but are not instrumented)."""
import ctypes as C
import os
import sys

os.environ.setdefault("A3D_ATTN_VARIANT", "3")

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animate3d_b200 import _lib as L
from tools import kernel_bench as kb

lib = L.load()
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
lib.a3d_debug_set_attn_trace(C.c_void_p(buf.data_ptr()))
kb.attn_case("l0 cross-view (traced)", 2, 4, 16, 1024, 40)
lib.a3d_debug_set_attn_trace(C.c_void_p(None))
t = buf.cpu().view(64, 16)
t0 = int(t[0, 4])
names = {0: "m:s_free", 10: "m:k_ready", 1: "m:qk_iss", 2: "m:p_full", 11: "m:v_ready", 3: "m:pv_iss", 12: "m:commits",
         4: "s:wait_s", 5: "s:got_s", 6: "s:tmem_ld", 7: "s:exp_done", 9: "s:arrive"}
order = [0, 10, 1, 2, 11, 3, 12, 4, 5, 6, 7, 9]
print("step " + " ".join(f"{names[k]:>10s}" for k in order))
for j in list(range(0, 12)) + [30, 31, 32, 61]:
    print(f"{j:4d} " + " ".join(f"{int(t[j, k]) - t0:10d}" for k in order))
