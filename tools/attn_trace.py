"""Dump the per-step timeline (SM clock cycles) of one CTA of the head-dim-40 attention kernel."""
import ctypes as C
import os
import sys

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
names = {0: "mma:p_full", 1: "mma:pv_issued", 2: "mma:qk_issued", 4: "sm:wait_s", 5: "sm:got_s", 6: "sm:tmem_ld", 7: "sm:max", 8: "sm:exp_st", 9: "sm:arrive"}
print("step  " + "  ".join(f"{names[k]:>13s}" for k in sorted(names)))
for j in list(range(0, 12)) + [30, 31, 32, 62, 63]:
    print(f"{j:4d}  " + "  ".join(f"{int(t[j, k]) - t0:13d}" for k in sorted(names)))
