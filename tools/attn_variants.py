"""Head-dim-40 cross-view attention: accuracy and time of the kernel's switchable modes (a3d_debug_set_attn_poly flags: bit 0 =
one-step-ahead barrier tests, bit 1 = probabilities through tensor memory / TS-mode P V product, bit 2 = one query tile per CTA and
two CTAs per SM, bit 3 = TS: no explicit P-buffer wait, implied by s_full) at the bench shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animate3d_b200 import _lib as L
from animate3d_b200 import ops
from tools import kernel_bench as kb

lib = L.load()
heads, d, dqk, dv = 8, 40, 48, 48


def accuracy(kind):
    g = torch.Generator(device="cuda").manual_seed(5)
    Lk = 4096
    q = torch.randn(2 * Lk, heads, d, device="cuda", generator=g).half().float()
    k = torch.randn(2 * Lk, heads, d, device="cuda", generator=g).half().float()
    v = torch.randn(2 * Lk, heads, d, device="cuda", generator=g).half().float()
    if kind == "wide":
        q *= 2.5
    pad = lambda x, dp: torch.nn.functional.pad(x, (0, dp - d))
    vb = pad(v, dv)
    vb[..., d] = 1.0
    buf = torch.cat([pad(q, dqk).reshape(2 * Lk, -1), pad(k, dqk).reshape(2 * Lk, -1), vb.reshape(2 * Lk, -1)], 1).half().contiguous()
    ld = buf.shape[1]
    st = (ld, Lk * ld, Lk * ld, Lk * ld)
    ext = (Lk, 1, 1, 2)
    c = heads * d
    out = torch.zeros(2 * Lk, c, device="cuda", dtype=torch.float16)
    ops.attention(ops.view5(buf, 0, ld, st, ext), ops.view5(buf, heads * dqk, ld - heads * dqk, st, ext),
                  ops.view5(buf, 2 * heads * dqk, ld - 2 * heads * dqk, st, ext), out, (c, Lk * c, Lk * c, Lk * c), heads=heads, d=d,
                  scale=d ** -0.5, impl=L.IMPL_TC)
    qq = q.reshape(2, Lk, heads, d).permute(0, 2, 1, 3)
    kk = k.reshape(2, Lk, heads, d).permute(0, 2, 1, 3)
    vv = v.reshape(2, Lk, heads, d).permute(0, 2, 1, 3)
    ref = torch.einsum("bhqk,bhkd->bhqd", (torch.einsum("bhqd,bhkd->bhqk", qq, kk) * d ** -0.5).softmax(-1), vv)
    ref = ref.permute(0, 2, 1, 3).reshape(2 * Lk, c)
    torch.cuda.synchronize()
    e = out.float() - ref
    return (e.norm() / ref.norm()).item(), (e.abs().max() / ref.abs().max()).item()


for flags in (7, 15, 7, 15, 3):
    L.check(lib.a3d_debug_set_attn_poly(flags))
    r1, m1 = accuracy("randn")
    r2, m2 = accuracy("wide")
    name = f"early={flags & 1} ts={(flags >> 1) & 1} tiles/CTA={1 if flags & 4 else 2} implied-P-free={(flags >> 3) & 1}"
    print(f"{name}: randn rel-l2 {r1:.2e} max {m1:.2e} | wide rel-l2 {r2:.2e} max {m2:.2e}")
    kb.attn_case(f"l0 {name}", 2, 4, 16, 1024, 40)
L.check(lib.a3d_debug_set_attn_poly(15))
