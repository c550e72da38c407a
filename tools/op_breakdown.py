"""Per-shape time breakdown of one eager CFG denoise step: every a3d op call is bracketed with CUDA events (launch gaps do
not count) and aggregated by (op, shape signature).  Answers "which GEMM shapes carry the step"."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animate3d_b200 import ops

records = []


def wrap(name, sig):
    fn = getattr(ops, name)

    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        records.append((name, sig(*a, **k), e0, e1))
        return r

    setattr(ops, name, timed)


def gemm_sig(A, B, out, *, M, N, K, conv=None, geglu=False, R1=None, R2=None, rowbias=None, out_f32=False, **kw):
    flops = 2.0 * M * N * K
    return (f"M={M} N={N} K={K}" + (f" conv{conv[1]}x{conv[2]}s{conv[4]}" if conv else "") + (" geglu" if geglu else "")
            + (" +R1" if R1 is not None else "") + (" +R2" if R2 is not None else "") + (" rowbias" if rowbias is not None else "")
            + (" f32" if out_f32 else ""), flops)


def attn_sig(q, k, v, out, ostrides, *, heads, d, **kw):
    lq, lk = q.e1 * q.e2, k.e1 * k.e2
    b = q.e3 * q.e4
    return (f"d={d} Lq={lq} Lk={lk} batches={b}", 4.0 * b * heads * lq * lk * d)


wrap("gemm", gemm_sig)
wrap("attention", attn_sig)
wrap("temporal_attn", lambda qkv, out, pixels, frames, heads, d, scale, **kw: (f"pixels={pixels} F={frames} d={d}", 4.0 * pixels * heads * frames * frames * d))
wrap("group_norm", lambda x1, c1, x2, c2, *a, **k: (f"c={c1 + c2} samples={a[3]} rows={a[4]}", 0.0))
wrap("layer_norm", lambda x, g, b, y, rows, c, eps=1e-5: (f"rows={rows} c={c}", 0.0))
wrap("conv_in", lambda sample, w, b, y, bn, cin, f, h, wd, cout: (f"{cin}->{cout} pixels={bn * f * h * wd}", 2.0 * bn * f * h * wd * cout * cin * 9))
wrap("conv_out", lambda x, w, b, y, bn, cin, f, h, wd, cout: (f"{cin}->{cout} pixels={bn * f * h * wd}", 2.0 * bn * f * h * wd * cout * cin * 9))
wrap("upsample2x", lambda x, y, n, h, w, c: (f"n={n} {h}x{w} c={c}", 0.0))

os.environ["A3D_STEPS"] = "2"
import tools.one_step  # noqa: E402  (runs two eager steps with the wrapped ops)

torch.cuda.synchronize()
records = records[len(records) // 2:]   # first step pays lazy module loading / tensor-map creation: keep the second
agg = collections.OrderedDict()
for name, (sig, fl), e0, e1 in records:
    key = (name, sig)
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)
    a[2] += fl
tot = sum(a[1] for a in agg.values())
print(f"total bracketed time {tot:.2f} ms over {len(records)} calls")
by_op = collections.defaultdict(float)
for (name, sig), (n, ms, fl) in agg.items():
    by_op[name] += ms
print("  ".join(f"{k}={v:.2f}ms" for k, v in by_op.items()))
for (name, sig), (n, ms, fl) in sorted(agg.items(), key=lambda x: -x[1][1])[:60]:
    tf = f"{fl / ms / 1e9:7.1f} TF/s" if fl else " " * 12
    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  n={n:3d}  {ms / n * 1e3:8.1f} us/call {tf}  {name:14s} {sig}")
