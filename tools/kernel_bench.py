"""Time the hot kernels standalone at the bench workload's shapes (CUDA events, inputs >> L2 or rotated).  Prints one
line per case: achieved TFLOP/s and output GB/s.  Used to iterate on kernels without running the whole UNet."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from animate3d_b200 import ops, _lib as L

L.load()
dev = "cuda"


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_case(name, M, N, K, conv=None, geglu=False, rowbias=False, res=False, rb=(16, 1024)):
    A = torch.randn(M if conv is None else conv[0] * conv[1] * conv[2], K if conv is None else conv[3], device=dev).half()
    B = (torch.randn(N, K, device=dev) * 0.05).half()
    out = torch.empty(M, N // 2 if geglu else N, device=dev, dtype=torch.float16)
    bias = torch.randn(N, device=dev)
    rbt = torch.randn(1024, N, device=dev) if rowbias else None
    R2 = torch.randn(M, N, device=dev).half() if res else None
    fn = lambda: ops.gemm(A, B, out, M=M, N=N, K=K, conv=conv, bias=bias, geglu=geglu, rowbias=rbt, rb_div=rb[0], rb_mod=rb[1], R2=R2)
    ms = timeit(fn)
    fl = 2.0 * M * N * K
    ob = out.numel() * 2
    print(f"gemm {name:28s} M={M} N={N} K={K}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  out {ob/ms/1e6:7.1f} GB/s")


def attn_case(name, B, Nv, F, hw, d, variant=None):
    heads = 8
    dqk = (d + 15) // 16 * 16
    dv = (d + 16) // 16 * 16
    M = B * Nv * F * hw
    nq = heads * (3 * dqk + dv)
    qkv = torch.randn(M, nq, device=dev, dtype=torch.float16)
    vcol = 3 * heads * dqk
    vv_ = qkv[:, vcol:].view(M, heads, dv)
    vv_[:, :, d] = 1.0
    vv_[:, :, d + 1:] = 0.0
    c = heads * d
    out = torch.empty(M, c, device=dev, dtype=torch.float16)
    st = (nq, F * hw * nq, hw * nq, Nv * F * hw * nq)
    ext = (hw, Nv, F, B)
    ostr = (c, F * hw * c, hw * c, Nv * F * hw * c)
    vq = ops.view5(qkv, 0, nq, st, ext)
    vk = ops.view5(qkv, 2 * heads * dqk, nq - 2 * heads * dqk, st, ext)
    vv = ops.view5(qkv, vcol, nq - vcol, st, ext)
    fn = lambda: ops.attention(vq, vk, vv, out, ostr, heads=heads, d=d, scale=d ** -0.5)
    ms = timeit(fn)
    fl = B * F * 4.0 * (Nv * hw) ** 2 * c
    print(f"attn {name:28s} L={Nv*hw} d={d} batches={B*F}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    M0 = 131072
    if which in ("all", "gemm"):
        gemm_case("l0 conv3x3 320->320", M0, 320, 2880, conv=(128, 32, 32, 320, 1))
        gemm_case("l0 proj 320->320", M0, 320, 320)
        gemm_case("l0 proj+res 320->320", M0, 320, 320, res=True)
        gemm_case("l0 qkv 320->1536", M0, 1536, 320)
        gemm_case("l0 tqkv+rowbias(16 rows) 960", M0, 960, 320, rowbias=True, rb=(1, 16))
        gemm_case("l0 sqkv+rowbias(8 rows) 1152", M0, 1152, 320, rowbias=True)
        gemm_case("l0 geglu 320->2560", M0, 2560, 320, geglu=True)
        gemm_case("l0 ffout 1280->320", M0, 320, 1280, res=True)
        gemm_case("l1 conv3x3 640->640", M0 // 4, 640, 5760, conv=(128, 16, 16, 640, 1))
        gemm_case("l1 geglu 640->5120", M0 // 4, 5120, 640, geglu=True)
        gemm_case("l2 conv3x3 1280->1280", M0 // 16, 1280, 11520, conv=(128, 8, 8, 1280, 1))
        gemm_case("l2 geglu 1280->10240", M0 // 16, 10240, 1280, geglu=True)
        gemm_case("up conv3x3 960->320", M0, 320, 8640, conv=(128, 32, 32, 960, 1))
    if which == "attn0":
        attn_case("l0 cross-view", 2, 4, 16, 1024, 40)
    if which in ("all", "attn"):
        attn_case("l0 cross-view", 2, 4, 16, 1024, 40)
        attn_case("l1 cross-view", 2, 4, 16, 256, 80)
        attn_case("l2 cross-view", 2, 4, 16, 64, 160)
