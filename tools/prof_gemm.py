"""Time every GEMM shape of the Swin-V2-B config-2 step (48 frames of 480x480) on its own: forward (K-major
operands), dgrad (B read MN-major) and wgrad (both MN-major, split-K).  Prints ms, TFLOP/s and the
algorithmic GB/s so each launch can be put against the tensor and HBM rooflines."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops, _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=48)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--only", default="")
a = ap.parse_args()
lib = L.load()
dev = "cuda"
stages = [(a.frames * 14400, 128, 2), (a.frames * 3600, 256, 2), (a.frames * 900, 512, 18), (a.frames * 225, 1024, 2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


total = 0.0
print(f"{'stage':5s} {'op':10s} {'M':>8s} {'N':>6s} {'K':>8s} {'ms':>8s} {'TF/s':>8s} {'GB/s':>8s}  x blocks -> ms/step")
for si, (T, C, nblk) in enumerate(stages):
    for name, N, K in (("qkv", 3 * C, C), ("proj", C, C), ("fc1", 4 * C, C), ("fc2", C, 4 * C)):
        x = torch.randn(T, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        dy = torch.randn(T, N, device=dev).bfloat16()
        splits = lib.b200swin_gemm_splits(N, K, T)
        cases = {
            "fwd": lambda: ops.gemm(ops.Operand(x), ops.Operand(w), T, N, K),
            "dgrad": lambda: ops.gemm(ops.Operand(dy), ops.Operand(w), T, K, N, b_mn=True),
            "wgrad": lambda: ops.gemm(ops.Operand(dy), ops.Operand(x), N, K, T, a_mn=True, b_mn=True,
                                      out_dtype=torch.float32, splits=splits),
        }
        if name == "fc1":
            z = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
            bias = torch.randn(N, device=dev)
            cases["fwd+gelu"] = lambda: ops.gemm(ops.Operand(x), ops.Operand(w), T, N, K, epilogue=L.EPI_GELU, bias=bias,
                                                 aux_out=z)
            cases["fwd+gelu-noaux"] = lambda: ops.gemm(ops.Operand(x), ops.Operand(w), T, N, K, epilogue=L.EPI_GELU,
                                                       bias=bias)
        if name == "fc2":
            gd = torch.randn(T, K, device=dev).bfloat16()
            cases["dgrad*g'"] = lambda: ops.gemm(ops.Operand(dy), ops.Operand(w), T, K, N, b_mn=True,
                                                 epilogue=L.EPI_DGELU, aux_in=gd)
        if name == "qkv":
            inv = torch.empty(T, 2, C // 32, device=dev)
            qb, vb = torch.randn(C, device=dev), torch.randn(C, device=dev)
            cases["fwd+norm"] = lambda: ops.gemm(ops.Operand(x), ops.Operand(w), T, N, K, epilogue=L.EPI_QKV, bias=qb,
                                                 bias2=vb, inv_norm=inv, nH=C // 32)
        for kind, fn in cases.items():
            if a.only and a.only not in f"{name}.{kind}":
                continue
            ms = timeit(fn)
            fl = 2.0 * T * N * K
            by = 2.0 * (T * K + N * K + T * N)
            total += ms * nblk
            print(f"st{si}   {name + '.' + kind:10s} {T:8d} {N:6d} {K:8d} {ms:8.3f} {fl / ms / 1e9:8.1f} {by / ms / 1e6:8.1f}"
                  f"  x{nblk:2d} -> {ms * nblk:7.2f}" + (f"  splits={splits}" if kind == "wgrad" else ""))
print(f"sum over blocks: {total:.2f} ms/step")
