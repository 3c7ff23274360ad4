"""One fused-epilogue GEMM shape for ncu: python tools/prof_one_gemm.py [gelu|dgelu|qkv|plain] [stage]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops, _lib as L
kind = sys.argv[1] if len(sys.argv) > 1 else "gelu"
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 0
T, C = [(48 * 14400, 128), (48 * 3600, 256), (48 * 900, 512), (48 * 225, 1024)][stage]
dev = "cuda"
N, K = 4 * C, C
x = torch.randn(T, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
bias = torch.randn(N, device=dev)
z = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
dy = torch.randn(T, K, device=dev).bfloat16()
gd = torch.randn(T, N, device=dev).bfloat16()
for _ in range(3):
    if kind == "gelu":
        ops.gemm(ops.Operand(x), ops.Operand(w), T, N, K, epilogue=L.EPI_GELU, bias=bias, aux_out=z)
    elif kind == "dgelu":
        ops.gemm(ops.Operand(dy), ops.Operand(w), T, N, K, b_mn=False, epilogue=L.EPI_DGELU, aux_in=gd)
    else:
        ops.gemm(ops.Operand(x), ops.Operand(w), T, N, K, bias=bias)
torch.cuda.synchronize()
print("ok")
