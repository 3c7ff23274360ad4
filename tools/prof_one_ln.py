"""One LayerNorm+residual forward + backward on a stage shape, for ncu: python tools/prof_one_ln.py [stage]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops
stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T, C = [(48 * 14400, 128), (48 * 3600, 256), (48 * 900, 512), (48 * 225, 1024)][stage]
dev = "cuda"
x = torch.randn(T, C, device=dev).bfloat16().requires_grad_(True)
r = torch.randn(T, C, device=dev).bfloat16()
dy = torch.randn(T, C, device=dev).bfloat16()
g = torch.ones(C, device=dev, requires_grad=True); b = torch.zeros(C, device=dev, requires_grad=True)
pb = torch.zeros(C, device=dev, requires_grad=True)
for _ in range(3):
    y = ops.layer_norm_residual(x, g, b, 1e-6, residual=r, producer_bias=pb)
    torch.autograd.grad(y, [x, g, b, pb], dy)
torch.cuda.synchronize()
print("ok")
