"""Driver for ncu: run the attention core forward + backward on one stage shape (no model around it)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=48)
ap.add_argument("--H", type=int, default=30)
ap.add_argument("--C", type=int, default=512)
ap.add_argument("--ws", type=int, default=12)
ap.add_argument("--shift", type=int, default=0)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--impl", default="auto")
a = ap.parse_args()
ops.ATTN_IMPL["mode"] = ops.ATTN_IMPL["bwd_mode"] = a.impl
B, H, W, C, ws = a.B, a.H, a.H, a.C, a.ws
nH = C // 32
dev = "cuda"
torch.manual_seed(0)
qkv = torch.randn(B, H, W, 3 * C, device=dev)
q, k, v = qkv.split(C, -1)
qn = torch.nn.functional.normalize(q.reshape(B, H, W, nH, 32), dim=-1).reshape(B, H, W, C)
kn = torch.nn.functional.normalize(k.reshape(B, H, W, nH, 32), dim=-1).reshape(B, H, W, C)
qkv = torch.cat([qn, kn, v], -1).bfloat16().requires_grad_(True)
inv = torch.ones(B * H * W, 2, nH, device=dev)
tab = (16 * torch.sigmoid(torch.randn((2 * ws - 1) ** 2, nH, device=dev))).requires_grad_(True)
sc = torch.full((nH,), 10.0, device=dev, requires_grad=True)
qpad = torch.nn.functional.normalize(torch.randn(nH, 32, device=dev), dim=-1).reshape(C)
vpad = torch.randn(C, device=dev, requires_grad=True)
cot = torch.randn(B, H, W, C, device=dev).bfloat16()
for i in range(a.iters):
    out = ops.attention_core(qkv, inv, tab, sc, qpad, vpad, None, B, H, W, C, nH, ws, a.shift)
    out.backward(cot)
    qkv.grad = None
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
out = ops.attention_core(qkv, inv, tab, sc, qpad, vpad, None, B, H, W, C, nH, ws, a.shift)
e1.record()
out.backward(cot)
e2.record()
torch.cuda.synchronize()
items = B * ((H + ws - 1) // ws) ** 2 * nH
print(f"B={B} H={H} C={C} ws={ws} shift={a.shift} items={items}: fwd {e0.elapsed_time(e1):.3f} ms  bwd {e1.elapsed_time(e2):.3f} ms")
