import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops, SiLogLoss
dev = "cuda"
torch.manual_seed(0)
def attempt(name, fn):
    ops._weight_cache.clear()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    ops._weight_cache.clear()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "OK")
    except Exception as e:
        print(name, "FAILED:", str(e).splitlines()[0])
which = sys.argv[1]
W = torch.randn(1024, 1024, device=dev, requires_grad=True); b = torch.zeros(1024, device=dev, requires_grad=True)
feat = torch.randn(8, 1024, 15, 15, device=dev, requires_grad=True)
tgt = torch.rand(8, 480, 480, device=dev) * 9 + 0.5
crit = SiLogLoss()
def head(upto):
    with torch.autocast("cuda", torch.bfloat16):
        B, C, h, w = feat.shape
        tok = feat.permute(0, 2, 3, 1).reshape(B, h * w, C)
        if upto == "tok": return tok.sum()
        d = ops.linear(tok, W, b)
        if upto == "lin": return d.float().sum()
        d = d.view(B, h, w, 32, 32).permute(0, 1, 3, 2, 4).reshape(B, h * 32, w * 32)
        if upto == "shuffle": return d.float().sum()
        d = torch.sigmoid(d.float()) * 10.0
        if upto == "sigmoid": return d.sum()
        p1, p2 = d.chunk(2, dim=0)
        if upto == "chunk": return p1.sum() + p2.sum()
        return (crit(p1, tgt[:4]) + crit(p2, tgt[4:])) / 2
mode = sys.argv[2] if len(sys.argv) > 2 else ""
if "eager" in mode:
    for _ in range(2):
        head(which).backward()
    torch.cuda.synchronize()
if "opt" in mode:
    opt = torch.optim.AdamW([W, b], lr=1e-3, fused=True, capturable=True)
    head(which).backward(); opt.step(); torch.cuda.synchronize()
if "zero" in mode:
    def f():
        W.grad = None; b.grad = None; feat.grad = None
        head(which).backward()
    attempt(which + " " + mode, f)
else:
    attempt(which + " " + mode, lambda: head(which).backward())
