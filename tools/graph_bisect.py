"""Find which part of the training step cannot be captured in a CUDA graph."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from b200swin import SiLogLoss, ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = bench.DepthModel().to(dev).train()
crit = SiLogLoss()
params = list(model.parameters())
opt = torch.optim.AdamW(params, lr=5e-4, weight_decay=0.05, fused=True, capturable=True)
batch = [t.to(dev) for t in bench.make_batch(4, 1234)]

def fwd():
    with torch.autocast("cuda", torch.bfloat16):
        p1, p2 = model(batch[0], batch[1])
    return (crit(p1, batch[2]) + crit(p2, batch[3])) / 2

for _ in range(3):
    opt.zero_grad(set_to_none=True); l = fwd(); l.backward(); opt.step()
torch.cuda.synchronize()

def attempt(name, fn):
    ops._weight_cache.clear()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
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
        torch.cuda.synchronize()

stage = sys.argv[1]
if stage == "enc_fwd":
    def f():
        with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
            model.encoder(torch.cat([batch[0], batch[1]]))
    attempt("encoder forward (no grad)", f)
elif stage == "fwd":
    attempt("forward + loss", lambda: fwd())
elif stage == "fwd_bwd":
    def f():
        opt.zero_grad(set_to_none=True); l = fwd(); l.backward()
    attempt("forward + backward", f)
elif stage == "opt":
    attempt("optimizer", lambda: opt.step())
elif stage == "silog":
    p = torch.rand(2, 64, 64, device=dev, requires_grad=True)
    def f():
        l = crit(p * 5 + 0.1, batch[2][:2, :64, :64]); l.backward()
    attempt("silog fwd+bwd", f)
if stage in ("model_sum", "model_silog1", "readout"):
    if stage == "model_sum":
        def f():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", torch.bfloat16):
                p1, p2 = model(batch[0], batch[1])
            (p1.sum() + p2.sum()).backward()
        attempt("model fwd+bwd (sum loss)", f)
    elif stage == "model_silog1":
        def f():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", torch.bfloat16):
                p1, p2 = model(batch[0], batch[1])
            crit(p1, batch[2]).backward()
        attempt("model fwd+bwd (one silog)", f)
    else:
        feat = torch.randn(8, 1024, 15, 15, device=dev, requires_grad=True)
        def f():
            with torch.autocast("cuda", torch.bfloat16):
                B, C, h, w = feat.shape
                tok = feat.permute(0, 2, 3, 1).reshape(B, h * w, C)
                d = ops.linear(tok, model.readout.weight, model.readout.bias)
                d = d.view(B, h, w, 32, 32).permute(0, 1, 3, 2, 4).reshape(B, h * 32, w * 32)
                d = torch.sigmoid(d.float()) * 10.0
                p1, p2 = d.chunk(2, dim=0)
            ((crit(p1, batch[2]) + crit(p2, batch[3])) / 2).backward()
        attempt("readout + silog x2", f)
