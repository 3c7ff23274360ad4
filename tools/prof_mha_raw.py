"""Launch the global-attention entry points (b200swin_mha_fwd / _bwd) directly at the config-3 shape -- the program the ncu
captures of profiles/r02_ncu_gattn_*.json profile.  python tools/prof_mha_raw.py [--B 16] [--iters 3]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=16)
ap.add_argument("--N", type=int, default=1200)
ap.add_argument("--nH", type=int, default=8)
ap.add_argument("--hd", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
lib = L.load()
B, N, nH, hd = a.B, a.N, a.nH, a.hd
E = nH * hd
torch.manual_seed(0)
q, k, v, dout = [torch.randn(B, N, E, device="cuda").bfloat16() for _ in range(4)]
out = torch.empty_like(q)
dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
lse = torch.empty(B, nH, N, device="cuda")
wsb = lib.b200swin_mha_bwd_workspace_bytes(B, N, nH)
wsp = torch.empty(wsb, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
sc = hd ** -0.5


def fwd():
    L.check(lib.b200swin_mha_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), E, E, E, out.data_ptr(), E, lse.data_ptr(), B, N, N,
                                 nH, hd, sc, L.BF16, st), "mha_fwd")


def bwd():
    L.check(lib.b200swin_mha_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), E, E, E, out.data_ptr(), E, dout.data_ptr(), E,
                                 lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), E, E, E, B, N, N, nH, hd, sc,
                                 L.BF16, wsp.data_ptr(), wsb, st), "mha_bwd")


for fn, name, fl in ((fwd, "fwd", 4.0), (bwd, "bwd (prep + dK/dV pass + dQ pass)", 10.0)):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / a.iters * 1e3
    print(f"{name}: {us:.1f} us  {fl * B * nH * N * N * hd / us * 1e-6:.1f} TF/s (algorithmic)")
