"""A/B of the attention kernel families on one shape: outputs / gradients of impl X against impl Y + timings.
python tools/check_mma.py [--ws 12] [--B 48] [--H 30] [--C 512] [--shift 6] [--bwd 1]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=48); ap.add_argument("--H", type=int, default=30); ap.add_argument("--W", type=int, default=0)
ap.add_argument("--C", type=int, default=512); ap.add_argument("--ws", type=int, default=12); ap.add_argument("--shift", type=int, default=6)
ap.add_argument("--bwd", type=int, default=0); ap.add_argument("--a", type=int, default=4); ap.add_argument("--b", type=int, default=3)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
B, H, W, C, ws = a.B, a.H, a.W or a.H, a.C, a.ws
nH = C // 32
lib = L.load()
dev = "cuda"
torch.manual_seed(0)
T = B * H * W
q = torch.nn.functional.normalize(torch.randn(T, nH, 32, device=dev), dim=-1)
k = torch.nn.functional.normalize(torch.randn(T, nH, 32, device=dev), dim=-1)
v = torch.randn(T, nH, 32, device=dev)
qkv = torch.cat([q.reshape(T, C), k.reshape(T, C), v.reshape(T, C)], 1).bfloat16().contiguous()
tab = (16 * torch.sigmoid(torch.randn((2 * ws - 1) ** 2, nH, device=dev))).contiguous()
scale = (torch.rand(nH, device=dev) * 20 + 3).contiguous()
qpad = torch.nn.functional.normalize(torch.randn(nH, 32, device=dev), dim=-1).reshape(C).contiguous()
vpad = torch.randn(C, device=dev)
inv_norm = (torch.rand(T, 2, nH, device=dev) + 0.5).contiguous()
dout = torch.randn(T, C, device=dev).bfloat16()
Hp, Wp = (H + ws - 1) // ws * ws, (W + ws - 1) // ws * ws
nwin = B * (Hp // ws) * (Wp // ws)
st = torch.cuda.current_stream().cuda_stream

def run(impl, bwd):
    out = torch.zeros(T, C, device=dev, dtype=torch.bfloat16)
    out_lo = torch.zeros_like(out)
    lse = torch.zeros(nwin, nH, ws * ws, device=dev)
    L.check(lib.b200swin_attn_fwd(qkv.data_ptr(), out.data_ptr(), out_lo.data_ptr(), lse.data_ptr(), tab.data_ptr(), scale.data_ptr(),
                                  qpad.data_ptr(), vpad.data_ptr(), None, 0, B, H, W, C, nH, ws, a.shift, 1, impl, st), "fwd")
    res = [out.float(), out_lo.float(), lse]
    if bwd:
        wsb = lib.b200swin_attn_bwd_workspace_bytes(B, H, W, nH, ws, 1, impl)
        wsp = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
        dqkv = torch.zeros(T, 3 * C, device=dev, dtype=torch.bfloat16)
        acc = torch.zeros(tab.numel() + nH + C, device=dev)
        cs_ok = bool(lib.b200swin_attn_bwd_colsum_supported(ws, 1, impl))
        dcol = torch.zeros(3 * C, device=dev) if cs_ok else None
        L.check(lib.b200swin_attn_bwd(qkv.data_ptr(), out.data_ptr(), out_lo.data_ptr(), dout.data_ptr(), lse.data_ptr(), inv_norm.data_ptr(),
                                      tab.data_ptr(), scale.data_ptr(), qpad.data_ptr(), vpad.data_ptr(), None, 0, dqkv.data_ptr(),
                                      acc.data_ptr(), acc.data_ptr() + 4 * tab.numel(), acc.data_ptr() + 4 * (tab.numel() + nH),
                                      dcol.data_ptr() if cs_ok else None,
                                      B, H, W, C, nH, ws, a.shift, 1, impl, wsp.data_ptr(), wsb, st), "bwd")
        res += [dqkv[:, :C].float(), dqkv[:, C:2 * C].float(), dqkv[:, 2 * C:].float(), acc[:tab.numel()].clone(),
                acc[tab.numel():tab.numel() + nH].clone(), acc[tab.numel() + nH:].clone()]
        if cs_ok:
            ref_q, ref_v = dqkv[:, :C].float().sum(0), dqkv[:, 2 * C:].float().sum(0)
            eq = ((dcol[:C] - ref_q).norm() / ref_q.norm()).item()
            ev = ((dcol[2 * C:] - ref_v).norm() / ref_v.norm()).item()
            print(f"impl {impl}: column sums from the kernel vs sum over dqkv rows: dq {eq:.2e}  dv {ev:.2e}  (k part untouched: {float(dcol[C:2 * C].abs().max()):.1e})")
    return res

def timeit(impl, bwd):
    def f():
        run(impl, bwd)
    for _ in range(3): f()
    torch.cuda.synchronize()
    # time the C-ABI calls only: pre-allocate
    out = torch.zeros(T, C, device=dev, dtype=torch.bfloat16); out_lo = torch.zeros_like(out)
    lse = torch.zeros(nwin, nH, ws * ws, device=dev)
    wsb = lib.b200swin_attn_bwd_workspace_bytes(B, H, W, nH, ws, 1, impl)
    wsp = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)
    dqkv = torch.zeros(T, 3 * C, device=dev, dtype=torch.bfloat16); acc = torch.zeros(tab.numel() + nH + C, device=dev)
    e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
    tf = tb = 0.0
    for _ in range(a.iters):
        e0.record()
        lib.b200swin_attn_fwd(qkv.data_ptr(), out.data_ptr(), out_lo.data_ptr(), lse.data_ptr(), tab.data_ptr(), scale.data_ptr(),
                              qpad.data_ptr(), vpad.data_ptr(), None, 0, B, H, W, C, nH, ws, a.shift, 1, impl, st)
        e1.record()
        if bwd:
            lib.b200swin_attn_bwd(qkv.data_ptr(), out.data_ptr(), out_lo.data_ptr(), dout.data_ptr(), lse.data_ptr(), inv_norm.data_ptr(),
                                  tab.data_ptr(), scale.data_ptr(), qpad.data_ptr(), vpad.data_ptr(), None, 0, dqkv.data_ptr(),
                                  acc.data_ptr(), acc.data_ptr() + 4 * tab.numel(), acc.data_ptr() + 4 * (tab.numel() + nH), None,
                                  B, H, W, C, nH, ws, a.shift, 1, impl, wsp.data_ptr(), wsb, st)
        e2.record()
        torch.cuda.synchronize()
        tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
    return tf / a.iters * 1e3, tb / a.iters * 1e3

ra, rb = run(a.a, a.bwd), run(a.b, a.bwd)
names = ["out", "out_lo", "lse", "dq", "dk", "dv", "dtable", "dscale", "dvpad"]
for n, x, y in zip(names, ra, rb):
    den = y.double().norm().item()
    err = (x.double() - y.double()).norm().item()
    print(f"{n:8s} rel-L2 {err / den if den else err:.3e}  max|d| {(x - y).abs().max().item():.3e}  nan {int(torch.isnan(x).any())}")
items = nwin * nH
for impl in (a.a, a.b):
    tf, tb = timeit(impl, a.bwd)
    print(f"impl {impl}: fwd {tf:.1f} us ({tf * 1e-6 * 1.9e9 * 148 / items:.0f} cyc/item/SM)  bwd {tb:.1f} us ({tb * 1e-6 * 1.9e9 * 148 / items:.0f} cyc/item/SM)")
