"""Time the attention-core C-ABI entry points directly (no autograd, no per-call host work between the events):
`--iters` back-to-back launches between two CUDA events on the launching stream."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=48)
ap.add_argument("--H", type=int, default=30)
ap.add_argument("--C", type=int, default=512)
ap.add_argument("--ws", type=int, default=12)
ap.add_argument("--shift", type=int, default=0)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--bwd", type=int, default=1)
ap.add_argument("--impl", type=int, default=1, help="1 = tcgen05 (by window size), 2 = KV-blocked tcgen05, 0 = CUDA cores")
ap.add_argument("--W", type=int, default=0)
a = ap.parse_args()
lib = L.load()
B, H, W, C, ws = a.B, a.H, (a.W or a.H), a.C, a.ws
nH = C // 32
dev = "cuda"
torch.manual_seed(0)
qkv = torch.randn(B, H, W, 3 * C, device=dev)
q, k, v = qkv.split(C, -1)
nrm = lambda t: torch.nn.functional.normalize(t.reshape(B, H, W, nH, 32), dim=-1).reshape(B, H, W, C)
qkv = torch.cat([nrm(q), nrm(k), v], -1).bfloat16().contiguous()
Hp = (H + ws - 1) // ws * ws
Wp = (W + ws - 1) // ws * ws
nwin = B * (Hp // ws) * (Wp // ws)
inv = torch.ones(B * H * W, 2, nH, device=dev)
tab = (16 * torch.sigmoid(torch.randn((2 * ws - 1) ** 2, nH, device=dev))).contiguous()
sc = torch.full((nH,), 10.0, device=dev)
qpad = torch.nn.functional.normalize(torch.randn(nH, 32, device=dev), dim=-1).reshape(C).contiguous()
vpad = torch.randn(C, device=dev)
out = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
out_lo = torch.empty_like(out)
lse = torch.empty(nwin, nH, ws * ws, device=dev)
dout = torch.randn(B, H, W, C, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
acc = torch.zeros(tab.numel() + nH + C, device=dev)
st = torch.cuda.current_stream().cuda_stream
wsb = lib.b200swin_attn_bwd_workspace_bytes(B, H, W, nH, ws, 1, a.impl)
wsp = torch.empty(max(wsb, 16), dtype=torch.uint8, device=dev)


def fwd():
    L.check(lib.b200swin_attn_fwd(qkv.data_ptr(), out.data_ptr(), out_lo.data_ptr(), lse.data_ptr(), tab.data_ptr(), sc.data_ptr(),
                                  qpad.data_ptr(), vpad.data_ptr(), None, 0, B, H, W, C, nH, ws, a.shift, 1, a.impl, st), "fwd")


def bwd():
    L.check(lib.b200swin_attn_bwd(qkv.data_ptr(), out.data_ptr(), out_lo.data_ptr(), dout.data_ptr(), lse.data_ptr(), inv.data_ptr(),
                                  tab.data_ptr(), sc.data_ptr(), qpad.data_ptr(), vpad.data_ptr(), None, 0,
                                  dqkv.data_ptr(), acc.data_ptr(), acc.data_ptr() + 4 * tab.numel(),
                                  acc.data_ptr() + 4 * (tab.numel() + nH), None, B, H, W, C, nH, ws, a.shift, 1, a.impl,
                                  wsp.data_ptr(), wsb, st), "bwd")


def timeit(fn):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


items = nwin * nH
N = ws * ws
tf = fwd and timeit(fwd)
gb_f = (B * H * W * C * 2 * 4 + items * N * 4) / 1e9          # qkv in, o out (bf16) + lse
flops_f = 4.0 * items * N * N * 32
msg = f"impl={a.impl} B={B} H={H}x{W} C={C} ws={ws} shift={a.shift} items={items}: fwd {tf*1e3:.1f} us ({gb_f/tf*1e3:.0f} GB/s, {flops_f/tf/1e9:.1f} TF/s, {tf*1e-3*1.9e9*148/items:.0f} cyc/item/SM)"
if a.bwd:
    tb = timeit(bwd)
    msg += f"  bwd {tb*1e3:.1f} us ({tb*1e-3*1.9e9*148/items:.0f} cyc/item/SM)"
print(msg)
