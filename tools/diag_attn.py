import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops
B, H, C, ws, shift = [int(a) for a in sys.argv[1:6]]
dev = "cuda"; W, nH = H, C // 32
gen = torch.Generator(device=dev).manual_seed(H * 7 + ws)
T = B * H * W
nrm = torch.nn.functional.normalize
q = nrm(torch.randn(T, nH, 32, device=dev, generator=gen), dim=-1).reshape(T, C)
k = nrm(torch.randn(T, nH, 32, device=dev, generator=gen), dim=-1).reshape(T, C)
v = torch.randn(T, C, device=dev, generator=gen)
inv = torch.rand(T, 2, nH, device=dev, generator=gen) + 0.5
tab = 16 * torch.sigmoid(torch.randn((2 * ws - 1) ** 2, nH, device=dev, generator=gen))
sc = torch.rand(nH, device=dev, generator=gen) * 20 + 1
qpad = nrm(torch.randn(nH, 32, device=dev, generator=gen), dim=-1).reshape(C)
vpad = torch.randn(C, device=dev, generator=gen)
cot = torch.randn(B, H, W, C, device=dev, generator=gen).bfloat16()
res = {}
for impl in ("tc", "simt"):
    ops.ATTN_IMPL["mode"] = ops.ATTN_IMPL["bwd_mode"] = impl
    leaf = torch.cat([q, k, v], 1).bfloat16().view(B, H, W, 3 * C).requires_grad_(True)
    tl, sl, vl = (t.clone().requires_grad_(True) for t in (tab, sc, vpad))
    o = ops.attention_core(leaf, inv, tl, sl, qpad, vl, None, B, H, W, C, nH, ws, shift)
    o.backward(cot)
    res[impl] = [o.detach().float(), leaf.grad.float(), tl.grad, sl.grad, vl.grad]
if os.environ.get("DIAG_REPEAT"):
    # run-to-run determinism of the tensor-core backward
    ops.ATTN_IMPL["mode"] = ops.ATTN_IMPL["bwd_mode"] = "tc"
    grads = []
    for _ in range(3):
        leaf = torch.cat([q, k, v], 1).bfloat16().view(B, H, W, 3 * C).requires_grad_(True)
        o = ops.attention_core(leaf, inv, tab, sc, qpad, vpad, None, B, H, W, C, nH, ws, shift)
        o.backward(cot)
        grads.append(leaf.grad.float().clone())
    print("tc run-to-run max abs diff:", (grads[0] - grads[1]).abs().max().item(), (grads[1] - grads[2]).abs().max().item())
def rel(a, b): return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
a, b = res["tc"][1].view(B, H, W, 3, nH, 32), res["simt"][1].view(B, H, W, 3, nH, 32)
print("out", rel(res["tc"][0], res["simt"][0]))
oa, ob = res["tc"][0].view(B, H, W, nH, 32), res["simt"][0].view(B, H, W, nH, 32)
eo = ((oa - ob).pow(2).sum(-1) / ob.pow(2).sum(-1).clamp_min(1e-12)).sqrt()      # [B,H,W,nH]
print("out err per head", [round(rel(oa[:, :, :, h], ob[:, :, :, h]), 4) for h in range(min(nH, 16))])
print("out err per batch", [round(rel(oa[i], ob[i]), 4) for i in range(min(B, 12))])
badt = (eo > 0.1).nonzero()
print("bad out (token,head)", badt.shape[0], badt[:12].tolist())
rowerr = eo.mean((0, 2, 3)).cpu().numpy().round(3)
print("out err by image row", rowerr.tolist())
for i, nm in enumerate("qkv"):
    print("d" + nm, rel(a[:, :, :, i], b[:, :, :, i]), "per head:", [round(rel(a[:, :, :, i, h], b[:, :, :, i, h]), 3) for h in range(min(nH, 16))])
# where are the bad tokens? per (batch) and per row
d = (a - b).pow(2).sum((3, 4, 5)).sqrt() / b.pow(2).sum((3, 4, 5)).sqrt().clamp_min(1e-9)
print("per-batch max token err", [round(d[i].max().item(), 3) for i in range(min(B, 12))])
bad = (d > 0.1).nonzero()
print("bad tokens", bad.shape[0], "of", T, bad[:10].tolist())
print("dtable", rel(res["tc"][2], res["simt"][2]), "dscale", rel(res["tc"][3], res["simt"][3]), "dvpad", rel(res["tc"][4], res["simt"][4]))
# error by in-window key position and by head for dk
Hp = (H + ws - 1) // ws * ws
dkerr = (a[:, :, :, 1] - b[:, :, :, 1]).pow(2).sum(-1)        # [B,H,W,nH]
dkref = b[:, :, :, 1].pow(2).sum(-1)
pos = torch.zeros(ws, ws); cnt = torch.zeros(ws, ws)
e = (dkerr.sum(-1) / dkref.sum(-1).clamp_min(1e-12)).sqrt().cpu()   # [B,H,W]
for i in range(H):
    for j in range(W):
        pos[i % ws, j % ws] += e[:, i, j].mean(); cnt[i % ws, j % ws] += 1
print("dk err by in-window position (unshifted windows):")
print((pos / cnt).numpy().round(2))
print("dk err by batch index:", [round(e[i].mean().item(), 3) for i in range(B)][:16])
