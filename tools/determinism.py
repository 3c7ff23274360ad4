"""Run the tensor-core attention forward / backward repeatedly on the same inputs: any run-to-run difference in the
outputs that are not accumulated with atomics (out, lse, dqkv) is a race."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops
dev = "cuda"
def run(B, H, C, ws, shift, reps=6):
    W, nH = H, C // 32
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
    ops.ATTN_IMPL["mode"] = ops.ATTN_IMPL["bwd_mode"] = "tc"
    outs, grads = [], []
    for _ in range(reps):
        leaf = torch.cat([q, k, v], 1).bfloat16().view(B, H, W, 3 * C).requires_grad_(True)
        o = ops.attention_core(leaf, inv, tab, sc, qpad, vpad, None, B, H, W, C, nH, ws, shift)
        o.backward(cot)
        outs.append(o.detach().clone()); grads.append(leaf.grad.clone())
    do = max((outs[0].float() - x.float()).abs().max().item() for x in outs[1:])
    dg = max((grads[0].float() - x.float()).abs().max().item() for x in grads[1:])
    nbo = max(int((outs[0] != x).sum().item()) for x in outs[1:])
    nbg = max(int((grads[0] != x).sum().item()) for x in grads[1:])
    if nbg:
        # where: part (q / k / v), in-window row of the token, window position
        from collections import Counter
        cnt = Counter()
        for x in grads[1:]:
            idx = (grads[0] != x).nonzero()
            if idx.numel() == 0:
                continue
            Hp = (H + ws - 1) // ws * ws
            b, i, j, col = idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]
            si, sj = (i - shift) % Hp, (j - shift) % Hp          # coordinates on the shifted padded grid
            r = (si % ws) * ws + (sj % ws)
            part = col // C
            for pp, rr, wh, ww in zip(part.tolist(), r.tolist(), (si // ws).tolist(), (sj // ws).tolist()):
                cnt[(pp, rr // 16 * 16, wh, ww)] += 1
            break
        print("   mismatches by (part q0/k1/v2, in-window row bucket of 16, window row, window col):",
              sorted(cnt.items(), key=lambda kv: -kv[1])[:12])
    print(f"B={B} H={H} C={C} ws={ws} shift={shift}: fwd max|diff| {do:.3g} ({nbo} elems)   bwd dqkv max|diff| {dg:.3g} ({nbg} elems)")
for cfg in [(8, 120, 128, 12, 6), (8, 120, 128, 12, 6), (8, 120, 128, 12, 0), (48, 30, 512, 12, 0), (48, 15, 1024, 6, 0), (48, 12, 1024, 6, 0), (16, 64, 128, 8, 4), (8, 60, 256, 12, 6)]:
    run(*cfg)
