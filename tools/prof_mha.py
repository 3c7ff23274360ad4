"""Timing of the global-attention core (b200swin_mha_fwd / _bwd) at the config-3 shape against torch's fused SDPA on the
same GPU.  python tools/prof_mha.py [B]"""
import sys

import torch

sys.path.insert(0, ".")
import b200swin  # noqa: E402,F401
from b200swin import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N, nH, hd = 1200, 8, 64
E = nH * hd
for dtype in (torch.bfloat16, torch.float32):
    q, k, v = [torch.randn(B, N, E, device="cuda", dtype=dtype, requires_grad=True) for _ in range(3)]
    cot = torch.randn(B, N, E, device="cuda", dtype=dtype)
    out, _ = ops.mha_core(q, k, v, nH)
    t_f = timeit(lambda: ops.mha_core(q, k, v, nH))
    t_b = timeit(lambda: torch.autograd.grad(out, [q, k, v], cot, retain_graph=True))
    qh, kh, vh = [t.detach().view(B, N, nH, hd).transpose(1, 2).requires_grad_(True) for t in (q, k, v)]
    o2 = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh)
    c2 = cot.view(B, N, nH, hd).transpose(1, 2)
    t_f2 = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
    t_b2 = timeit(lambda: torch.autograd.grad(o2, [qh, kh, vh], c2, retain_graph=True))
    fl = 4.0 * B * nH * N * N * hd
    print(f"{dtype}: B={B} N={N} {nH}x{hd}  fwd {t_f:.0f} us ({fl / t_f * 1e-6:.1f} TF/s)  bwd {t_b:.0f} us "
          f"({2.5 * fl / t_b * 1e-6:.1f} TF/s) | torch SDPA fwd {t_f2:.0f} us bwd {t_b2:.0f} us")
