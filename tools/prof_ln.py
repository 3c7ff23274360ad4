"""LayerNorm+residual fwd/bwd and colsum on the config-2 shapes: ms and achieved GB/s (algorithmic bytes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=5):
    fn(); fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for T, C, nblk in [(48 * 14400, 128, 4), (48 * 3600, 256, 4), (48 * 900, 512, 36), (48 * 225, 1024, 4)]:
    x = torch.randn(T, C, device=dev).bfloat16().requires_grad_(True)
    r = torch.randn(T, C, device=dev).bfloat16()
    g = torch.ones(C, device=dev, requires_grad=True); b = torch.zeros(C, device=dev, requires_grad=True)
    dy = torch.randn(T, C, device=dev).bfloat16()
    y = ops.layer_norm_residual(x, g, b, 1e-6, residual=r)
    tf = timeit(lambda: ops.layer_norm_residual(x, g, b, 1e-6, residual=r))
    def bw():
        x.grad = None
        y.backward(dy, retain_graph=True)
    tb = timeit(bw)
    tc = timeit(lambda: ops.colsum(dy))
    by = T * C * 2
    print(f"T={T} C={C}: ln fwd {tf*1e3:7.1f} us ({3*by/tf/1e6:6.0f} GB/s)  ln bwd {tb*1e3:7.1f} us ({3*by/tb/1e6:6.0f} GB/s)  "
          f"colsum {tc*1e3:7.1f} us ({by/tc/1e6:6.0f} GB/s)   x{nblk} -> {nblk*(tf+tb)+nblk*tc:.2f} ms/step")
