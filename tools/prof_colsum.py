"""Column-sum kernel on the bias-gradient shapes of config 2 (one call per event pair; L2 flushed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200swin import ops
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for M, N, c0, nc in [(43200, 2048, 0, 2048), (172800, 1024, 0, 1024), (691200, 512, 0, 512), (43200, 1536, 0, 512), (43200, 1536, 1024, 512)]:
    x = torch.randn(M, N, device="cuda").bfloat16()
    fn = lambda: ops.colsum(x, c0, nc)
    fn(); fn()
    ts = []
    for _ in range(5):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]
    print(f"M={M} ld={N} cols [{c0},{c0+nc}): {t*1e3:.1f} us  {M*nc*2/t/1e6:.0f} GB/s")
