"""torch.profiler view of one bench step: which aten ops (with shapes) own the non-b200swin GPU time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = bench.DepthModel().to(dev).train()
from b200swin import SiLogLoss
crit = SiLogLoss()
opt = torch.optim.AdamW(model.parameters(), lr=5e-4, weight_decay=0.05, fused=True)
batch = [t.to(dev) for t in bench.make_batch(24, 1234)]

def step():
    img1, img2, d1, d2 = batch
    with torch.autocast("cuda", torch.bfloat16):
        p1, p2 = model(img1, img2)
    loss = (crit(p1, d1) + crit(p2, d2)) / 2
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = [(e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:150]) for e in ka if e.key.startswith("aten::") and e.self_device_time_total > 0]
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"aten ops with GPU time: {tot/1e3:.2f} ms total")
for t, n, k, sh in rows[:40]:
    print(f"{t/1e3:8.3f} ms  x{n:4d}  {k:32s} {sh}")
