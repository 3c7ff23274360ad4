"""N > 1 host logic on CPU (gloo, world_size 2): the drop-in modules' parameters under DistributedDataParallel.

The CUDA kernels cannot run here, so the forward is executed by the CPU oracle over the DROP-IN module's own
parameters (same names, same tensors); what is under test is the multi-process plumbing the GPU path uses
unchanged: per-rank data sharding, gradient all-reduce as the only collective, and that the averaged gradient
equals the single-process gradient of the mean of the per-rank losses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

CFG = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=[4, 4], pretrain_window_size=[4, 4],
           use_shift=[True, False], drop_path_rate=0.0, out_indices=(1,))


class OracleRunner(nn.Module):
    """Runs oracle.swin_ref over the parameters of a b200swin SwinTransformerV2 (CPU only, tests only)."""

    def __init__(self):
        super().__init__()
        from b200swin.swin_transformer_v2 import SwinTransformerV2
        torch.manual_seed(0)
        self.net = SwinTransformerV2(**CFG)
        self.net.init_weights(None)
        with torch.no_grad():
            for n, p in self.net.named_parameters():
                if "norm" in n and n.endswith("weight"):
                    p.fill_(1.0)

    def forward(self, img, target):
        from oracle import silog_ref, swin_ref
        sd = dict(self.net.named_parameters())
        sd.update(dict(self.net.named_buffers()))
        feat = swin_ref.swin_v2(img, sd, CFG["embed_dim"], CFG["depths"], CFG["num_heads"], CFG["window_size"],
                                CFG["use_shift"], CFG["out_indices"])[0]
        pred = torch.sigmoid(feat.mean(1)) * 10.0 + 0.1
        return silog_ref.silog_torch(pred, target)


def _data(rank):
    g = torch.Generator().manual_seed(1234 + rank)            # bench.py's per-rank data seeding
    img = torch.rand(2, 3, 32, 40, generator=g)
    tgt = torch.rand(2, 4, 5, generator=g) * 9 + 0.5
    tgt[tgt < 1.5] = 0
    return img, tgt


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    model = OracleRunner()
    ddp = nn.parallel.DistributedDataParallel(model)
    img, tgt = _data(rank)
    loss = ddp(img, tgt)
    loss.backward()
    grads = {n: p.grad.clone() for n, p in model.net.named_parameters() if p.grad is not None}
    # timing contract of bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret["grads"] = grads
        ret["loss"] = loss.item()
        ret["tmax"] = t.item()
    dist.destroy_process_group()


def test_ddp_world2_matches_mean_of_rank_losses():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["tmax"] == 2.0
    # single process: mean of the two per-rank losses
    model = OracleRunner()
    total = 0
    for r in range(2):
        total = total + model(*_data(r)) / 2
    total.backward()
    checked = 0
    for n, p in model.net.named_parameters():
        if p.grad is None:
            continue
        g = ret["grads"][n]
        torch.testing.assert_close(g, p.grad, rtol=2e-4, atol=1e-6 + 2e-5 * p.grad.abs().max().item())
        checked += 1
    assert checked > 40


def test_rank_shards_are_disjoint():
    a, b = _data(0)[0], _data(1)[0]
    assert not torch.equal(a, b)
