"""The product's data-parallel wrapper (b200swin.data_parallel.DataParallel: flat gradient buffer, bucketed all-reduce,
optional overlap through post-accumulate hooks) with world_size 2 over gloo on CPU.  What is under test is exactly the
code the GPU path runs (NCCL there): broadcast of the parameters, per-rank data, averaged gradients equal to the
single-process gradient of the mean of the per-rank losses, every bucket covered, gradients left as flat views."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _net(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(6, 2000), torch.nn.GELU(), torch.nn.Linear(2000, 9), torch.nn.LayerNorm(9),
                               torch.nn.Linear(9, 1))


def _data(rank):
    g = torch.Generator().manual_seed(1234 + rank)
    return torch.randn(5, 6, generator=g), torch.randn(5, 1, generator=g)


def _worker(rank, world, port, overlap, buckets, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from b200swin.data_parallel import DataParallel
    net = _net(rank)                                   # different initial weights per rank: the wrapper must broadcast
    dp = DataParallel(net, buckets=buckets, overlap=overlap, bf16_copies=False)
    x, y = _data(rank)
    loss = torch.nn.functional.mse_loss(dp(x), y)
    loss.backward()
    dp.reduce_gradients()
    ok = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(dp.flat.params, dp.flat.grad_views))
    covered = sorted((lo, hi) for lo, hi, _, _ in dp._bounds)
    ret[rank] = dict(grads=[p.grad.clone() for p in net.parameters()], weights=[p.detach().clone() for p in net.parameters()],
                     flat_views=ok, covered=covered, total=dp.flat.total, nb=len(dp._bounds))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap,buckets", [(False, 1), (False, 3), (True, 2)])
def test_flat_buffer_data_parallel_world2(overlap, buckets):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, overlap, buckets, ret), nprocs=2, join=True)
    r0, r1 = ret[0], ret[1]
    # replicas: rank 0's weights everywhere
    ref = _net(0)
    for a, b, c in zip(r0["weights"], r1["weights"], ref.parameters()):
        assert torch.equal(a, b) and torch.equal(a, c.detach())
    # averaged gradient == gradient of the mean of the two per-rank losses in one process
    total = 0
    for rank in range(2):
        x, y = _data(rank)
        total = total + torch.nn.functional.mse_loss(ref(x), y)
    (total / 2).backward()
    for a, b, p in zip(r0["grads"], r1["grads"], ref.parameters()):
        assert torch.equal(a, b)
        torch.testing.assert_close(a, p.grad, rtol=1e-5, atol=1e-7)
    assert r0["flat_views"] and r1["flat_views"]
    # the buckets tile the flat buffer exactly
    lo = 0
    for a, b in r0["covered"]:
        assert a == lo
        lo = b
    assert lo == r0["total"] and r0["nb"] <= max(1, buckets)
