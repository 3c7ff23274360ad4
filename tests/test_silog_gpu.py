"""SiLog kernel parity (GPU): golden vectors from the reference, the numpy oracle on seeded inputs
at the BASELINE sizes, bf16 predictions, and size-independent properties.
Tolerances: fp32 loss/grad 1e-4 relative (north_star); bf16 2e-2."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import silog_ref

pytestmark = pytest.mark.gpu


def _run(pred, tgt, gout=1.0, lambd=0.5):
    import b200swin
    p = pred.clone().cuda().requires_grad_(True)
    loss = b200swin.SiLogLoss(lambd)(p, tgt.cuda())
    (loss * gout).backward()
    return loss.item(), p.grad.float().cpu().numpy()


@pytest.mark.parametrize("nm", ["kat", "nyu", "kitti", "void", "allvalid"])
def test_golden(nm):
    g = load_golden("silog")
    loss, grad = _run(torch.from_numpy(g[f"{nm}.pred"]), torch.from_numpy(g[f"{nm}.target"]), gout=1.7)
    ref = float(g[f"{nm}.loss"][0])
    assert abs(loss - ref) <= 1e-4 * abs(ref)
    gr = g[f"{nm}.grad_x1p7"]
    np.testing.assert_allclose(grad, gr, rtol=1e-4, atol=1e-4 * np.abs(gr).max())


def test_all_invalid_is_nan_like_reference():
    loss, _ = _run(torch.ones(64), torch.zeros(64))
    assert np.isnan(loss)


@pytest.mark.parametrize("shape,maxd,inval", [((24, 480, 480), 10.0, 0.05), ((8, 352, 1216), 80.0, 0.9),
                                              ((16, 480, 640), 10.0, 0.3), ((1, 7, 13), 10.0, 0.2)])
def test_full_size_against_oracle(shape, maxd, inval):
    gen = torch.Generator().manual_seed(7)
    tgt = 0.5 + (maxd - 0.5) * torch.rand(shape, generator=gen)
    tgt = torch.where(torch.rand(shape, generator=gen) < inval, torch.zeros(()), tgt)
    pred = 0.2 + maxd * torch.rand(shape, generator=gen)
    loss, grad = _run(pred, tgt)
    ref = silog_ref.silog_np(pred.numpy(), tgt.numpy())
    assert abs(loss - ref) <= 1e-4 * abs(ref)
    gr = silog_ref.silog_grad_np(pred.numpy(), tgt.numpy())
    np.testing.assert_allclose(grad, gr, rtol=2e-4, atol=1e-4 * np.abs(gr).max())
    assert (grad[tgt.numpy() <= 0] == 0).all()


def test_bf16_pred():
    gen = torch.Generator().manual_seed(9)
    tgt = 0.5 + 9.5 * torch.rand(4, 120, 160, generator=gen)
    pred = (0.2 + 10 * torch.rand(4, 120, 160, generator=gen)).bfloat16()
    loss, grad = _run(pred, tgt)
    ref = silog_ref.silog_np(pred.float().numpy(), tgt.numpy())
    assert abs(loss - ref) <= 1e-4 * abs(ref)          # same bf16-rounded inputs -> fp32 math
    gr = silog_ref.silog_grad_np(pred.float().numpy(), tgt.numpy())
    np.testing.assert_allclose(grad, gr, rtol=2e-2, atol=2e-2 * np.abs(gr).max())


def test_properties_scale_and_determinism():
    gen = torch.Generator().manual_seed(11)
    tgt = 0.5 + 9.5 * torch.rand(3, 97, 131, generator=gen)      # odd sizes: scalar tail path
    tgt[tgt < 2.0] = 0
    pred = 0.2 + 10 * torch.rand(3, 97, 131, generator=gen)
    l1, g1 = _run(pred, tgt)
    l2, g2 = _run(pred, tgt)
    assert l1 == l2 and np.array_equal(g1, g2)                   # fixed-order reduction -> bitwise repeatable
    # lambd = 1: loss is the std of d, invariant to a global scale of pred
    la, _ = _run(pred, tgt, lambd=1.0)
    lb, _ = _run(pred * 3.7, tgt, lambd=1.0)
    assert abs(la - lb) <= 2e-4 * la
    # gradient is linear in the upstream gradient
    _, g3 = _run(pred, tgt, gout=-2.5)
    np.testing.assert_allclose(g3, -2.5 * g1, rtol=1e-5, atol=1e-9)
    # non-contiguous inputs and a squeezed channel dim (train.py:215 passes pred.squeeze(1))
    import b200swin
    p4 = pred.unsqueeze(1).cuda()
    l4 = b200swin.SiLogLoss()(p4.squeeze(1), tgt.cuda()).item()
    assert l4 == l1
