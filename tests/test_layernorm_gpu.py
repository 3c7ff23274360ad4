"""LayerNorm(+DropPath scale +residual) kernel parity against the oracle restatement of
LayerNormFP32 and the post-norm residual adds (swin_transformer_v2.py:41-47, 472-474)."""
import numpy as np
import pytest
import torch

from oracle import swin_ref

pytestmark = pytest.mark.gpu


def _ref(x, res, g, b, rs, rps, eps, cot):
    x = x.double().requires_grad_(True)
    g = g.double().requires_grad_(True)
    b = b.double().requires_grad_(True)
    y = swin_ref.layer_norm_fp32(x, g, b, eps)
    if rs is not None:
        y = y * rs.double().repeat_interleave(rps).view(-1, 1)
    if res is not None:
        res = res.double().requires_grad_(True)
        y = y + res
    gr = torch.autograd.grad((y * cot.double()).sum(), [x, g, b] + ([res] if res is not None else []))
    return y.detach(), gr


@pytest.mark.parametrize("rows,C,dtype,with_res,with_scale", [
    (14400, 128, torch.float32, True, False), (900 * 2, 512, torch.float32, True, True),
    (225 * 3, 1024, torch.bfloat16, True, True), (37, 96, torch.float32, False, False),
    (64, 1536, torch.float32, True, False), (3600, 256, torch.bfloat16, False, False),
    (50, 352, torch.float32, True, True)])
def test_ln_residual(rows, C, dtype, with_res, with_scale):
    from b200swin import ops
    gen = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=gen) * 2 + 0.5).to(dtype)
    res = torch.randn(rows, C, generator=gen).to(dtype) if with_res else None
    g = 1 + 0.3 * torch.randn(C, generator=gen)
    b = 0.2 * torch.randn(C, generator=gen)
    rps = rows // 3 if rows % 3 == 0 else rows
    rs = torch.tensor([0.0, 1.25, 1.25])[: rows // rps].contiguous() if with_scale else None
    cot = torch.randn(rows, C, generator=gen).to(dtype)
    eps = 1e-6
    yr, gr = _ref(x.float(), None if res is None else res.float(), g, b, rs, rps, eps, cot.float())

    xg = x.cuda().requires_grad_(True)
    gg = g.cuda().requires_grad_(True)
    bg = b.cuda().requires_grad_(True)
    rg = res.cuda().requires_grad_(True) if with_res else None
    y = ops.layer_norm_residual(xg, gg, bg, eps, residual=rg, row_scale=None if rs is None else rs.cuda(),
                                rows_per_scale=rps)
    assert y.dtype == dtype
    (y * cot.cuda()).sum().backward()
    tol = 1e-4 if dtype == torch.float32 else 2e-2

    def close(a, ref, name):
        a = a.detach().double().cpu()
        err = (a - ref).norm() / ref.norm().clamp_min(1e-30)
        assert err <= tol, f"{name}: rel-L2 {err:.3e}"
        if dtype == torch.float32:
            np.testing.assert_allclose(a.numpy(), ref.numpy(), rtol=1e-3, atol=1e-4 * ref.abs().max().item(),
                                       err_msg=name)

    close(y, yr, "y")
    close(xg.grad, gr[0], "dx")
    close(gg.grad, gr[1], "dgamma")
    close(bg.grad, gr[2], "dbeta")
    if with_res:
        close(rg.grad, gr[3], "dres")
