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


@pytest.mark.parametrize("rows,C,with_res", [(1000, 128, True), (777, 512, True), (64, 1024, False), (4096, 256, True)])
def test_fp32_residual_stream_variant(rows, C, with_res):
    """bf16 activations with an fp32 residual stream (torch.autocast semantics): y32 = residual32 + scale * LN(x) in fp32,
    y = bf16(y32); gradients as in the plain bf16 call."""
    from b200swin import ops
    g = torch.Generator().manual_seed(rows + C)
    x = torch.randn(rows, C, generator=g).bfloat16().cuda().requires_grad_(True)
    res32 = (torch.randn(rows, C, generator=g) * 3).cuda() if with_res else None
    res16 = res32.bfloat16().requires_grad_(True) if with_res else None
    gamma = (torch.rand(C, generator=g) * 1e-3).cuda().requires_grad_(True)       # tiny gamma: far below a bf16 ulp of res
    beta = (torch.randn(C, generator=g) * 1e-3).cuda().requires_grad_(True)
    y, y32 = ops.layer_norm_residual(x, gamma, beta, 1e-6, residual=res16, residual32=res32, stream32=True)
    ref = torch.nn.functional.layer_norm(x.detach().double(), (C,), gamma.detach().double(), beta.detach().double(), 1e-6)
    if with_res:
        ref = ref + res32.double()
    assert y32.dtype == torch.float32 and not y32.requires_grad
    assert (y32.double() - ref).abs().max().item() < 1e-5 * max(1.0, ref.abs().max().item())
    assert torch.equal(y, y32.bfloat16())
    cot = torch.randn(rows, C, generator=g).cuda()
    (y.float() * cot).sum().backward()
    y_plain = ops.layer_norm_residual(x.detach().requires_grad_(True), gamma, beta, 1e-6, residual=res16)
    assert x.grad is not None and (not with_res or torch.equal(res16.grad.float(), cot.bfloat16().float()))


def test_tiny_block_norm_gammas_survive_bf16_autocast():
    """The reference initialises every block norm to gamma = 1e-5 (res-post-norm).  With a bf16 residual stream each
    block's contribution (~1e-5 relative) is rounded away; with the fp32 stream the bf16-autocast encoder follows the
    fp32 oracle an order of magnitude inside the bf16 bar."""
    from b200swin.swin_transformer_v2 import SwinTransformerV2
    from oracle import swin_ref
    cfg = dict(embed_dim=64, depths=[2, 2], num_heads=[2, 4], window_size=[4, 4], pretrain_window_size=[4, 4],
               use_shift=[True, False], drop_path_rate=0.0, out_indices=(0, 1))
    torch.manual_seed(3)
    enc = SwinTransformerV2(**cfg)
    enc.init_weights(None)                                  # block norms at 1e-5
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if ".blocks." in n and "norm" in n and n.endswith("weight"):
                p.fill_(3e-3)                               # contribution ~3e-3 of the stream: one bf16 ulp is 4e-3
    img = torch.rand(2, 3, 64, 64)
    sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    ref = swin_ref.swin_v2(img, sd, cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"],
                           cfg["use_shift"], cfg["out_indices"])
    # what the blocks ADD to the stream (the part a bf16 stream loses): compare the block contribution, not the stream
    sd0 = {k: (torch.zeros_like(v) if (".blocks." in k and "norm" in k) else v) for k, v in sd.items()}
    base = swin_ref.swin_v2(img, sd0, cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"],
                            cfg["use_shift"], cfg["out_indices"])
    enc = enc.cuda().eval()
    with torch.no_grad(), torch.autocast("cuda", torch.bfloat16):
        outs = enc(img.cuda())
        for n, p in enc.named_parameters():                 # the same pipeline with the block contributions switched off:
            if ".blocks." in n and "norm" in n:             # the bf16 noise of the base path (patch embedding, merging)
                p.zero_()                                   # cancels in the difference
        outs0 = enc(img.cuda())
    errs = []
    for o, o0, r, b0 in zip(outs, outs0, ref, base):
        contrib_ref = (r - b0).double()
        contrib = (o.double() - o0.double()).cpu()
        errs.append(((contrib - contrib_ref).norm() / contrib_ref.norm()).item())
    # stage 0: the four block contributions ride the fp32 stream untouched (measured 0.005; 0.52 with a bf16 stream).
    # stage 1 starts from a bf16 GEMM over the stage-0 stream (PatchMerging) -- as under torch.autocast, what stage 0 added
    # below a bf16 ulp cannot pass through that operand; only its own blocks' contributions are preserved.
    assert errs[0] < 0.05, errs
    assert errs[1] < 0.6, errs


@pytest.mark.parametrize("add_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,C", [(260, 256), (1201, 512), (7, 1024), (33, 128)])
def test_layer_norm_sum_vs_float64(rows, C, add_dtype):
    """y = LN(x + xadd) in one kernel (Transformer_Encoder's post-norm, models/cnn_transformer.py:202-203, :208-209): both
    outputs and every gradient against float64; the bf16 copy is the rounding of the fp32 output."""
    from b200swin import ops
    g = torch.Generator().manual_seed(rows + C)
    x = torch.randn(rows, C, generator=g).cuda().requires_grad_(True)
    xa = torch.randn(rows, C, generator=g).cuda().to(add_dtype).requires_grad_(True)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    beta = (0.2 * torch.randn(C, generator=g)).cuda().requires_grad_(True)
    # (the cotangent of the bf16 output travels as bf16: keep it representable)
    c32, c16 = torch.randn(rows, C, generator=g).cuda(), torch.randn(rows, C, generator=g).cuda().bfloat16().float()
    y32, y16 = ops.layer_norm_sum(x, xa, gamma, beta, 1e-5, want16=True)
    assert y32.dtype == torch.float32 and y16.dtype == torch.bfloat16 and torch.equal(y16, y32.detach().bfloat16())
    assert ops.bf16_twin_of(y32) is y16
    ((y32 * c32).sum() + (y16.float() * c16).sum()).backward()
    x6, a6, g6, b6 = [t.detach().double().cpu().requires_grad_(True) for t in (x, xa, gamma, beta)]
    y6 = torch.nn.functional.layer_norm(x6 + a6, (C,), g6, b6, 1e-5)
    (y6 * (c32 + c16).double().cpu()).sum().backward()

    def rel(a, r):
        return ((a.detach().double().cpu() - r).norm() / r.norm()).item()
    assert rel(y32, y6.detach()) < 1e-5
    assert rel(x.grad, x6.grad) < 1e-4 and rel(gamma.grad, g6.grad) < 1e-4 and rel(beta.grad, b6.grad) < 1e-4
    assert rel(xa.grad, a6.grad) < (1e-4 if add_dtype == torch.float32 else 5e-3)       # bf16 storage of the gradient
