"""tcgen05 GEMM parity (GPU): every operand-layout combination the block uses (forward, dgrad with W read
MN-major, wgrad with both operands MN-major + split-K), the fp32-accurate hi/lo mode and the fused
epilogues, against float64 matmuls of the same inputs.
Tolerances: fp32-accurate mode 1e-4 relative (north_star fp32 bar), bf16 2e-2."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _relerr(a, ref):
    a, ref = a.double().cpu(), ref.double().cpu()
    return ((a - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def _ops():
    from b200swin import ops, _lib
    return ops, _lib


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 128), (1000, 288, 96), (4096 + 40, 512, 2048),
                                   (14400, 384, 128), (77, 96, 32)])
def test_nt_bf16(M, N, K):
    ops, L = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).bfloat16().cuda()
    b = torch.randn(N, K, generator=g).bfloat16().cuda()
    ref = a.double() @ b.double().t()
    out32 = ops.gemm(ops.Operand(a), ops.Operand(b), M, N, K, out_dtype=torch.float32)
    assert _relerr(out32, ref) < 5e-6          # fp32 accumulation over K
    out16 = ops.gemm(ops.Operand(a), ops.Operand(b), M, N, K, out_dtype=torch.bfloat16)
    assert _relerr(out16, ref) < 4e-3


@pytest.mark.parametrize("M,N,K", [(512, 256, 384), (1000, 96, 288), (14400, 128, 384)])
def test_dgrad_layout_b_mn_major(M, N, K):
    # dX[M,N] = dY[M,K] . W[K,N]  with W stored [K,N] (i.e. nn.Linear.weight [out,in] read MN-major)
    ops, L = _ops()
    g = torch.Generator().manual_seed(1 + M)
    dy = torch.randn(M, K, generator=g).bfloat16().cuda()
    w = torch.randn(K, N, generator=g).bfloat16().cuda()
    ref = dy.double() @ w.double()
    out = ops.gemm(ops.Operand(dy), ops.Operand(w), M, N, K, b_mn=True, out_dtype=torch.float32)
    assert _relerr(out, ref) < 2e-6


@pytest.mark.parametrize("T,N,K,splits", [(1024, 384, 128, 1), (14400, 384, 128, 0), (5000, 288, 96, 7),
                                          (43200, 2048, 512, 0), (136, 128, 128, 3)])
def test_wgrad_layout_both_mn_major_splitk(T, N, K, splits):
    # dW[N,K] = dY[T,N]^T . X[T,K]: both operands read as stored (MN-major), contraction over tokens
    ops, L = _ops()
    g = torch.Generator().manual_seed(2 + T)
    dy = torch.randn(T, N, generator=g).bfloat16().cuda()
    x = torch.randn(T, K, generator=g).bfloat16().cuda()
    ref = dy.double().t() @ x.double()
    if splits == 0:
        splits = L.load().b200swin_gemm_splits(N, K, T)
    out = ops.gemm(ops.Operand(dy), ops.Operand(x), N, K, T, a_mn=True, b_mn=True, out_dtype=torch.float32,
                   splits=splits)
    assert _relerr(out, ref) < 2e-5          # fp32 accumulation over up to 43200 tokens


@pytest.mark.parametrize("M,N,K", [(640, 384, 128), (1000, 512, 512)])
def test_fp32_accurate_mode(M, N, K):
    ops, L = _ops()
    g = torch.Generator().manual_seed(3 + M)
    a = torch.randn(M, K, generator=g).cuda()
    b = (torch.randn(N, K, generator=g) * 0.05).cuda()
    ref = a.double() @ b.double().t()
    ao, bo = ops.stage_operand(a, True), ops.stage_operand(b, True)
    # hi + lo reproduces the fp32 value to ~2^-17
    assert _relerr(ao.hi.float() + ao.lo.float(), a) < 1e-5
    out = ops.gemm(ao, bo, M, N, K, out_dtype=torch.float32)
    assert _relerr(out, ref) < 2e-5
    # wgrad-style and dgrad-style in the accurate mode
    bt = b.t().contiguous()                                  # [K,N]
    out2 = ops.gemm(ao, ops.stage_operand(bt, True), M, N, K, b_mn=True, out_dtype=torch.float32)
    assert _relerr(out2, ref) < 2e-5
    dy = torch.randn(M, N, generator=g).cuda()
    refw = dy.double().t() @ a.double()
    outw = ops.gemm(ops.stage_operand(dy, True), ao, N, K, M, a_mn=True, b_mn=True, out_dtype=torch.float32, splits=3)
    assert _relerr(outw, refw) < 2e-5


def test_epilogues():
    ops, L = _ops()
    g = torch.Generator().manual_seed(5)
    M, C, nH = 300, 128, 4
    x = torch.randn(M, C, generator=g).bfloat16().cuda()
    # bias
    w = (torch.randn(C, C, generator=g) * 0.1).bfloat16().cuda()
    bias = torch.randn(C, generator=g).cuda()
    ref = x.double() @ w.double().t() + bias.double()
    out = ops.gemm(ops.Operand(x), ops.Operand(w), M, C, C, bias=bias, out_dtype=torch.float32)
    assert _relerr(out, ref) < 2e-6
    # GELU with the derivative gelu'(pre-activation) as side output
    w1 = (torch.randn(4 * C, C, generator=g) * 0.1).bfloat16().cuda()
    b1 = torch.randn(4 * C, generator=g).cuda()
    z_ref = x.double() @ w1.double().t() + b1.double()
    h_ref = 0.5 * z_ref * (1 + torch.erf(z_ref / math.sqrt(2)))
    z = torch.empty(M, 4 * C, dtype=torch.float32, device="cuda")
    h = ops.gemm(ops.Operand(x), ops.Operand(w1), M, 4 * C, C, epilogue=L.EPI_GELU, bias=b1, aux_out=z,
                 out_dtype=torch.float32)
    gp = 0.5 * (1 + torch.erf(z_ref / math.sqrt(2))) + z_ref * torch.exp(-0.5 * z_ref ** 2) / math.sqrt(2 * math.pi)
    assert _relerr(z, gp) < 3e-6 and _relerr(h, h_ref) < 3e-6
    # DGELU: out = acc * aux (the saved gelu')
    dy = torch.randn(M, C, generator=g).bfloat16().cuda()
    w2 = (torch.randn(C, 4 * C, generator=g) * 0.1).bfloat16().cuda()      # fc2.weight [C, 4C] read MN-major
    zz = gp.float().cuda()
    dz_ref = (dy.double() @ w2.double()) * zz.double()
    dz = ops.gemm(ops.Operand(dy), ops.Operand(w2), M, 4 * C, C, b_mn=True, epilogue=L.EPI_DGELU, aux_in=zz,
                  out_dtype=torch.float32)
    assert _relerr(dz, dz_ref) < 1e-5
    # the bf16 variant prefetches the aux tile by TMA (ragged M: 300 rows)
    dz16 = ops.gemm(ops.Operand(dy), ops.Operand(w2), M, 4 * C, C, b_mn=True, epilogue=L.EPI_DGELU, aux_in=zz.bfloat16(),
                    out_dtype=torch.bfloat16)
    assert _relerr(dz16, (dy.double() @ w2.double()) * zz.bfloat16().double()) < 4e-3
    # QKV: biases on q and v only, q/k L2-normalised per 32-wide head, inverse norms exported
    wq = (torch.randn(3 * C, C, generator=g) * 0.1).bfloat16().cuda()
    qb, vb = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    raw = x.double() @ wq.double().t() + torch.cat([qb, torch.zeros_like(vb), vb]).double()
    q, k, v = raw[:, :C].view(M, nH, 32), raw[:, C:2 * C].view(M, nH, 32), raw[:, 2 * C:]
    ref_qkv = torch.cat([(q / q.norm(dim=-1, keepdim=True)).view(M, C), (k / k.norm(dim=-1, keepdim=True)).view(M, C), v], 1)
    inv = torch.empty(M, 2, nH, dtype=torch.float32, device="cuda")
    out = ops.gemm(ops.Operand(x), ops.Operand(wq), M, 3 * C, C, epilogue=L.EPI_QKV, bias=qb, bias2=vb, inv_norm=inv,
                   nH=nH, out_dtype=torch.float32)
    assert _relerr(out, ref_qkv) < 3e-6
    assert _relerr(inv[:, 0], 1 / q.norm(dim=-1)) < 3e-6 and _relerr(inv[:, 1], 1 / k.norm(dim=-1)) < 3e-6


def test_colsum_and_split():
    ops, L = _ops()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(5000, 384, generator=g).cuda()
    assert _relerr(ops.colsum(x), x.double().sum(0)) < 1e-6
    assert _relerr(ops.colsum(x, 256, 128), x[:, 256:].double().sum(0)) < 1e-6
    xb = x.bfloat16()
    assert _relerr(ops.colsum(xb, 0, 128, extra=torch.ones(128, device="cuda")), xb[:, :128].double().sum(0) + 1) < 1e-6


def test_linear_and_mlp_autograd_match_fp64():
    import torch.nn.functional as F
    ops, L = _ops()
    g = torch.Generator().manual_seed(8)
    B, T, C = 2, 150, 128
    x = torch.randn(B, T, C, generator=g)
    w1, b1 = torch.randn(4 * C, C, generator=g) * 0.1, torch.randn(4 * C, generator=g) * 0.1
    w2, b2 = torch.randn(C, 4 * C, generator=g) * 0.1, torch.randn(C, generator=g) * 0.1
    cot = torch.randn(B, T, C, generator=g)
    leaves = [t.double().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    ref = F.linear(F.gelu(F.linear(leaves[0], leaves[1], leaves[2])), leaves[3], leaves[4])
    gref = torch.autograd.grad((ref * cot.double()).sum(), leaves)
    for mode, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        mine = [t.clone().cuda().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
        if mode == "bf16":
            with torch.autocast("cuda", torch.bfloat16):
                y = ops.mlp(*mine)
        else:
            y = ops.mlp(*mine)
        assert _relerr(y, ref.detach()) < tol, mode
        (y.float() * cot.cuda()).sum().backward()
        for m, r, nm in zip(mine, gref, ["dx", "dw1", "db1", "dw2", "db2"]):
            assert _relerr(m.grad, r) < tol, (mode, nm, _relerr(m.grad, r))
        # plain linear
        mine = [t.clone().cuda().requires_grad_(True) for t in (x, w1, b1)]
        yl = ops.linear(*mine) if mode == "fp32" else None
        if yl is not None:
            l2 = [t.double().requires_grad_(True) for t in (x, w1, b1)]
            rl = F.linear(*l2)
            c2 = torch.randn(rl.shape, generator=g)
            gl = torch.autograd.grad((rl * c2.double()).sum(), l2)
            (yl * c2.cuda()).sum().backward()
            assert _relerr(yl, rl.detach()) < 1e-4
            for m, r in zip(mine, gl):
                assert _relerr(m.grad, r) < 1e-4
