"""Window gather/scatter/mask kernels: bit-exact index maps against the reference-generated golden
maps and the numpy oracle, round trips at full BASELINE sizes, and adjointness."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import index_maps as im

pytestmark = pytest.mark.gpu


def test_index_maps_bit_exact_vs_golden():
    from b200swin import ops
    g = load_golden("index_maps")
    for ci, (B, H, W, ws) in enumerate(g["meta.cases"].tolist()):
        s = ws // 2
        x = (torch.arange(B * H * W, dtype=torch.float32) + 1).view(B, H, W, 1).cuda()
        for shift, key in [(0, "partition"), (s, "gather_shift")]:
            out = ops.window_gather(x, ws, shift).view(-1, ws * ws).long().cpu().numpy() - 1
            assert np.array_equal(out, g[f"c{ci}.{key}"]), (ci, key)
        nslots = g[f"c{ci}.gather_shift"].size
        slots = torch.arange(nslots, dtype=torch.float32).view(-1, ws * ws, 1).cuda()
        back = ops.window_scatter(slots, B, H, W, ws, s).view(B, H, W).long().cpu().numpy()
        assert np.array_equal(back, g[f"c{ci}.scatter_shift"]), (ci, "scatter")
        m = ops.shift_mask(H, W, ws, s, "cuda").cpu().numpy()
        assert np.array_equal(m, g[f"c{ci}.mask"]), (ci, "mask")


@pytest.mark.parametrize("B,H,W,C,ws,dtype", [(4, 120, 120, 128, 12, torch.bfloat16), (2, 30, 30, 512, 12, torch.float32),
                                              (2, 88, 304, 192, 24, torch.bfloat16), (3, 15, 15, 96, 6, torch.float32),
                                              (1, 5, 9, 6, 4, torch.bfloat16)])
def test_round_trip_and_oracle_full_size(B, H, W, C, ws, dtype):
    from b200swin import ops
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(B, H, W, C, generator=gen).to(dtype).cuda()
    for shift in (0, ws // 2):
        win = ops.window_gather(x, ws, shift)
        idx = torch.from_numpy(im.fused_gather_index(B, H, W, ws, shift)).cuda()
        flat = torch.cat([x.view(-1, C), torch.zeros(1, C, dtype=dtype, device="cuda")])
        assert torch.equal(win, flat[idx.view(-1)].view(win.shape))
        assert torch.equal(ops.window_scatter(win, B, H, W, ws, shift), x)      # crop(unroll(reverse(.))) inverts


def test_adjoint_through_autograd():
    from b200swin import ops
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 10, 7, 8, generator=gen).cuda().requires_grad_(True)
    w = ops.window_gather(x, 4, 2)
    cot = torch.randn(w.shape, generator=gen).cuda()
    (w * cot).sum().backward()
    # <gather(x), cot> == <x, scatter(cot)>
    assert torch.equal(x.grad, ops.window_scatter(cot, 2, 10, 7, 4, 2))


def _patch_merge_ref(x):
    """The reference's PatchMerging gather (models/swin_transformer_v2.py:660-672), torch ops on the same device."""
    B, H, W, C = x.shape
    if H % 2 == 1 or W % 2 == 1:
        x = torch.nn.functional.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
    return x.view(B, -1, 4 * C)


@pytest.mark.parametrize("B,H,W,C,dtype", [(4, 120, 120, 128, torch.bfloat16), (2, 15, 15, 512, torch.bfloat16),
                                           (3, 30, 29, 96, torch.float32), (1, 7, 12, 8, torch.bfloat16),
                                           (2, 1, 1, 4, torch.float32)])
def test_patch_merge_bit_exact_forward_and_adjoint(B, H, W, C, dtype):
    from b200swin import ops
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(B, H, W, C, generator=gen).to(dtype).cuda().requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    out = ops.patch_merge(x)
    ref = _patch_merge_ref(xr)
    assert out.shape == ref.shape and torch.equal(out, ref)
    cot = torch.randn(ref.shape, generator=gen).to(dtype).cuda()
    out.backward(cot)
    ref.backward(cot)
    assert torch.equal(x.grad, xr.grad)


@pytest.mark.parametrize("B,H,W,E,dtype", [(2, 480, 480, 128, torch.float32), (2, 64, 96, 96, torch.bfloat16),
                                           (1, 30, 41, 128, torch.float32),
                                           (1, 8, 1100, 32, torch.float32)])     # wide frame: the per-thread patchify
def test_patch_embed_gemm_matches_conv(B, H, W, E, dtype):
    """PatchEmbed as patchify + GEMM against the conv it replaces (models/swin_transformer_v2.py:941-957), forward and
    the weight / bias gradients; fp32 within 1e-4, bf16 (autocast) within 2e-2 relative."""
    from b200swin.swin_transformer_v2 import PatchEmbed
    torch.manual_seed(7)
    pe = PatchEmbed(4, 3, E, None).cuda()
    x = torch.rand(B, 3, H, W, device="cuda")
    ph = 4
    xp = torch.nn.functional.pad(x, (0, (-W) % ph, 0, (-H) % ph))
    ref = torch.nn.functional.conv2d(xp.double(), pe.proj.weight.double(), pe.proj.bias.double(), stride=4)
    cot = torch.randn(ref.shape, device="cuda", dtype=torch.float64)
    gw, gb = torch.autograd.grad((ref * cot).sum(), [pe.proj.weight, pe.proj.bias])
    with torch.autocast("cuda", torch.bfloat16, enabled=dtype == torch.bfloat16):
        out = pe(x)
    assert out.shape == ref.shape
    (out.double() * cot).sum().backward()
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for a, r, name in [(out, ref, "out"), (pe.proj.weight.grad, gw, "dW"), (pe.proj.bias.grad, gb, "db")]:
        err = (a.double() - r).norm() / r.norm()
        assert err <= tol, f"{name}: rel-L2 {err:.3e}"
