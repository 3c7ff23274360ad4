"""Continuous-position-bias table kernel (cpb_mlp + 16*sigmoid) against the PyTorch restatement of
models/swin_transformer_v2.py:304-313, forward and all three parameter gradients (fp32, 1e-4 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("ws,nH,hid", [(12, 16, 512), (6, 32, 512), (24, 48, 512), (7, 3, 64)])
def test_cpb_table_forward_backward(ws, nH, hid):
    from b200swin import ops
    from b200swin.swin_transformer_v2 import WindowAttention
    torch.manual_seed(ws * 100 + nH)
    attn = WindowAttention(nH * 32, (ws, ws), nH, attn_type='cosine_mh', relative_coords_table_type='norm8_log_bylayer',
                           rpe_output_type='sigmoid', rpe_hidden_dim=hid, pretrain_window_size=ws).cuda()
    torch.nn.init.normal_(attn.rpe_mlp[2].weight, std=0.3)
    l0, l2 = attn.rpe_mlp[0], attn.rpe_mlp[2]
    coords = attn.relative_coords_table
    T = (2 * ws - 1) ** 2
    cot = torch.randn(T, nH, device="cuda", dtype=torch.float64)
    # reference in float64
    w0, b0, w2 = (p.detach().double().requires_grad_(True) for p in (l0.weight, l0.bias, l2.weight))
    ref = 16 * torch.sigmoid(torch.relu(coords.double().view(-1, 2) @ w0.t() + b0) @ w2.t())
    gr = torch.autograd.grad((ref * cot).sum(), [w0, b0, w2])
    out = ops.cpb_table(coords, l0.weight, l0.bias, l2.weight)
    assert out.shape == (T, nH) and out.dtype == torch.float32
    (out.double() * cot).sum().backward()
    for a, r, name in [(out, ref, "table"), (l0.weight.grad, gr[0], "dW0"), (l0.bias.grad, gr[1], "db0"),
                       (l2.weight.grad, gr[2], "dW2")]:
        err = (a.double() - r).norm() / r.norm().clamp_min(1e-30)
        assert err <= 1e-4, f"{name}: rel-L2 {err:.3e}"
    # the module takes the same route
    assert torch.equal(attn._bias_table(), out)


def test_temperature_clamp_exp_forward_backward():
    """scale = exp(min(logit_scale, ln 100)) and its gradient (zero where clamped) out of the CPB kernels
    (models/swin_transformer_v2.py:294): against torch.clamp(...).exp() in float64."""
    from b200swin.swin_transformer_v2 import WindowAttention
    import math
    torch.manual_seed(3)
    attn = WindowAttention(8 * 32, (6, 6), 8, attn_type='cosine_mh', relative_coords_table_type='norm8_log_bylayer',
                           rpe_output_type='sigmoid', pretrain_window_size=6).cuda()
    with torch.no_grad():
        attn.logit_scale.copy_(torch.tensor([0.5, 2.3, 4.0, 4.6, 4.7, 6.0, -1.0, math.log(100.0)]).view(8, 1, 1))
    table, scale = attn._table_and_scale()
    # the reference clamps in float32 (the parameter's dtype): the head sitting exactly on ln 100 keeps its gradient
    ref_in = attn.logit_scale.detach().clone().requires_grad_(True)
    ref = torch.clamp(ref_in, max=math.log(1.0 / 0.01)).exp().view(8).double()
    cot = torch.randn(8, device="cuda", dtype=torch.float64)
    (scale.double() * cot).sum().backward()
    (ref * cot).sum().backward()
    assert ((scale.double() - ref).abs() / ref).max() < 1e-5
    g, gr = attn.logit_scale.grad.double().view(-1), ref_in.grad.double().view(-1)
    assert (g - gr).abs().max() <= 1e-5 * gr.abs().max()
    assert g[4] == 0 and g[5] == 0                  # clamped heads receive no gradient
    assert torch.equal(table, attn._bias_table())
