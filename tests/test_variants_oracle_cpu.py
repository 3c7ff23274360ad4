"""Pin the oracle's SURVEY 8f-4 restatements (attn_type='normal', learned bias table, ConvMlp, the global-attention
encoder layer of models/cnn_transformer.py) against golden vectors generated from the reference itself
(tests/golden/make_golden.py variants tenc).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mha_ref, swin_ref


def _close(a, ref, rtol, atol_rel, msg=""):
    ref = np.asarray(ref)
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=atol_rel * max(1e-6, np.abs(ref).max()), err_msg=msg)


@pytest.mark.parametrize("name", ["wattn_normal_c64_h2_ws4_masked", "wattn_normal_none_c96_h3_ws6"])
def test_window_attention_normal_matches_reference(name):
    g = load_golden(name)
    C, nH, ws, _, B_, nW = g["meta.cfg"].tolist()
    at, rct, rot = [str(s) for s in g["meta.types"]]
    assert at == "normal"
    sd = swin_ref.npz_to_sd(g)
    assert "logit_scale" not in sd and (("relative_position_bias_table" in sd) == (rct == "none"))
    x = torch.from_numpy(g["in.x"]).requires_grad_(True)
    mask = torch.from_numpy(g["in.mask"]) if nW else None
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and k != "relative_coords_table"}
    full = dict(sd)
    full.update(params)
    y = swin_ref.window_attention(x, full, nH, mask, rpe_output_type=rot)
    _close(y.detach().numpy(), g["out.y"], 1e-5, 2e-6)
    names = sorted(params)
    grads = torch.autograd.grad((y * torch.from_numpy(g["in.cot"])).sum(), [x] + [params[n] for n in names])
    _close(grads[0].numpy(), g["grad.x"], 1e-4, 1e-5)
    for n, gr in zip(names, grads[1:]):
        _close(gr.numpy(), g["grad.sd." + n], 2e-4, 2e-5, n)


@pytest.mark.parametrize("name", ["layer_post_convln_c64_ws4_pad", "layer_pre_conv_normal_c64_ws4"])
def test_basic_layer_variants_match_reference(name):
    g = load_golden(name)
    dim, nH, ws, _, H, W, B, depth, down, post, shift, Wh, Ww = g["meta.cfg"].tolist()
    at, rct, rot, mt = [str(s) for s in g["meta.types"]]
    sd = swin_ref.npz_to_sd(g)
    assert any("mlp.conv_proj.weight" in k for k in sd) and (any("proj_ln" in k for k in sd) == (mt == "conv_ln"))
    x = torch.from_numpy(g["in.x"]).requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "relative_coords_table" not in k}
    full = dict(sd)
    full.update(params)
    x_out, _, _, x_down, Wh1, Ww1 = swin_ref.basic_layer(x, full, H, W, depth, nH, ws, bool(shift), bool(down), bool(post),
                                                         rpe_output_type=rot)
    assert (Wh1, Ww1) == (Wh, Ww)
    _close(x_out.detach().numpy(), g["out.x"], 2e-5, 2e-5)
    _close(x_down.detach().numpy(), g["out.x_down"], 2e-5, 2e-5)
    total = (x_out * torch.from_numpy(g["in.cot2"])).sum() + (x_down * torch.from_numpy(g["in.cot"])).sum()
    names = sorted(params)
    grads = torch.autograd.grad(total, [x] + [params[n] for n in names], allow_unused=True)
    _close(grads[0].numpy(), g["grad.x"], 1e-3, 1e-4)
    for n, gr in zip(names, grads[1:]):
        ref = g["grad.sd." + n]
        _close(gr.numpy() if gr is not None else np.zeros_like(ref), ref, 1e-3, 2e-4, n)


def test_transformer_encoder_layer_matches_reference():
    g = load_golden("tenc_h256_n130")
    B, N, E, nH, ff = g["meta.cfg"].tolist()
    sd = swin_ref.npz_to_sd(g)
    feat = torch.from_numpy(g["in.feat"]).requires_grad_(True)
    pos = torch.from_numpy(g["in.pos"]).requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y = mha_ref.transformer_encoder_layer(feat, pos, params, nH)
    _close(y.detach().numpy(), g["out.y"], 2e-5, 2e-5)
    names = sorted(params)
    grads = torch.autograd.grad((y * torch.from_numpy(g["in.cot"])).sum(), [feat, pos] + [params[n] for n in names])
    _close(grads[0].numpy(), g["grad.feat"], 1e-3, 1e-4)
    _close(grads[1].numpy(), g["grad.pos"], 1e-3, 1e-4)
    for n, gr in zip(names, grads[2:]):
        _close(gr.numpy(), g["grad.sd." + n], 1e-3, 2e-4, n)
    # the head-averaged weights the reference receives (and drops) at cnn_transformer.py:201
    sub = {k[len("self_attn."):]: v for k, v in sd.items() if k.startswith("self_attn.")}
    qk = (feat + pos).detach()
    _, w = mha_ref.multihead_attention(qk, qk, feat.detach(), sub, nH)
    _close(w.numpy(), g["out.weights"], 1e-4, 1e-5)


@pytest.mark.parametrize("E,nH,Nq,Nk", [(128, 2, 37, 37), (128, 4, 20, 45)])
def test_mha_oracle_equals_torch_multihead_attention(E, nH, Nq, Nk):
    """The restated algorithm against the third-party implementation the reference calls (torch.nn.MultiheadAttention)."""
    torch.manual_seed(E + Nk)
    m = torch.nn.MultiheadAttention(E, nH, batch_first=True).double()
    with torch.no_grad():
        m.in_proj_bias.normal_(0, 0.3)
        m.out_proj.bias.normal_(0, 0.3)
    q = torch.randn(2, Nq, E, dtype=torch.float64)
    k = torch.randn(2, Nk, E, dtype=torch.float64)
    v = torch.randn(2, Nk, E, dtype=torch.float64)
    y_ref, w_ref = m(q, k, v)
    y, w = mha_ref.multihead_attention(q, k, v, dict(m.state_dict()), nH)
    np.testing.assert_allclose(y.detach().numpy(), y_ref.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(w.detach().numpy(), w_ref.detach().numpy(), rtol=1e-10, atol=1e-12)
