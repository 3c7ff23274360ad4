"""Pin the CPU oracle against the golden vectors generated from the reference itself
(tests/golden/make_golden.py).  CPU only; no reference needed at run time."""
import numpy as np
import pytest
import torch

from oracle import index_maps as im
from oracle import silog_ref, swin_ref
from conftest import load_golden


# ------------------------------------------------------------------ integer maps: bit exact
def test_index_maps_bit_exact():
    g = load_golden("index_maps")
    for ci, (B, H, W, ws) in enumerate(g["meta.cases"].tolist()):
        s = ws // 2
        Hp, Wp = im.padded_size(H, W, ws)
        part = g[f"c{ci}.partition"]
        # reference partition ran on the padded grid with pads = -1; rebuild from our map
        mine = im.fused_gather_index(B, H, W, ws, 0)
        assert np.array_equal(mine, part), (ci, "partition")
        assert np.array_equal(im.fused_gather_index(B, H, W, ws, s), g[f"c{ci}.gather_shift"]), (ci, "gather")
        assert np.array_equal(im.reverse_src_index(B, Hp, Wp, ws), g[f"c{ci}.reverse"]), (ci, "reverse")
        assert np.array_equal(im.fused_scatter_index(B, H, W, ws, s), g[f"c{ci}.scatter_shift"]), (ci, "scatter")
        assert np.array_equal(im.shift_attn_mask(H, W, ws, s), g[f"c{ci}.mask"]), (ci, "mask")
        assert np.array_equal(im.shift_region_ids(Hp, Wp, ws, s), im.shift_region_ids_closed_form(Hp, Wp, ws, s))


def test_partition_is_plain_window_partition():
    # no padding: fused gather with shift 0 equals the pure partition map
    for (B, Hp, Wp, ws) in [(2, 8, 12, 4), (1, 24, 24, 12)]:
        assert np.array_equal(im.fused_gather_index(B, Hp, Wp, ws, 0), im.partition_src_index(B, Hp, Wp, ws))


def test_known_answers_from_survey():
    assert im.relative_position_index(2, 2).tolist() == [[4, 3, 1, 0], [5, 4, 2, 1], [7, 6, 4, 3], [8, 7, 5, 4]]
    ids = im.shift_region_ids(8, 8, 4, 2)
    assert ids[0].tolist() == [0, 0, 0, 0, 1, 1, 2, 2] and ids[4].tolist() == [3, 3, 3, 3, 4, 4, 5, 5]
    assert ids[6].tolist() == [6, 6, 6, 6, 7, 7, 8, 8]
    m = im.shift_attn_mask(8, 8, 4, 2)
    assert [(m[w] != 0).sum() for w in range(4)] == [0, 128, 128, 192]
    row = im.relative_coords_table(3, 3, 2)[0, :, 0, 0]
    np.testing.assert_allclose(row, [-1.362488, -1.056642, 0, 1.056642, 1.362488], atol=2e-6)


def test_buffers_match_reference():
    g = load_golden("index_maps")
    for k in g.files:
        if k.startswith("rpi.ws"):
            ws = int(k[len("rpi.ws"):])
            assert np.array_equal(im.relative_position_index(ws, ws), g[k])
        if k.startswith("rct.ws"):
            ws, pre = k[len("rct.ws"):].split(".pre")
            np.testing.assert_allclose(im.relative_coords_table(int(ws), int(ws), int(pre)), g[k], rtol=0, atol=2e-7)


# ------------------------------------------------------------------ WindowAttention fwd/bwd
@pytest.mark.parametrize("name", ["wattn_c64_h2_ws4_masked", "wattn_c96_h3_ws6_pre12", "wattn_c128_h4_ws12"])
def test_window_attention_matches_reference(name):
    g = load_golden(name)
    C, nH, ws, pre, B_, nW = g["meta.cfg"].tolist()
    sd = swin_ref.npz_to_sd(g)
    x = torch.from_numpy(g["in.x"]).requires_grad_(True)
    mask = torch.from_numpy(g["in.mask"]) if nW else None
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()
              and k not in ("relative_coords_table",)}
    full = dict(sd)
    full.update(params)
    y = swin_ref.window_attention(x, full, nH, mask)
    np.testing.assert_allclose(y.detach().numpy(), g["out.y"], rtol=1e-5, atol=2e-6)
    cot = torch.from_numpy(g["in.cot"])
    names = sorted(params)
    grads = torch.autograd.grad((y * cot).sum(), [x] + [params[n] for n in names])
    scale = np.abs(g["grad.x"]).max()
    np.testing.assert_allclose(grads[0].numpy(), g["grad.x"], rtol=1e-4, atol=1e-5 * scale)
    for n, gr in zip(names, grads[1:]):
        ref = g["grad.sd." + n]
        np.testing.assert_allclose(gr.numpy(), ref, rtol=2e-4, atol=2e-5 * max(1e-6, np.abs(ref).max()), err_msg=n)
    # hand-derived backward (appendix A) against the same golden gradients, in fp64
    sd64 = swin_ref.to_dtype(sd, torch.float64)
    b = swin_ref.window_attention_backward(cot.double(), x.detach().double(), sd64, nH,
                                           mask.double() if mask is not None else None)
    for mine, key in [("dx", "grad.x"), ("d_qkv_w", "grad.sd.qkv.weight"), ("d_q_bias", "grad.sd.q_bias"),
                      ("d_v_bias", "grad.sd.v_bias"), ("d_proj_w", "grad.sd.proj.weight"),
                      ("d_proj_b", "grad.sd.proj.bias")]:
        ref = g[key]
        np.testing.assert_allclose(b[mine].reshape(ref.shape).numpy(), ref, rtol=2e-4,
                                   atol=3e-5 * np.abs(ref).max(), err_msg=mine)
    ref = g["grad.sd.logit_scale"].reshape(-1)
    np.testing.assert_allclose(b["d_logit_scale"].numpy(), ref, rtol=2e-4, atol=3e-5 * np.abs(ref).max())
    assert ref[0] == 0.0     # head 0 sits above the clamp (swin_transformer_v2.py:294)
    # table-space gradient chains into rpe_mlp.2.weight: d_table^T @ hidden
    hid = torch.relu(torch.nn.functional.linear(sd64["relative_coords_table"].reshape(-1, 2),
                                                sd64["rpe_mlp.0.weight"], sd64["rpe_mlp.0.bias"]))
    ref = g["grad.sd.rpe_mlp.2.weight"]
    np.testing.assert_allclose((b["d_table"].t() @ hid).numpy(), ref, rtol=2e-4, atol=3e-5 * np.abs(ref).max())


# ------------------------------------------------------------------ BasicLayer (blocks, mask, merging)
LAYERS = ["layer_post_c64_ws4_pad", "layer_post_c32_ws6_nopad", "layer_pre_c64_ws4",
          "layer_post_c64_ws4_noshift", "layer_post_c128_ws12_pad"]


@pytest.mark.parametrize("name", LAYERS)
def test_basic_layer_matches_reference(name):
    g = load_golden(name)
    dim, nH, ws, pre, H, W, B, depth, down, post, shift, Wh, Ww = g["meta.cfg"].tolist()
    sd = swin_ref.npz_to_sd(g)
    x = torch.from_numpy(g["in.x"]).requires_grad_(True)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and "relative_coords_table" not in k}
    full = dict(sd)
    full.update(params)
    x_out, H1, W1, x_down, Wh1, Ww1 = swin_ref.basic_layer(x, full, H, W, depth, nH, ws, bool(shift), bool(down),
                                                            bool(post))
    assert (H1, W1, Wh1, Ww1) == (H, W, Wh, Ww)
    np.testing.assert_allclose(x_out.detach().numpy(), g["out.x"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(x_down.detach().numpy(), g["out.x_down"], rtol=2e-5, atol=2e-5)
    total = (x_out * torch.from_numpy(g["in.cot2"])).sum()
    if down:
        total = total + (x_down * torch.from_numpy(g["in.cot"])).sum()
    names = sorted(params)
    grads = torch.autograd.grad(total, [x] + [params[n] for n in names], allow_unused=True)
    ref = g["grad.x"]
    np.testing.assert_allclose(grads[0].numpy(), ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max())
    for n, gr in zip(names, grads[1:]):
        ref = g["grad.sd." + n]
        mine = gr.numpy() if gr is not None else np.zeros_like(ref)
        np.testing.assert_allclose(mine, ref, rtol=1e-3, atol=2e-4 * max(1e-6, np.abs(ref).max()), err_msg=n)


def test_swin_small_matches_reference():
    import json
    g = load_golden("swin_small")
    cfg = json.loads(str(g["meta.cfg_json"]))
    sd = swin_ref.npz_to_sd(g)
    img = torch.from_numpy(g["in.img"])
    outs = swin_ref.swin_v2(img, sd, cfg["embed_dim"], cfg["depths"], cfg["num_heads"], cfg["window_size"],
                            cfg["use_shift"], tuple(cfg["out_indices"]))
    assert len(outs) == 2
    for i, o in enumerate(outs):
        np.testing.assert_allclose(o.numpy(), g[f"out.{i}"], rtol=1e-4, atol=5e-5)


# ------------------------------------------------------------------ SiLog + metrics
def test_silog_matches_reference():
    g = load_golden("silog")
    assert abs(float(g["kat.loss"][0]) - 0.5659523) < 1e-6
    for nm in ["kat", "nyu", "kitti", "void", "allvalid"]:
        pred, tgt = g[f"{nm}.pred"], g[f"{nm}.target"]
        np.testing.assert_allclose(silog_ref.silog_np(pred, tgt), float(g[f"{nm}.loss"][0]), rtol=2e-6)
        np.testing.assert_allclose(silog_ref.silog_grad_np(pred, tgt, gout=1.7), g[f"{nm}.grad_x1p7"],
                                   rtol=2e-4, atol=1e-9)
        lt = silog_ref.silog_torch(torch.from_numpy(pred), torch.from_numpy(tgt))
        np.testing.assert_allclose(float(lt), float(g[f"{nm}.loss"][0]), rtol=1e-6)
    assert np.isnan(g["allinvalid.loss"][0]) and np.isnan(silog_ref.silog_np(np.ones(4), np.zeros(4)))


def test_eval_depth_matches_reference():
    g = load_golden("silog")
    keys = [str(k) for k in g["metrics.keys"]]
    m = silog_ref.eval_depth_np([1.0, 2.0, 3.0], [1.0, 1.0, 6.0])
    np.testing.assert_allclose([m[k] for k in keys], g["metrics.kat"], rtol=1e-6)
    assert abs(m["d1"] - 1 / 3) < 1e-7 and abs(m["abs_rel"] - 0.5) < 1e-7 and abs(m["rmse"] - 1.8257419) < 1e-6
    m = silog_ref.eval_depth_np(g["metrics.rand.pred"], g["metrics.rand.target"])
    np.testing.assert_allclose([m[k] for k in keys], g["metrics.rand"], rtol=2e-5)
