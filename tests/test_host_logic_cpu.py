"""Host-side logic of the drop-in modules that needs no GPU: stochastic-depth mask drawing (timm semantics,
reference models/swin_transformer_v2.py DropPath usage in the blocks :419-488) and constructor validation."""
import pytest
import torch

from b200swin.swin_transformer_v2 import DropPath, SwinTransformerV2

CFG = dict(embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=[4, 4], pretrain_window_size=[4, 4],
           use_shift=[True, True], out_indices=(0, 1))


def _drop_paths(net):
    return [m for m in net.modules() if isinstance(m, DropPath)]


def test_drop_path_rates_follow_the_linear_schedule():
    net = SwinTransformerV2(drop_path_rate=0.3, **CFG)
    rates = sorted(m.drop_prob for m in _drop_paths(net))
    assert rates == pytest.approx([0.1, 0.2, 0.3], abs=1e-6)     # linspace(0, rate, sum(depths)); p = 0 is an Identity


def test_masks_of_a_forward_are_drawn_in_one_batch_and_consumed_in_call_order():
    torch.manual_seed(3)
    net = SwinTransformerV2(drop_path_rate=0.5, **CFG).train()
    B = 64
    mods = net._draw_drop_paths(B, torch.device("cpu"))
    assert [m.drop_prob for m in mods] == pytest.approx([1 / 6, 1 / 3, 0.5], abs=1e-6)   # the p = 0 block draws nothing
    x = torch.zeros(B, 3)
    for m in mods:
        keep = 1.0 - m.drop_prob
        a, b = m.sample_scale(x), m.sample_scale(x)                   # two draws per block: attention and MLP branch
        for s in (a, b):
            assert s.shape == (B,) and s.dtype == torch.float32
            assert all(v == 0.0 or v == pytest.approx(1.0 / keep) for v in s.unique().tolist())   # Bernoulli(keep) / keep
        assert not torch.equal(a, b)
        assert m._drawn == []
        c = m.sample_scale(x)                                         # a third call falls back to its own draw
        assert c.shape == (B,)
    big = net._draw_drop_paths(20000, torch.device("cpu"))
    for m in big:                                                     # E[scale] = 1
        assert m._drawn[0].mean().item() == pytest.approx(1.0, abs=0.05)


def test_pre_drawn_masks_are_ignored_for_another_batch_size_and_in_eval():
    net = SwinTransformerV2(drop_path_rate=0.5, **CFG).train()
    mods = net._draw_drop_paths(8, torch.device("cpu"))
    m = mods[-1]
    s = m.sample_scale(torch.zeros(5, 2))                             # stale batch size: own draw, queue untouched
    assert s.shape == (5,) and len(m._drawn) == 2
    net.eval()
    assert net._draw_drop_paths(8, torch.device("cpu")) == []
    assert m.sample_scale(torch.zeros(8, 2)) is None


def test_reference_rng_mode_reproduces_timm_drop_path_masks():
    """reference_rng=True: per-call draws in the reference's order with timm's call -- x.new_empty((B, 1, 1)).bernoulli_(keep)
    per branch (timm.models.layers.DropPath as used at swin_transformer_v2.py:473 / :483) -- so the same seed gives the
    same masks as the reference."""
    net = SwinTransformerV2(drop_path_rate=0.5, reference_rng=True, **CFG).train()
    B = 16
    assert net._draw_drop_paths(B, torch.device("cpu")) == []
    x = torch.zeros(B, 7, 3)
    torch.manual_seed(11)
    mine = []
    for blk in [b for l in net.layers for b in l.blocks]:
        if isinstance(blk.drop_path, DropPath):
            mine += [blk.drop_path.sample_scale(x), blk.drop_path.sample_scale(x)]      # attention branch, MLP branch
    torch.manual_seed(11)
    k = 0
    for blk in [b for l in net.layers for b in l.blocks]:
        if isinstance(blk.drop_path, DropPath):
            keep = 1.0 - blk.drop_path.drop_prob
            for _ in range(2):
                ref = x.new_empty((B, 1, 1)).bernoulli_(keep).div_(keep)
                assert torch.equal(mine[k], ref.view(B))
                k += 1
    assert k == len(mine) == 6


def test_no_batched_draw_under_activation_checkpointing():
    """The recompute of a checkpointed block replays the RNG of per-call draws; pre-drawn masks would be gone by then."""
    net = SwinTransformerV2(drop_path_rate=0.5, use_checkpoint=True, **CFG).train()
    assert net._draw_drop_paths(8, torch.device("cpu")) == []


def test_forward_clears_the_queues_even_when_it_raises():
    net = SwinTransformerV2(drop_path_rate=0.5, **CFG).train()
    with pytest.raises(Exception):                                    # CPU tensors: the CUDA ops refuse loudly
        net(torch.rand(2, 3, 32, 32))
    assert all(m._drawn == [] for m in _drop_paths(net))


def test_pad_queries_are_normalised_once_per_stage():
    """Padded grids: every block's pad-token query (normalize(q_bias) per head) comes from one batched call and equals
    the per-block computation; nothing is staged when the grid needs no padding."""
    torch.manual_seed(5)
    net = SwinTransformerV2(drop_path_rate=0.0, **CFG)
    layer = net.layers[1]
    with torch.no_grad():
        for blk in layer.blocks:
            blk.attn.q_bias.normal_()
    assert layer._stage_pad_queries(8, 8) == []                       # 8 % 4 == 0: no padding
    staged = layer._stage_pad_queries(7, 9)
    assert len(staged) == len(layer.blocks)
    for a in staged:
        batched, _ = a._pads(True)
        a._qpad_pre = None
        own, _ = a._pads(True)
        assert batched.shape == own.shape == (a.dim,)
        assert torch.allclose(batched, own, rtol=0, atol=1e-7)
        assert a._pads(False) == (None, None)


def test_cnn_transformer_dropins_share_the_reference_state_dict_keys():
    """b200swin.cnn_transformer.Transformer_Encoder / MultiheadAttention expose the keys of the reference's
    Transformer_Encoder (golden fixture generated from models/cnn_transformer.py:176-216) and of torch's
    nn.MultiheadAttention, with the same shapes."""
    import types
    import numpy as np
    import torch
    from conftest import load_golden
    from b200swin.cnn_transformer import MultiheadAttention, Transformer_Encoder
    g = load_golden("tenc_h256_n130")
    B, N, E, nH, ff = g["meta.cfg"].tolist()
    enc = Transformer_Encoder(types.SimpleNamespace(transformer_ff_dim=ff), hidden_dim=E)
    ref = {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
    mine = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert mine == ref
    assert enc.self_attn.num_heads == nH
    t = torch.nn.MultiheadAttention(512, 8, batch_first=True)
    m = MultiheadAttention(512, 8, batch_first=True)
    assert {k: v.shape for k, v in t.state_dict().items()} == {k: v.shape for k, v in m.state_dict().items()}
    m.load_state_dict(t.state_dict(), strict=True)


def test_variant_constructors_build_the_reference_keys():
    """attn_type='normal' / relative_coords_table_type='none' / mlp_type='conv_ln': same state_dict keys as the reference
    modules that produced the golden fixtures."""
    from functools import partial
    from conftest import load_golden
    from b200swin import swin_transformer_v2 as S
    for name in ["layer_post_convln_c64_ws4_pad", "layer_pre_conv_normal_c64_ws4"]:
        g = load_golden(name)
        dim, nH, ws, _, H, W, B, depth, down, post, shift, Wh, Ww = g["meta.cfg"].tolist()
        at, rct, rot, mt = [str(s) for s in g["meta.types"]]
        layer = S.BasicLayer(dim=dim, depth=depth, num_heads=nH, window_size=ws, norm_layer=partial(S.LayerNormFP32, eps=1e-6),
                             downsample=S.PatchMerging, use_shift=True, init_values=0.5 if not post else None,
                             relative_coords_table_type=rct, rpe_output_type=rot, attn_type=at, mlp_type=mt,
                             postnorm=bool(post), pretrain_window_size=ws)
        assert {k: tuple(v.shape) for k, v in layer.state_dict().items()} == \
            {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
    g = load_golden("wattn_normal_none_c96_h3_ws6")
    C, nH, ws = g["meta.cfg"].tolist()[:3]
    wa = S.WindowAttention(C, (ws, ws), nH, attn_type="normal", relative_coords_table_type="none", rpe_output_type="sigmoid")
    assert {k: tuple(v.shape) for k, v in wa.state_dict().items()} == {k[3:]: g[k].shape for k in g.files if k.startswith("sd.")}
