#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own, so these
files - outputs of the reference modules themselves on seeded inputs - are what pins the
oracle (tests/test_oracle_golden.py) and, on the GPU box, the CUDA path.

Every fixture is an .npz with
  sd.<state_dict key>   parameters and buffers of the reference module
  in.<name>             inputs
  out.<name>            outputs
  grad.<name>           d(sum(out * cot))/d(<name>) for inputs ("in.x" -> "grad.x") and
                        parameters ("grad.sd.<key>"), cot = "in.cot"
  meta.<name>           small integer / float configuration arrays
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_import as R  # noqa: E402


def _np(t):
    return t.detach().cpu().numpy()


def _save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB  ({len(arrs)} arrays)")


def _sd_arrays(mod):
    return {"sd." + k: _np(v) for k, v in mod.state_dict().items()}


def _randomise(mod, gen, norm_gamma=True):
    """Give every parameter an O(1)-informative random value (the reference's init sets block
    norm gammas to 1e-5, which would hide attention bugs - SURVEY.md section 4)."""
    with torch.no_grad():
        for n, p in mod.named_parameters():
            if n.endswith("logit_scale"):
                p.copy_(torch.log(10 * torch.ones_like(p)) + 0.5 * torch.randn(p.shape, generator=gen))
            elif "norm" in n and n.endswith("weight"):
                p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=gen))
            elif n.endswith("bias") or "gamma" in n:
                p.copy_(0.2 * torch.randn(p.shape, generator=gen))
            elif "rpe_mlp" in n:
                p.copy_(0.5 * torch.randn(p.shape, generator=gen))
            elif p.ndim >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=gen) * (1.5 / fan_in ** 0.5))
            else:
                p.copy_(0.2 * torch.randn(p.shape, generator=gen))


def _grads(mod, out, cot, inputs):
    loss = (out * cot).sum()
    params = [(n, p) for n, p in mod.named_parameters() if p.requires_grad]
    gs = torch.autograd.grad(loss, [t for _, t in inputs] + [p for _, p in params], allow_unused=True)
    res = {}
    for (n, _), g in zip(inputs, gs[:len(inputs)]):
        res["grad." + n] = _np(g)
    for (n, p), g in zip(params, gs[len(inputs):]):
        res["grad.sd." + n] = _np(g if g is not None else torch.zeros_like(p))
    return res


# ------------------------------------------------------------------ index maps / buffers
def gen_index_maps():
    S = R.load_swin()
    arrs = {}
    cases = [(2, 7, 7, 4), (1, 30, 30, 12), (2, 15, 20, 8), (1, 11, 38, 12), (1, 8, 8, 4), (1, 6, 6, 6)]
    arrs["meta.cases"] = np.array(cases, dtype=np.int64)
    for ci, (B, H, W, ws) in enumerate(cases):
        s = ws // 2
        Hp, Wp = (H + ws - 1) // ws * ws, (W + ws - 1) // ws * ws
        # tokens numbered from 1 so that the zero pad is distinguishable (-> index -1)
        x = (torch.arange(B * H * W, dtype=torch.float64) + 1).view(B, H, W, 1)
        xp = torch.nn.functional.pad(x, (0, 0, 0, Wp - W, 0, Hp - H))
        part = S.window_partition(xp, ws).view(-1, ws * ws)
        arrs[f"c{ci}.partition"] = (part.long() - 1).numpy()
        rolled = torch.roll(xp, shifts=(-s, -s), dims=(1, 2))
        gat = S.window_partition(rolled, ws).view(-1, ws * ws)
        arrs[f"c{ci}.gather_shift"] = (gat.long() - 1).numpy()
        # output side: number window-major slots, reverse, un-roll, crop
        slots = torch.arange(gat.numel(), dtype=torch.float64).view(-1, ws, ws, 1)
        rev = S.window_reverse(slots, ws, Hp, Wp)
        arrs[f"c{ci}.reverse"] = rev.long().view(B, Hp, Wp).numpy()
        back = torch.roll(rev, shifts=(s, s), dims=(1, 2))[:, :H, :W, :].contiguous()
        arrs[f"c{ci}.scatter_shift"] = back.long().view(B, H, W).numpy()
        # shift mask exactly as BasicLayer.forward builds it (captured through a pre-hook)
        layer = R.quiet(S.BasicLayer, dim=8, depth=2, num_heads=1, window_size=ws, attn_type="cosine_mh",
                        relative_coords_table_type="norm8_log_bylayer", rpe_output_type="sigmoid",
                        pretrain_window_size=ws, norm_layer=S.LayerNormFP32)
        cap = {}

        def _grab(m, a, cap=cap):
            cap["mask"] = a[1]

        layer.blocks[1].register_forward_pre_hook(_grab)
        with torch.no_grad():
            layer(torch.zeros(1, H * W, 8), H, W)
        arrs[f"c{ci}.mask"] = _np(cap["mask"]).astype(np.float32)
    for ws, pre in [(2, 2), (3, 2), (4, 4), (6, 6), (12, 12), (24, 12), (8, 12)]:
        wa = R.quiet(S.WindowAttention, 32, (ws, ws), 1, attn_type="cosine_mh",
                     relative_coords_table_type="norm8_log_bylayer", rpe_output_type="sigmoid",
                     pretrain_window_size=pre)
        arrs[f"rpi.ws{ws}"] = _np(wa.relative_position_index)
        arrs[f"rct.ws{ws}.pre{pre}"] = _np(wa.relative_coords_table)
    _save("index_maps", **arrs)


# ------------------------------------------------------------------ WindowAttention
def gen_window_attention():
    S = R.load_swin()
    for name, (C, nH, ws, pre, B_, nW, seed) in {
        "wattn_c64_h2_ws4_masked": (64, 2, 4, 4, 6, 3, 11),
        "wattn_c96_h3_ws6_pre12": (96, 3, 6, 12, 4, 0, 12),
        "wattn_c128_h4_ws12": (128, 4, 12, 12, 2, 0, 13),
    }.items():
        gen = torch.Generator().manual_seed(seed)
        wa = R.quiet(S.WindowAttention, C, (ws, ws), nH, attn_type="cosine_mh",
                     relative_coords_table_type="norm8_log_bylayer", rpe_output_type="sigmoid",
                     pretrain_window_size=pre)
        _randomise(wa, gen)
        with torch.no_grad():
            wa.logit_scale[0] = 5.0                     # above ln(100)=4.605 -> clamped, zero grad
        N = ws * ws
        x = torch.randn(B_, N, C, generator=gen, requires_grad=True)
        cot = torch.randn(B_, N, C, generator=gen)
        mask = None
        arrs = {}
        if nW:
            mask = torch.where(torch.rand(nW, N, N, generator=gen) < 0.3, -100.0, 0.0)
            arrs["in.mask"] = _np(mask)
        y = wa(x, mask)
        arrs.update(_sd_arrays(wa))
        arrs.update({"in.x": _np(x), "in.cot": _np(cot), "out.y": _np(y),
                     "meta.cfg": np.array([C, nH, ws, pre, B_, nW], dtype=np.int64)})
        arrs.update(_grads(wa, y, cot, [("x", x)]))
        _save(name, **arrs)


# ------------------------------------------------------------------ BasicLayer (blocks + mask + merging)
def gen_basic_layer():
    S = R.load_swin()
    from functools import partial
    cfgs = {
        # name: dim, nH, ws, pre, H, W, B, depth, downsample, postnorm, use_shift, seed
        "layer_post_c64_ws4_pad": (64, 2, 4, 4, 10, 7, 2, 2, True, True, True, 21),
        "layer_post_c32_ws6_nopad": (32, 1, 6, 12, 12, 12, 1, 2, False, True, True, 22),
        "layer_pre_c64_ws4": (64, 2, 4, 4, 8, 9, 2, 2, True, False, True, 23),
        "layer_post_c64_ws4_noshift": (64, 2, 4, 4, 6, 6, 2, 2, False, True, False, 24),
        "layer_post_c128_ws12_pad": (128, 4, 12, 12, 15, 15, 1, 2, True, True, True, 25),
    }
    for name, (dim, nH, ws, pre, H, W, B, depth, down, post, shift, seed) in cfgs.items():
        gen = torch.Generator().manual_seed(seed)
        layer = R.quiet(S.BasicLayer, dim=dim, depth=depth, num_heads=nH, window_size=ws,
                        norm_layer=partial(S.LayerNormFP32, eps=1e-6),
                        downsample=S.PatchMerging if down else None, use_shift=shift,
                        init_values=0.5 if not post else None,
                        relative_coords_table_type="norm8_log_bylayer", rpe_output_type="sigmoid",
                        attn_type="cosine_mh", postnorm=post, pretrain_window_size=pre)
        _randomise(layer, gen)
        layer.eval()
        x = torch.randn(B, H * W, dim, generator=gen, requires_grad=True)
        x_out, H1, W1, x_down, Wh, Ww = layer(x, H, W)
        cot = torch.randn(x_down.shape, generator=gen)
        cot2 = torch.randn(x_out.shape, generator=gen)
        arrs = _sd_arrays(layer)
        arrs.update({"in.x": _np(x), "in.cot": _np(cot), "in.cot2": _np(cot2),
                     "out.x": _np(x_out), "out.x_down": _np(x_down),
                     "meta.cfg": np.array([dim, nH, ws, pre, H, W, B, depth, int(down), int(post), int(shift), Wh, Ww],
                                          dtype=np.int64)})
        if down:
            total = (x_down * cot).sum() + (x_out * cot2).sum()
        else:
            total = (x_out * cot2).sum()
        params = [(n, p) for n, p in layer.named_parameters()]
        gs = torch.autograd.grad(total, [x] + [p for _, p in params], allow_unused=True)
        arrs["grad.x"] = _np(gs[0])
        for (n, p), g in zip(params, gs[1:]):
            arrs["grad.sd." + n] = _np(g if g is not None else torch.zeros_like(p))
        _save(name, **arrs)


# ------------------------------------------------------------------ whole encoder (small)
def gen_swin_small():
    S = R.load_swin()
    gen = torch.Generator().manual_seed(31)
    # three stages (the reference's MODEL_SCALE 16 call, models/model.py:57-67) keep the file small
    cfg = dict(embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4], window_size=[4, 4, 2],
               pretrain_window_size=[4, 4, 2], drop_path_rate=0.0, use_checkpoint=False,
               use_shift=[True, True, False], out_indices=(0, 2))
    net = R.quiet(S.SwinTransformerV2, **cfg)
    R.quiet(net.init_weights, None)
    _randomise(net, gen)
    S.SwinTransformerV2.train(net, False)
    img = torch.rand(2, 3, 72, 56, generator=gen, requires_grad=True)
    outs = net(img)
    cots = [torch.randn(o.shape, generator=gen) for o in outs]
    total = sum((o * c).sum() for o, c in zip(outs, cots))
    params = [(n, p) for n, p in net.named_parameters()]
    gs = torch.autograd.grad(total, [img] + [p for _, p in params], allow_unused=True)
    arrs = _sd_arrays(net)
    arrs.update({"in.img": _np(img), "grad.img": _np(gs[0])})
    for i, (o, c) in enumerate(zip(outs, cots)):
        arrs[f"out.{i}"] = _np(o)
        arrs[f"in.cot{i}"] = _np(c)
    # full gradients for a subset, (sum, L2) checksums for every parameter
    keep = ("patch_embed", "layers.0.", "layers.1.blocks.1.", "norm0", "norm2", "layers.1.downsample")
    names, sums = [], []
    for (n, p), g in zip(params, gs[1:]):
        g = g if g is not None else torch.zeros_like(p)
        if n.startswith(keep):
            arrs["grad.sd." + n] = _np(g)
        names.append(n)
        sums.append([g.double().sum().item(), g.double().pow(2).sum().sqrt().item()])
    arrs["gradsum.names"] = np.array(names)
    arrs["gradsum.values"] = np.array(sums, dtype=np.float64)
    arrs["meta.cfg_json"] = np.array(__import__("json").dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}))
    _save("swin_small", **arrs)


# ------------------------------------------------------------------ SiLog + metrics
def gen_silog():
    C = R.load_criterion()
    M = R.load_metrics()
    crit = C.SiLogLoss()
    arrs = {}
    cases = {}
    cases["kat"] = (torch.tensor([1.0, 2.0, 4.0, 3.0]), torch.tensor([1.0, 1.0, 0.0, 6.0]))
    gen = torch.Generator().manual_seed(41)
    for nm, (shape, maxd, inval) in {"nyu": ((2, 48, 64), 10.0, 0.05), "kitti": ((1, 22, 76), 80.0, 0.9),
                                     "void": ((3, 30, 40), 10.0, 0.3), "allvalid": ((1, 16, 16), 10.0, 0.0)}.items():
        tgt = 0.5 + (maxd - 0.5) * torch.rand(shape, generator=gen)
        tgt = torch.where(torch.rand(shape, generator=gen) < inval, torch.zeros(()), tgt)
        pred = 0.2 + maxd * torch.rand(shape, generator=gen)
        cases[nm] = (pred, tgt)
    for nm, (pred, tgt) in cases.items():
        p = pred.clone().requires_grad_(True)
        loss = crit(p, tgt)
        (g,) = torch.autograd.grad(loss * 1.7, p)
        arrs[f"{nm}.pred"] = _np(pred)
        arrs[f"{nm}.target"] = _np(tgt)
        arrs[f"{nm}.loss"] = _np(loss).reshape(1)
        arrs[f"{nm}.grad_x1p7"] = _np(g)
    allinv = crit(torch.ones(4), torch.zeros(4))
    arrs["allinvalid.loss"] = _np(allinv).reshape(1)
    # eval_depth known-answer + random
    m = M.eval_depth(torch.tensor([1.0, 2.0, 3.0]), torch.tensor([1.0, 1.0, 6.0]))
    arrs["metrics.kat"] = np.array([m[k] for k in sorted(m)], dtype=np.float64)
    p = 0.5 + 9 * torch.rand(4000, generator=gen)
    t = 0.5 + 9 * torch.rand(4000, generator=gen)
    m = M.eval_depth(p, t)
    arrs["metrics.rand.pred"], arrs["metrics.rand.target"] = _np(p), _np(t)
    arrs["metrics.rand"] = np.array([m[k] for k in sorted(m)], dtype=np.float64)
    arrs["metrics.keys"] = np.array(sorted(m))
    _save("silog", **arrs)


# ------------------------------------------------------------------ SURVEY 8f-4 variants
def gen_variants():
    """attn_type='normal', the learned bias table (relative_coords_table_type='none'), ConvMlp (mlp_type 'conv' /
    'conv_ln'): reachable through non-default constructor arguments of the reference file."""
    S = R.load_swin()
    from functools import partial
    # WindowAttention, plain scaled dot product
    for name, (C, nH, ws, B_, nW, rct, rot, seed) in {
        "wattn_normal_c64_h2_ws4_masked": (64, 2, 4, 6, 3, "norm8_log", "normal", 51),
        "wattn_normal_none_c96_h3_ws6": (96, 3, 6, 4, 0, "none", "sigmoid", 52),
    }.items():
        gen = torch.Generator().manual_seed(seed)
        wa = R.quiet(S.WindowAttention, C, (ws, ws), nH, attn_type="normal", relative_coords_table_type=rct,
                     rpe_output_type=rot)
        _randomise(wa, gen)
        if rct == "none":
            with torch.no_grad():
                wa.relative_position_bias_table.copy_(torch.randn(wa.relative_position_bias_table.shape, generator=gen))
        N = ws * ws
        x = torch.randn(B_, N, C, generator=gen, requires_grad=True)
        cot = torch.randn(B_, N, C, generator=gen)
        mask = None
        arrs = {}
        if nW:
            mask = torch.where(torch.rand(nW, N, N, generator=gen) < 0.3, -100.0, 0.0)
            arrs["in.mask"] = _np(mask)
        y = wa(x, mask)
        arrs.update(_sd_arrays(wa))
        arrs.update({"in.x": _np(x), "in.cot": _np(cot), "out.y": _np(y),
                     "meta.cfg": np.array([C, nH, ws, ws, B_, nW], dtype=np.int64),
                     "meta.types": np.array(["normal", rct, rot])})
        arrs.update(_grads(wa, y, cot, [("x", x)]))
        _save(name, **arrs)
    # BasicLayer with ConvMlp
    cfgs = {
        # name: dim, nH, ws, H, W, B, postnorm, attn_type, mlp_type, rct, rot, seed
        "layer_post_convln_c64_ws4_pad": (64, 2, 4, 10, 7, 2, True, "cosine_mh", "conv_ln", "norm8_log_bylayer", "sigmoid", 53),
        "layer_pre_conv_normal_c64_ws4": (64, 2, 4, 8, 9, 2, False, "normal", "conv", "norm8_log", "normal", 54),
    }
    for name, (dim, nH, ws, H, W, B, post, at, mt, rct, rot, seed) in cfgs.items():
        gen = torch.Generator().manual_seed(seed)
        layer = R.quiet(S.BasicLayer, dim=dim, depth=2, num_heads=nH, window_size=ws,
                        norm_layer=partial(S.LayerNormFP32, eps=1e-6), downsample=S.PatchMerging, use_shift=True,
                        init_values=0.5 if not post else None, relative_coords_table_type=rct, rpe_output_type=rot,
                        attn_type=at, mlp_type=mt, postnorm=post, pretrain_window_size=ws)
        _randomise(layer, gen)
        layer.eval()
        x = torch.randn(B, H * W, dim, generator=gen, requires_grad=True)
        x_out, H1, W1, x_down, Wh, Ww = layer(x, H, W)
        cot = torch.randn(x_down.shape, generator=gen)
        cot2 = torch.randn(x_out.shape, generator=gen)
        arrs = _sd_arrays(layer)
        arrs.update({"in.x": _np(x), "in.cot": _np(cot), "in.cot2": _np(cot2), "out.x": _np(x_out),
                     "out.x_down": _np(x_down),
                     "meta.cfg": np.array([dim, nH, ws, ws, H, W, B, 2, 1, int(post), 1, Wh, Ww], dtype=np.int64),
                     "meta.types": np.array([at, rct, rot, mt])})
        total = (x_down * cot).sum() + (x_out * cot2).sum()
        params = [(n, p) for n, p in layer.named_parameters()]
        gs = torch.autograd.grad(total, [x] + [p for _, p in params], allow_unused=True)
        arrs["grad.x"] = _np(gs[0])
        for (n, p), g in zip(params, gs[1:]):
            arrs["grad.sd." + n] = _np(g if g is not None else torch.zeros_like(p))
        _save(name, **arrs)


def gen_transformer_encoder():
    """The reference's global-attention encoder layer (models/cnn_transformer.py:176-216) on a 10 x 13 token map:
    hidden 256 = 4 heads of 64 (the 512 = 8 x 64 configuration shares the code path; it is compared with
    torch.nn.MultiheadAttention directly in the GPU tests, the fixture would be 13 MB)."""
    import types
    T = R.load_cnn_transformer()
    B, N, E, FF = 2, 130, 256, 128
    args = types.SimpleNamespace(transformer_ff_dim=FF)
    # ReLU makes the gradient discontinuous where a pre-activation crosses zero: a single element of ffn1's output whose
    # sign an implementation's rounding flips costs 1/sqrt(B*N*FF) = 5e-3 of relative gradient error.  The fixture is
    # therefore drawn (seed search, deterministic) such that no pre-activation lies within 2e-4 of zero -- 50x the
    # fp32 rounding error of the row sums -- so that the 1e-4 parity bar measures arithmetic, not sign luck.
    seed = 61
    while True:
        gen = torch.Generator().manual_seed(seed)
        enc = T.Transformer_Encoder(args, hidden_dim=E)
        with torch.no_grad():
            for n, p in enc.named_parameters():
                if "norm" in n and n.endswith("weight"):
                    p.copy_(1.0 + 0.3 * torch.randn(p.shape, generator=gen))
                elif n.endswith("bias"):
                    p.copy_(0.2 * torch.randn(p.shape, generator=gen))
                else:
                    p.copy_(torch.randn(p.shape, generator=gen) * (1.5 / p.shape[-1] ** 0.5))
        feat = torch.randn(B, N, E, generator=gen, requires_grad=True)
        pos = torch.randn(B, N, E, generator=gen, requires_grad=True)
        cot = torch.randn(B, N, E, generator=gen)
        pre = {}
        hook = enc.ffn1[0].register_forward_hook(lambda m, i, o: pre.__setitem__("h", o.detach()))
        y = enc(feat, pos)
        hook.remove()
        margin = pre["h"].abs().min().item()
        if margin > 2e-4:
            break
        seed += 1
    print(f"tenc: seed {seed}, smallest |ffn1 pre-activation| = {margin:.2e}")
    qk = feat + pos
    _, w = enc.self_attn(qk, qk, feat)                       # the weights the reference receives at :201
    arrs = _sd_arrays(enc)
    arrs.update({"in.feat": _np(feat), "in.pos": _np(pos), "in.cot": _np(cot), "out.y": _np(y), "out.weights": _np(w),
                 "meta.cfg": np.array([B, N, E, 4, FF], dtype=np.int64), "meta.seed": np.array([seed], dtype=np.int64)})
    arrs.update(_grads(enc, y, cot, [("feat", feat), ("pos", pos)]))
    _save("tenc_h256_n130", **arrs)


if __name__ == "__main__":
    assert R.available(), "reference not mounted at " + R.REF_ROOT
    torch.manual_seed(0)
    torch.set_num_threads(4)
    only = sys.argv[1:]                 # e.g. `make_golden.py variants tenc` regenerates just those groups
    groups = {"index_maps": gen_index_maps, "wattn": gen_window_attention, "layers": gen_basic_layer,
              "swin_small": gen_swin_small, "silog": gen_silog, "variants": gen_variants,
              "tenc": gen_transformer_encoder}
    for k, fn in groups.items():
        if not only or k in only:
            fn()
