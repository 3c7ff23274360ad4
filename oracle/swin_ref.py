"""Functional torch-CPU restatement of the Swin-V2 hot path (fp32 or fp64).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Pure functions over a reference
``state_dict`` (same key names as the reference modules), differentiable through torch
autograd so gradients can be compared, plus the hand-derived attention backward of
SURVEY.md appendix A (``window_attention_backward``) that the CUDA backward follows.
Each function cites the reference lines it restates (relative to ``/root/reference``).

Nothing here copies tensors for layout's sake: roll / partition / reverse are applied as
index gathers built by ``oracle.index_maps``.
"""
from __future__ import annotations

import math
from typing import Mapping, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import index_maps as im

LOGIT_SCALE_MAX = math.log(1.0 / 0.01)       # swin_transformer_v2.py:294
MASK_VALUE = -100.0                          # swin_transformer_v2.py:892


def _sub(sd: Mapping[str, torch.Tensor], prefix: str) -> dict:
    p = prefix if (prefix == "" or prefix.endswith(".")) else prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


# ----------------------------------------------------------------------------- pieces
def layer_norm_fp32(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """LayerNormFP32.forward, swin_transformer_v2.py:41-47 (compute in >= fp32, return
    the input dtype).  eps=1e-6 is what SwinTransformerV2 instantiates (:1038)."""
    cd = torch.float64 if x.dtype == torch.float64 else torch.float32
    y = F.layer_norm(x.to(cd), (x.shape[-1],), w.to(cd), b.to(cd), eps)
    return y.to(x.dtype)


def gelu_exact(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() default (erf form), used by Mlp (swin_transformer_v2.py:60,80)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def mlp(x: torch.Tensor, sd: Mapping[str, torch.Tensor]) -> torch.Tensor:
    """Mlp.forward with norm=None, drop=0, mlpfp32=False (swin_transformer_v2.py:76-89)."""
    h = F.linear(x, sd["fc1.weight"], sd["fc1.bias"])
    h = gelu_exact(h)
    return F.linear(h, sd["fc2.weight"], sd["fc2.bias"])


def conv_mlp(x: torch.Tensor, sd: Mapping[str, torch.Tensor], H: int, W: int) -> torch.Tensor:
    """ConvMlp.forward (swin_transformer_v2.py:107-117): depthwise 3x3 conv on the [B,C,H,W] view, optional
    LayerNorm2D(LayerNormFP32, default eps 1e-5) when the state_dict holds proj_ln, then the Mlp."""
    B, L, C = x.shape
    y = x.view(B, H, W, C).permute(0, 3, 1, 2)
    y = F.conv2d(y, sd["conv_proj.weight"], None, stride=1, padding=1, groups=C)      # :98-104, :110
    y = y.permute(0, 2, 3, 1)
    if "proj_ln.ln.weight" in sd:                                                     # :112-113
        y = layer_norm_fp32(y, sd["proj_ln.ln.weight"], sd["proj_ln.ln.bias"], 1e-5)
    return mlp(y.reshape(B, L, C), _sub(sd, "mlp"))


def any_mlp(x: torch.Tensor, sd: Mapping[str, torch.Tensor], H: int, W: int) -> torch.Tensor:
    """mlp_type 'normal' or 'conv' / 'conv_ln' (swin_transformer_v2.py:401-409), told apart by the state_dict keys."""
    return conv_mlp(x, sd, H, W) if "conv_proj.weight" in sd else mlp(x, sd)


def cpb_table(sd: Mapping[str, torch.Tensor]) -> torch.Tensor:
    """rpe_mlp(relative_coords_table) -> [(2ws-1)^2, nH]  (swin_transformer_v2.py:185-187, 304).
    Linear(2,512,bias) -> ReLU -> Linear(512,nH,no bias, fp32)."""
    t = sd["relative_coords_table"]
    h = F.relu(F.linear(t, sd["rpe_mlp.0.weight"], sd["rpe_mlp.0.bias"]))
    out = F.linear(h, sd["rpe_mlp.2.weight"])
    return out.reshape(-1, out.shape[-1])


def cpb_bias(sd: Mapping[str, torch.Tensor], N: int, rpe_output_type: str = "sigmoid") -> torch.Tensor:
    """bias[h, i, j] = 16 * sigmoid(table[relative_position_index[i, j], h])
    (swin_transformer_v2.py:302-313) -> [nH, N, N]; the table is the learned relative_position_bias_table when the
    module was built with relative_coords_table_type='none' (:241-244, :305-306); rpe_output_type 'normal' skips the
    sigmoid (:309-310)."""
    table = sd["relative_position_bias_table"] if "relative_position_bias_table" in sd else cpb_table(sd)
    idx = sd["relative_position_index"].reshape(-1).long()
    b = table[idx].reshape(N, N, -1).permute(2, 0, 1)
    return 16.0 * torch.sigmoid(b) if rpe_output_type == "sigmoid" else b


def logit_scale_eff(logit_scale: torch.Tensor) -> torch.Tensor:
    """exp(min(logit_scale, ln 100)), swin_transformer_v2.py:294."""
    return torch.clamp(logit_scale, max=LOGIT_SCALE_MAX).exp()


def window_attention(x: torch.Tensor, sd: Mapping[str, torch.Tensor], num_heads: int,
                     mask: torch.Tensor | None = None, return_aux: bool = False, rpe_output_type: str = "sigmoid",
                     qk_scale: float | None = None):
    """WindowAttention.forward with qkv_bias=True (swin_transformer_v2.py:275-336).  attn_type is 'cosine_mh' when the
    state_dict holds logit_scale, else 'normal' (:296-298: q * scale . k with scale = qk_scale or head_dim^-0.5).
    x: [B_, N, C]; mask: [nW, N, N] or None."""
    B_, N, C = x.shape
    hd = C // num_heads
    qkv_bias = torch.cat((sd["q_bias"], torch.zeros_like(sd["v_bias"]), sd["v_bias"]))   # :283-285
    qkv = F.linear(x, sd["qkv.weight"], qkv_bias)                                        # :286
    qkv = qkv.reshape(B_, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)                    # :287
    q, k, v = qkv[0], qkv[1], qkv[2]
    if "logit_scale" in sd:
        if q.dtype in (torch.bfloat16, torch.float16):       # under autocast the reference widens first (.float(), :292-293)
            q, k = q.float(), k.float()
        qn = F.normalize(q, dim=-1)                                                      # :292 (eps 1e-12)
        kn = F.normalize(k, dim=-1)                                                      # :293
        scale = logit_scale_eff(sd["logit_scale"])                                       # :294
        cos = qn @ kn.transpose(-2, -1)
        attn = cos * scale                                                               # :295
    else:
        qn, kn = q, k
        scale = torch.tensor(qk_scale or hd ** -0.5, dtype=x.dtype)                      # :178-180
        cos = q @ k.transpose(-2, -1)
        attn = (q * scale) @ k.transpose(-2, -1)                                         # :297-298
    bias = cpb_bias(sd, N, rpe_output_type)
    attn = attn + bias.unsqueeze(0)                                                      # :317
    if mask is not None:                                                                 # :319-322
        nW = mask.shape[0]
        attn = attn.view(B_ // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    p = torch.softmax(attn, dim=-1)                                                      # :324
    if p.dtype != x.dtype and x.dtype in (torch.bfloat16, torch.float16):
        p = p.type_as(x)                                                                 # :325
    o = (p @ v).transpose(1, 2).reshape(B_, N, C)                                        # :328
    y = F.linear(o, sd["proj.weight"], sd["proj.bias"])                                  # :334
    if return_aux:
        return y, dict(q=q, k=k, v=v, qn=qn, kn=kn, cos=cos, p=p, o=o, bias=bias, scale=scale)
    return y


def window_attention_backward(g: torch.Tensor, x: torch.Tensor, sd: Mapping[str, torch.Tensor],
                              num_heads: int, mask: torch.Tensor | None = None) -> dict:
    """Hand-derived backward of ``window_attention``; returns every
    intermediate gradient the CUDA backward produces so each can be checked separately."""
    B_, N, C = x.shape
    hd = C // num_heads
    _, a = window_attention(x, sd, num_heads, mask, return_aux=True)
    q, k, v, qn, kn, cos, p, o, bias, scale = (a[n] for n in
                                               ("q", "k", "v", "qn", "kn", "cos", "p", "o", "bias", "scale"))
    g2 = g.reshape(-1, C)
    d_proj_w = g2.t() @ o.reshape(-1, C)
    d_proj_b = g2.sum(0)
    dO = (g2 @ sd["proj.weight"]).reshape(B_, N, num_heads, hd).transpose(1, 2)
    dV = p.transpose(-2, -1) @ dO
    dP = dO @ v.transpose(-2, -1)
    dS = p * (dP - (dP * p).sum(-1, keepdim=True))
    d_bias = dS.sum(0)                                         # [nH, N, N]
    d_scale = (dS * cos).sum(dim=(0, 2, 3))                    # [nH]
    ls = sd["logit_scale"].reshape(-1)
    d_logit_scale = d_scale * scale.reshape(-1) * (ls <= LOGIT_SCALE_MAX).to(g.dtype)
    dqn = (dS @ kn) * scale
    dkn = (dS.transpose(-2, -1) @ qn) * scale
    nq = q.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    nk = k.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    dq = (dqn - qn * (dqn * qn).sum(-1, keepdim=True)) / nq
    dk = (dkn - kn * (dkn * kn).sum(-1, keepdim=True)) / nk
    dqkv = torch.stack((dq, dk, dV), 0).permute(1, 3, 0, 2, 4).reshape(B_ * N, 3 * C)
    dx = (dqkv @ sd["qkv.weight"]).reshape(B_, N, C)
    d_qkv_w = dqkv.t() @ x.reshape(-1, C)
    d_q_bias = dqkv[:, :C].sum(0)
    d_v_bias = dqkv[:, 2 * C:].sum(0)
    # 16*sigmoid' = bias*(1-bias/16); scatter-add over relative_position_index
    idx = sd["relative_position_index"].reshape(-1).long()
    d_pre = (d_bias * bias * (1.0 - bias / 16.0)).permute(1, 2, 0).reshape(N * N, num_heads)
    n_tab = sd["relative_coords_table"].numel() // 2
    d_table = torch.zeros(n_tab, num_heads, dtype=g.dtype).index_add_(0, idx, d_pre)
    # 16*sigmoid(table) table-space gradient (what the kernels accumulate): sum of dS per rel index
    d_t16 = torch.zeros(n_tab, num_heads, dtype=g.dtype).index_add_(
        0, idx, d_bias.permute(1, 2, 0).reshape(N * N, num_heads))
    return dict(dx=dx, d_qkv_w=d_qkv_w, d_q_bias=d_q_bias, d_v_bias=d_v_bias, d_proj_w=d_proj_w,
                d_proj_b=d_proj_b, d_bias=d_bias, d_scale=d_scale, d_logit_scale=d_logit_scale,
                d_table=d_table, d_t16=d_t16, dqkv=dqkv, dS=dS)


# ------------------------------------------------------------------- window plumbing
def gather_windows(x: torch.Tensor, H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """pad + roll(-shift) + window_partition as one gather.  x: [B, H*W, C] -> [B*nW, N, C]
    (swin_transformer_v2.py:426-446)."""
    B, L, C = x.shape
    idx = torch.from_numpy(im.fused_gather_index(B, H, W, ws, shift))
    flat = torch.cat([x.reshape(B * L, C), x.new_zeros(1, C)], 0)      # row -1 -> zero pad token
    return flat[idx.reshape(-1)].reshape(idx.shape[0], ws * ws, C)


def scatter_windows(wins: torch.Tensor, B: int, H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """window_reverse + roll(+shift) + crop as one gather.  [B*nW, N, C] -> [B, H*W, C]
    (swin_transformer_v2.py:453-465)."""
    C = wins.shape[-1]
    idx = torch.from_numpy(im.fused_scatter_index(B, H, W, ws, shift)).reshape(-1)
    return wins.reshape(-1, C)[idx].reshape(B, H * W, C)


def shift_mask(H: int, W: int, ws: int, shift: int, dtype=torch.float32) -> torch.Tensor:
    """BasicLayer's attn_mask (swin_transformer_v2.py:874-892)."""
    return torch.from_numpy(im.shift_attn_mask(H, W, ws, shift)).to(dtype)


# --------------------------------------------------------------------------- blocks
def block_post(x: torch.Tensor, sd: Mapping[str, torch.Tensor], H: int, W: int, num_heads: int,
               ws: int, shift: int, eps: float = 1e-6,
               drop_scale1: torch.Tensor | None = None, drop_scale2: torch.Tensor | None = None,
               rpe_output_type: str = "sigmoid"):
    """SwinTransformerBlockPost.forward (swin_transformer_v2.py:419-488), post-norm:
    x = sc + DropPath(LN(attn(x)));  x = x + DropPath(LN(mlp(x))).
    ``drop_scale*`` are optional per-sample DropPath multipliers [B] (mask/keep_prob)."""
    B, L, C = x.shape
    assert L == H * W
    shortcut = x
    xw = gather_windows(x, H, W, ws, shift)
    mask = shift_mask(H, W, ws, shift, x.dtype).to(x.device) if shift > 0 else None              # :437-442
    aw = window_attention(xw, _sub(sd, "attn"), num_heads, mask, rpe_output_type=rpe_output_type)
    a = scatter_windows(aw, B, H, W, ws, shift)
    a = layer_norm_fp32(a, sd["norm1.weight"], sd["norm1.bias"], eps)                # :472
    if drop_scale1 is not None:
        a = a * drop_scale1.view(B, 1, 1)
    x = shortcut + a                                                                 # :473
    m = any_mlp(x, _sub(sd, "mlp"), H, W)                                            # :477
    m = layer_norm_fp32(m, sd["norm2.weight"], sd["norm2.bias"], eps)                # :482
    if drop_scale2 is not None:
        m = m * drop_scale2.view(B, 1, 1)
    return x + m                                                                     # :483


def block_pre(x: torch.Tensor, sd: Mapping[str, torch.Tensor], H: int, W: int, num_heads: int,
              ws: int, shift: int, eps: float = 1e-6, rpe_output_type: str = "sigmoid"):
    """SwinTransformerBlockPre.forward (swin_transformer_v2.py:561-630), pre-norm with
    optional gamma_1/gamma_2 (scalars 1.0 when init_values is None)."""
    B, L, C = x.shape
    shortcut = x
    y = layer_norm_fp32(x, sd["norm1.weight"], sd["norm1.bias"], eps)                # :567
    xw = gather_windows(y, H, W, ws, shift)
    mask = shift_mask(H, W, ws, shift, x.dtype).to(x.device) if shift > 0 else None
    aw = window_attention(xw, _sub(sd, "attn"), num_heads, mask, rpe_output_type=rpe_output_type)
    a = scatter_windows(aw, B, H, W, ws, shift)
    g1 = sd.get("gamma_1", 1.0)
    g2 = sd.get("gamma_2", 1.0)
    x = shortcut + g1 * a                                                            # :614-615
    m = any_mlp(layer_norm_fp32(x, sd["norm2.weight"], sd["norm2.bias"], eps), _sub(sd, "mlp"), H, W)
    return x + g2 * m                                                                # :624-625


def patch_merging(x: torch.Tensor, sd: Mapping[str, torch.Tensor], H: int, W: int, eps: float = 1e-6,
                  postnorm: bool = True):
    """PatchMerging.forward (swin_transformer_v2.py:648-678): 2x2 strided gather in
    (0,0),(1,0),(0,1),(1,1) order, then Linear(4C,2C,no bias) -> LN(2C) (postnorm) or
    LN(4C) -> Linear (pre-norm)."""
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if H % 2 == 1 or W % 2 == 1:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    x = x.reshape(B, -1, 4 * C)
    if postnorm:
        x = F.linear(x, sd["reduction.weight"])
        return layer_norm_fp32(x, sd["norm.weight"], sd["norm.bias"], eps)
    x = layer_norm_fp32(x, sd["norm.weight"], sd["norm.bias"], eps)
    return F.linear(x, sd["reduction.weight"])


def basic_layer(x: torch.Tensor, sd: Mapping[str, torch.Tensor], H: int, W: int, depth: int,
                num_heads: int, ws: int, use_shift: bool = True, downsample: bool = True,
                postnorm: bool = True, eps: float = 1e-6, rpe_output_type: str = "sigmoid"):
    """BasicLayer.forward (swin_transformer_v2.py:866-908) -> (x, H, W, x_down, Wh, Ww)."""
    shift = ws // 2
    for i in range(depth):
        s = 0 if (i % 2 == 0 or not use_shift) else shift                          # :814
        fn = block_post if postnorm else block_pre
        x = fn(x, _sub(sd, f"blocks.{i}"), H, W, num_heads, ws, s, eps, rpe_output_type=rpe_output_type)
    if downsample:
        xd = patch_merging(x, _sub(sd, "downsample"), H, W, eps, postnorm)
        return x, H, W, xd, (H + 1) // 2, (W + 1) // 2
    return x, H, W, x, H, W


def patch_embed(img: torch.Tensor, sd: Mapping[str, torch.Tensor], patch: int = 4, eps: float = 1e-6):
    """PatchEmbed.forward with patch_norm=True (swin_transformer_v2.py:941-957) -> [B, C, Wh, Ww]."""
    _, _, H, W = img.shape
    if W % patch != 0:
        img = F.pad(img, (0, patch - W % patch))
    if H % patch != 0:
        img = F.pad(img, (0, 0, 0, patch - H % patch))
    x = F.conv2d(img, sd["proj.weight"], sd["proj.bias"], stride=patch)
    B, C, Wh, Ww = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = layer_norm_fp32(x, sd["norm.weight"], sd["norm.bias"], eps)
    return x.transpose(1, 2).reshape(B, C, Wh, Ww)


def swin_v2(img: torch.Tensor, sd: Mapping[str, torch.Tensor], embed_dim: int, depths: Sequence[int],
            num_heads: Sequence[int], window_size: Sequence[int], use_shift: Sequence[bool],
            out_indices: Sequence[int] = (3,), eps: float = 1e-6) -> list:
    """SwinTransformerV2.forward (swin_transformer_v2.py:1251-1277), ape=False, eval mode."""
    x = patch_embed(img, _sub(sd, "patch_embed"), 4, eps)
    Wh, Ww = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    outs = []
    n = len(depths)
    for i in range(n):
        x_out, H, W, x, Wh, Ww = basic_layer(x, _sub(sd, f"layers.{i}"), Wh, Ww, depths[i], num_heads[i],
                                             window_size[i], use_shift[i], downsample=(i < n - 1), eps=eps)
        if i in out_indices:
            C = embed_dim * 2 ** i
            y = layer_norm_fp32(x_out.float() if x_out.dtype != torch.float64 else x_out,
                                sd[f"norm{i}.weight"], sd[f"norm{i}.bias"], eps)
            outs.append(y.view(-1, H, W, C).permute(0, 3, 1, 2).contiguous())
    return outs


def to_dtype(sd: Mapping[str, torch.Tensor], dtype) -> dict:
    """Cast the floating tensors of a state_dict (keeps integer buffers)."""
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def npz_to_sd(npz, prefix: str = "sd.") -> dict:
    """Collect ``sd.<key>`` arrays of a golden .npz file into a torch state_dict."""
    return {k[len(prefix):]: torch.from_numpy(np.asarray(npz[k])) for k in npz.files if k.startswith(prefix)}
