"""Drop-in for the reference's ``models/swin_transformer_v2.py`` on B200.

Same class names, constructor keywords, ``forward`` signatures and ``state_dict`` keys as the
reference for the configuration ``SwinTransformerV2`` instantiates (``attn_type='cosine_mh'``,
``relative_coords_table_type='norm8_log_bylayer'``, ``rpe_output_type='sigmoid'``, post-norm, plain
``Mlp``, ``PatchMerging``), so ``models/model.py`` / ``models/optimizer.py`` / checkpoints work
unchanged.  Every tensor op on the path runs in hand-written sm_100a kernels behind the C-ABI
(``include/b200swin.h``); there is no eager fallback.  How the block executes:

    x [B, H*W, C]  (natural token order, never permuted in HBM)
      -> qkv GEMM (tcgen05) with q_bias/v_bias + per-head L2-normalisation in the epilogue
      -> attention core: pad / cyclic roll / window partition / shift mask / CPB bias / softmax / PV /
         window reverse / roll back / crop are all address math + on-chip work in ONE kernel
      -> proj GEMM (tcgen05, bias epilogue)
      -> LayerNorm + DropPath scale + residual add (one kernel)
      -> fc1 GEMM (+GELU epilogue) -> fc2 GEMM -> LayerNorm + DropPath + residual (one kernel)

Variants of the reference file that no shipped config sets but SURVEY.md section 8f-4 lists are built on the same
kernels: ``attn_type='normal'`` (plain scaled dot product, :296-298), ``relative_coords_table_type='none'`` (learned
bias table, :241-244), ``mlp_type='conv' / 'conv_ln'`` (ConvMlp, :92-117, depthwise conv on the token layout) and the
pre-norm block.  ConvPatchMerging, ResNetDLNPatchEmbed, ape, endnorm, mlpfp32 and strid16 raise NotImplementedError.
"""
from __future__ import annotations

import math
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint as checkpoint

from . import ops

_LOGIT_MAX = math.log(1.0 / 0.01)


def to_2tuple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std)


class DropPath(nn.Module):
    """Per-sample stochastic depth (timm semantics: Bernoulli(keep) / keep).  The mask is drawn with the
    torch RNG; the multiply itself is fused into the LayerNorm+residual kernel via ``sample_scale``."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self._drawn = []          # masks pre-drawn for this forward by SwinTransformerV2._draw_drop_paths

    def sample_scale(self, x: torch.Tensor):
        if self.drop_prob == 0.0 or not self.training:
            return None
        if self._drawn and self._drawn[0].shape[0] == x.shape[0]:
            return self._drawn.pop(0)
        keep = 1.0 - self.drop_prob
        m = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device).bernoulli_(keep)
        return m.div_(keep) if keep > 0.0 else m

    def forward(self, x):
        s = self.sample_scale(x)
        return x if s is None else x * s.view(-1, *([1] * (x.ndim - 1))).to(x.dtype)

    def extra_repr(self):
        return f"drop_prob={self.drop_prob}"


class LayerNormFP32(nn.LayerNorm):
    """LayerNorm computed in fp32 whatever the storage type (reference :41-47)."""

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return ops.layer_norm_residual(input, self.weight, self.bias, self.eps)


class LinearFP32(nn.Linear):
    """fp32 Linear used as the last layer of rpe_mlp (reference :50-56).  [(2ws-1)^2, 512] x [512, nH]:
    a few kFLOP that stay in PyTorch so autograd produces the rpe_mlp gradients."""

    def forward(self, input: torch.Tensor) -> torch.Tensor:
        return F.linear(input.float(), self.weight.float(), self.bias.float() if self.bias is not None else None)


class Mlp(nn.Module):
    """fc1 -> GELU -> fc2 (reference :59-89), GELU fused into the fc1 GEMM epilogue."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.,
                 norm_layer=None, mlpfp32=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if norm_layer is not None or mlpfp32 or drop != 0. or act_layer is not nn.GELU:
            raise NotImplementedError("b200swin.Mlp: only the default GELU MLP (no norm, no dropout) is built")
        self.mlpfp32 = mlpfp32
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self.norm = None

    def forward(self, x, H=None, W=None, fc2_bias_grad_elsewhere=False, passthrough=False):
        # fc2_bias_grad_elsewhere: the caller's LayerNorm backward returns fc2.bias' gradient (ops.layer_norm_residual)
        # passthrough: also return an alias of x for the caller's residual branch (ops._Mlp.forward)
        b2 = self.fc2.bias.detach() if (fc2_bias_grad_elsewhere and self.fc2.bias is not None) else self.fc2.bias
        return ops.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, b2, passthrough)


class LayerNorm2D(nn.Module):
    """LayerNorm over the channels of an NCHW tensor (reference :26-38)."""

    def __init__(self, normalized_shape, norm_layer=None):
        super().__init__()
        self.ln = norm_layer(normalized_shape) if norm_layer is not None else nn.Identity()

    def forward(self, x):
        return self.ln(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class ConvMlp(nn.Module):
    """Depthwise 3x3 conv (+ optional LayerNorm) in front of the MLP (reference :92-117; mlp_type 'conv' / 'conv_ln').
    The conv runs on the token layout (csrc/dwconv.cu): no NHWC <-> NCHW permutes."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.,
                 norm_layer=None, mlpfp32=False, proj_ln=False):
        super().__init__()
        self.mlp = Mlp(in_features=in_features, hidden_features=hidden_features, out_features=out_features,
                       act_layer=act_layer, drop=drop, norm_layer=norm_layer, mlpfp32=mlpfp32)
        self.conv_proj = nn.Conv2d(in_features, in_features, kernel_size=3, padding=1, stride=1, bias=False,
                                   groups=in_features)
        self.proj_ln = LayerNorm2D(in_features, LayerNormFP32) if proj_ln else None

    def forward(self, x, H, W, fc2_bias_grad_elsewhere=False, passthrough=False):
        B, L, C = x.shape
        assert L == H * W
        y = ops.dwconv3x3(x.view(B, H, W, C), self.conv_proj.weight).view(B, L, C)
        if self.proj_ln is not None:
            ln = self.proj_ln.ln
            y = ops.layer_norm_residual(y, ln.weight, ln.bias, ln.eps)
        m = self.mlp(y, H, W, fc2_bias_grad_elsewhere)
        return (m, x) if passthrough else m


def window_partition(x, window_size):
    """(B, H, W, C) -> (num_windows*B, ws, ws, C)   (reference :120-131), one gather kernel."""
    B, H, W, C = x.shape
    if H % window_size or W % window_size:
        raise ValueError("window_partition: H and W must be multiples of window_size")
    return ops.window_gather(x, window_size, 0).view(-1, window_size, window_size, C)


def window_reverse(windows, window_size, H, W):
    """(num_windows*B, ws, ws, C) -> (B, H, W, C)   (reference :134-147), one scatter kernel."""
    B = int(windows.shape[0] / (H * W / window_size / window_size))
    C = windows.shape[-1]
    return ops.window_scatter(windows.reshape(-1, window_size * window_size, C), B, H, W, window_size, 0)


class ShiftMask:
    """What BasicLayer hands to its blocks in place of the materialised [nW,N,N] mask tensor
    (reference :874-892): the geometry, from which the attention kernel derives the {0,-100} mask on
    the fly.  ``.tensor()`` materialises it (one small kernel) for callers that want the real thing."""

    def __init__(self, H, W, window_size, shift_size, device):
        self.H, self.W, self.window_size, self.shift_size, self.device = H, W, window_size, shift_size, device

    def tensor(self):
        return ops.shift_mask(self.H, self.W, self.window_size, self.shift_size, self.device)


class WindowAttention(nn.Module):
    """Window multi-head self attention, Swin-V2 flavour: scaled-cosine logits with a clamped learnable
    per-head temperature and a log-spaced continuous position bias (reference :150-336)."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.,
                 relative_coords_table_type='norm8_log', rpe_hidden_dim=512, rpe_output_type='normal',
                 attn_type='normal', mlpfp32=False, pretrain_window_size=-1):
        super().__init__()
        if attn_type not in ('cosine_mh', 'normal'):
            raise NotImplementedError(f"attn_type={attn_type!r}")
        if relative_coords_table_type not in ('norm8_log_bylayer', 'norm8_log', 'linear', 'linear_bylayer', 'none'):
            raise NotImplementedError(f"relative_coords_table_type={relative_coords_table_type!r}")
        if rpe_output_type not in ('sigmoid', 'normal'):
            raise NotImplementedError(f"rpe_output_type={rpe_output_type!r}")
        if attn_drop != 0. or proj_drop != 0. or mlpfp32:
            raise NotImplementedError("b200swin.WindowAttention: dropout / mlpfp32 variants are not built")
        if dim % num_heads or dim // num_heads != 32:
            raise NotImplementedError("b200swin.WindowAttention: head_dim must be 32 (true for every Swin-V2 size)")
        self.dim = dim
        self.window_size = to_2tuple(window_size)
        if self.window_size[0] != self.window_size[1]:
            raise NotImplementedError("square windows only")
        self.num_heads = num_heads
        self.mlpfp32 = mlpfp32
        self.attn_type = attn_type
        self.rpe_output_type = rpe_output_type
        self.relative_coords_table_type = relative_coords_table_type

        if attn_type == 'cosine_mh':
            self.logit_scale = nn.Parameter(torch.log(10 * torch.ones((num_heads, 1, 1))), requires_grad=True)
        else:                                   # plain scaled dot product (reference :178-180, :296-298)
            self.scale = qk_scale or (dim // num_heads) ** -0.5
        Wh, Ww = self.window_size
        if relative_coords_table_type != 'none':
            self.rpe_mlp = nn.Sequential(nn.Linear(2, rpe_hidden_dim, bias=True), nn.ReLU(inplace=True),
                                         LinearFP32(rpe_hidden_dim, num_heads, bias=False))
            # relative_coords_table: offsets in [-(ws-1), ws-1]^2, normalised and log-spaced (reference :190-239)
            rh = torch.arange(-(Wh - 1), Wh, dtype=torch.float32)
            rw = torch.arange(-(Ww - 1), Ww, dtype=torch.float32)
            table = torch.stack(torch.meshgrid(rh, rw, indexing='ij'), dim=-1).unsqueeze(0).contiguous()
            if relative_coords_table_type in ('linear', 'norm8_log'):
                den = (Wh - 1, Ww - 1)
            else:
                den = (pretrain_window_size - 1, pretrain_window_size - 1)
            table[..., 0] /= den[0]
            table[..., 1] /= den[1]
            if relative_coords_table_type.startswith('norm8_log'):
                table *= 8
                table = torch.sign(table) * torch.log2(torch.abs(table) + 1.0) / np.log2(8)
            self.register_buffer("relative_coords_table", table)
        else:                                   # the Swin-V1 learned table (reference :241-244)
            self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * Wh - 1) * (2 * Ww - 1), num_heads))
            trunc_normal_(self.relative_position_bias_table, std=.02)
        # pair-wise relative position index (reference :249-259)
        ys = torch.arange(Wh).repeat_interleave(Ww)
        xs = torch.arange(Ww).repeat(Wh)
        rel = (ys[:, None] - ys[None, :] + Wh - 1) * (2 * Ww - 1) + (xs[:, None] - xs[None, :] + Ww - 1)
        self.register_buffer("relative_position_index", rel)

        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            self.q_bias = None
            self.v_bias = None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)

    # -- small host-side pieces (a few kFLOP; autograd carries their gradients) ---------------------
    def _fused_small_ops(self):
        if self.relative_coords_table_type == 'none' or self.attn_type != 'cosine_mh':
            return False
        l2 = self.rpe_mlp[2]
        return (self.rpe_output_type == 'sigmoid' and self.relative_coords_table.is_cuda and l2.bias is None
                and self.num_heads <= 64)

    def _bias_table(self):
        """[(2ws-1)^2, nH] fp32: rpe_mlp(coords) then 16*sigmoid (reference :304-313), or the learned table."""
        if self.relative_coords_table_type == 'none':
            t = self.relative_position_bias_table.float()
            return 16 * torch.sigmoid(t) if self.rpe_output_type == 'sigmoid' else t
        l0, l2 = self.rpe_mlp[0], self.rpe_mlp[2]
        if self._fused_small_ops():
            # one kernel forward, one backward (the PyTorch graph below is ~15 latency-bound launches per block)
            return ops.cpb_table(self.relative_coords_table, l0.weight, l0.bias, l2.weight)
        with torch.autocast('cuda', enabled=False):
            t = self.rpe_mlp(self.relative_coords_table.float()).view(-1, self.num_heads)
            if self.rpe_output_type == 'sigmoid':
                t = 16 * torch.sigmoid(t)
        return t

    def _scale(self):
        """exp(min(logit_scale, ln 100)) per head (reference :294, without its hard-coded cuda:0); the constant
        qk_scale for attn_type='normal'."""
        if self.attn_type == 'normal':
            return torch.full((self.num_heads,), float(self.scale), dtype=torch.float32, device=self.qkv.weight.device)
        return torch.clamp(self.logit_scale.float(), max=_LOGIT_MAX).exp().view(self.num_heads)

    def _table_and_scale(self):
        """Bias table and temperature; on the default configuration both come out of ONE kernel each way."""
        if self._fused_small_ops():
            l0, l2 = self.rpe_mlp[0], self.rpe_mlp[2]
            return ops.cpb_table_and_scale(self.relative_coords_table, l0.weight, l0.bias, l2.weight, self.logit_scale)
        return self._bias_table(), self._scale()

    def _pads(self, needed):
        if not needed or self.q_bias is None:
            return None, None
        if self.attn_type == 'normal':                             # pad tokens: q = q_bias as it is, k = 0, v = v_bias
            return self.q_bias.detach(), self.v_bias
        if getattr(self, '_qpad_pre', None) is not None:       # normalised for the whole stage by BasicLayer.forward
            return self._qpad_pre, self.v_bias
        with torch.no_grad():
            qpad = F.normalize(self.q_bias.float().view(self.num_heads, -1), dim=-1).reshape(-1)
        return qpad, self.v_bias

    def attend(self, x, B, H, W, shift, mask=None, proj_bias_grad_elsewhere=False, passthrough=False):
        """x: [B, H*W, C] natural order -> [B, H*W, C]: qkv GEMM, windowed attention over the (padded,
        rolled) grid, proj GEMM.  proj_bias_grad_elsewhere: the caller's LayerNorm backward returns proj.bias'
        gradient (ops.layer_norm_residual(..., producer_bias=...))."""
        C, nH, ws = self.dim, self.num_heads, self.window_size[0]
        if self.attn_type == 'normal':
            return self._attend_normal(x, B, H, W, shift, mask, proj_bias_grad_elsewhere, passthrough)
        x_alias = None
        if passthrough:        # alias of x for the caller's residual branch: its gradient is added in the qkv dgrad epilogue
            qkv, inv_norm, x_alias = ops.qkv_project(x, self.qkv.weight, self.q_bias, self.v_bias, nH, True)
        else:
            qkv, inv_norm = ops.qkv_project(x, self.qkv.weight, self.q_bias, self.v_bias, nH)
        qpad, vpad = self._pads(H % ws != 0 or W % ws != 0)
        table16, scale = self._table_and_scale()
        o = ops.attention_core(qkv.view(B, H, W, 3 * C), inv_norm, table16, scale, qpad, vpad, mask,
                               B, H, W, C, nH, ws, shift)
        pb = self.proj.bias.detach() if (proj_bias_grad_elsewhere and self.proj.bias is not None) else self.proj.bias
        y = ops.linear(o.view(B, H * W, C), self.proj.weight, pb)
        return (y, x_alias) if passthrough else y

    def _attend_normal(self, x, B, H, W, shift, mask, proj_bias_grad_elsewhere, passthrough):
        """attn_type='normal' (reference :296-298): q * scale . k, no normalisation, no learnable temperature.  The same
        attention kernels with the un-normalised projection, a constant scale and the plain backward (inv_norm = None:
        KV-blocked tcgen05 kernels for bf16, CUDA-core kernels for fp32 or an explicit mask)."""
        C, nH, ws = self.dim, self.num_heads, self.window_size[0]
        bias = None
        if self.q_bias is not None:
            bias = torch.cat((self.q_bias, torch.zeros_like(self.v_bias, requires_grad=False), self.v_bias))
        qkv = ops.linear(x, self.qkv.weight, bias)
        qpad, vpad = self._pads(H % ws != 0 or W % ws != 0)
        o = ops.attention_core(qkv.view(B, H, W, 3 * C), None, self._bias_table(), self._scale(), qpad, vpad, mask,
                               B, H, W, C, nH, ws, shift)
        pb = self.proj.bias.detach() if (proj_bias_grad_elsewhere and self.proj.bias is not None) else self.proj.bias
        y = ops.linear(o.view(B, H * W, C), self.proj.weight, pb)
        return (y, x) if passthrough else y

    def forward(self, x, mask=None):
        """x: (num_windows*B, N, C) window-major; mask: (nW, N, N) additive or None  (reference :275-336)."""
        B_, N, C = x.shape
        ws = self.window_size[0]
        if N != ws * ws:
            raise ValueError(f"WindowAttention: N={N} does not match window {ws}x{ws}")
        if isinstance(mask, ShiftMask):
            mask = mask.tensor()
        if mask is not None and B_ % mask.shape[0]:
            raise ValueError("WindowAttention: batch of windows is not a multiple of the mask's nW")
        return self.attend(x, B_, ws, ws, 0, mask)

    def extra_repr(self) -> str:
        return f'dim={self.dim}, window_size={self.window_size}, num_heads={self.num_heads}'


class _SwinBlockBase(nn.Module):
    def _init_common(self, dim, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop,
                     drop_path, use_mlp_norm, endnorm, act_layer, norm_layer, relative_coords_table_type,
                     rpe_hidden_dim, rpe_output_type, attn_type, mlp_type, mlpfp32, pretrain_window_size):
        if use_mlp_norm or endnorm or mlpfp32 or mlp_type not in ('normal', 'conv', 'conv_ln'):
            raise NotImplementedError("b200swin block: use_mlp_norm / endnorm / mlpfp32 variants are not built")
        self.dim, self.num_heads = dim, num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        self.use_mlp_norm, self.endnorm, self.mlpfp32 = use_mlp_norm, endnorm, mlpfp32
        assert 0 <= self.shift_size < self.window_size, "shift_size must in 0-window_size"
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=to_2tuple(window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop,
                                    relative_coords_table_type=relative_coords_table_type,
                                    rpe_output_type=rpe_output_type, rpe_hidden_dim=rpe_hidden_dim,
                                    attn_type=attn_type, mlpfp32=mlpfp32, pretrain_window_size=pretrain_window_size)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        if mlp_type == 'normal':
            self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        else:                                   # reference :404-409
            self.mlp = ConvMlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop,
                               proj_ln=mlp_type == 'conv_ln')
        self.enorm = None
        self.H = None
        self.W = None

    def _attention(self, x, mask_matrix, proj_bias_grad_elsewhere=False):
        """Attention half: shifted-window attention on natural-order tokens.  With the ShiftMask handle (or
        no shift) everything is fused; an explicit mask tensor takes the general window-major route."""
        H, W = self.H, self.W
        B, L, C = x.shape
        assert L == H * W, f"input feature has wrong size, with L = {L}, H = {H}, W = {W}"
        if self.shift_size > 0 and torch.is_tensor(mask_matrix):
            xw = ops.window_gather(x.view(B, H, W, C), self.window_size, self.shift_size)
            aw = self.attn.attend(xw, xw.shape[0], self.window_size, self.window_size, 0, mask_matrix,
                                  proj_bias_grad_elsewhere)
            return ops.window_scatter(aw, B, H, W, self.window_size, self.shift_size).view(B, L, C)
        return self.attn.attend(x, B, H, W, self.shift_size, None, proj_bias_grad_elsewhere)

    def _fc2(self):
        return self.mlp.mlp.fc2 if isinstance(self.mlp, ConvMlp) else self.mlp.fc2

    def _drop_scale(self, x):
        return self.drop_path.sample_scale(x) if isinstance(self.drop_path, DropPath) else None


class SwinTransformerBlockPost(_SwinBlockBase):
    """Post-norm Swin-V2 block (reference :355-488):
    x = x + DropPath(LN(attn(x)));  x = x + DropPath(LN(mlp(x)))."""

    def __init__(self, dim, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., use_mlp_norm=False, endnorm=False, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, relative_coords_table_type='norm8_log', rpe_hidden_dim=512,
                 rpe_output_type='normal', attn_type='normal', mlp_type='normal', mlpfp32=False,
                 pretrain_window_size=-1):
        super().__init__()
        self._init_common(dim, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop,
                          drop_path, use_mlp_norm, endnorm, act_layer, norm_layer, relative_coords_table_type,
                          rpe_hidden_dim, rpe_output_type, attn_type, mlp_type, mlpfp32, pretrain_window_size)

    def forward(self, x, mask_matrix):
        L = x.shape[1]
        # bf16 path: the bias gradients of proj and fc2 are the column sums of the LayerNorm backward's dx and come
        # out of that kernel (no separate pass over dx)
        fuse = torch.is_grad_enabled() and ops.ln_colsum_supported(ops.compute_dtype(x), self.dim)
        # the residual branches run through aliases of x handed back by the qkv / MLP functions, so that the residual
        # gradients are added inside the dgrad GEMM epilogues (no elementwise gradient-accumulation passes)
        direct = not (self.shift_size > 0 and torch.is_tensor(mask_matrix))
        # bf16 mode keeps the residual stream in fp32 beside the bf16 activations (torch.autocast semantics; with the
        # reference's 1e-5 block-norm initialisation a block's contribution is below a bf16 half-ulp of the stream)
        stream = ops.stream32_supported(ops.compute_dtype(x), self.dim)
        x32 = ops.stream_of_tensor(x) if stream else None
        if direct:
            assert L == self.H * self.W, f"input feature has wrong size, with L = {L}, H = {self.H}, W = {self.W}"
            a, xr = self.attn.attend(x, x.shape[0], self.H, self.W, self.shift_size, None, fuse, True)
        else:
            a, xr = self._attention(x, mask_matrix, fuse), x
        x = ops.layer_norm_residual(a, self.norm1.weight, self.norm1.bias, self.norm1.eps, residual=xr,
                                    row_scale=self._drop_scale(x), rows_per_scale=L,
                                    producer_bias=self.attn.proj.bias if fuse else None, residual32=x32, stream32=stream)
        x, x32 = x if stream else (x, None)
        m, xr = self.mlp(x, self.H, self.W, fuse, True)
        y = ops.layer_norm_residual(m, self.norm2.weight, self.norm2.bias, self.norm2.eps, residual=xr,
                                    row_scale=self._drop_scale(x), rows_per_scale=L,
                                    producer_bias=self._fc2().bias if fuse else None, residual32=x32, stream32=stream)
        return ops.attach_stream(*y) if stream else y


class SwinTransformerBlockPre(_SwinBlockBase):
    """Pre-norm variant with optional layer-scale gamma_1/gamma_2 (reference :491-630)."""

    def __init__(self, dim, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., use_mlp_norm=False, endnorm=False, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, init_values=None, relative_coords_table_type='norm8_log',
                 rpe_hidden_dim=512, rpe_output_type='normal', attn_type='normal', mlp_type='normal', mlpfp32=False,
                 pretrain_window_size=-1):
        super().__init__()
        self._init_common(dim, num_heads, window_size, shift_size, mlp_ratio, qkv_bias, qk_scale, drop, attn_drop,
                          drop_path, use_mlp_norm, endnorm, act_layer, norm_layer, relative_coords_table_type,
                          rpe_hidden_dim, rpe_output_type, attn_type, mlp_type, mlpfp32, pretrain_window_size)
        if init_values is not None and init_values >= 0:
            self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
            self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        else:
            self.gamma_1, self.gamma_2 = 1.0, 1.0

    def forward(self, x, mask_matrix):
        shortcut = x
        y = ops.layer_norm_residual(x, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        a = self._attention(y, mask_matrix)
        x = shortcut + self.drop_path(self.gamma_1 * a)
        m = self.mlp(ops.layer_norm_residual(x, self.norm2.weight, self.norm2.bias, self.norm2.eps), self.H, self.W)
        return x + self.drop_path(self.gamma_2 * m)


class PatchMerging(nn.Module):
    """2x2 patch merging (reference :633-678): strided gather + Linear(4C, 2C, no bias) + LayerNorm.
    The pad + gather is one index-map kernel (and its adjoint in the backward); the contraction and the norm run in
    the b200swin GEMM / LayerNorm kernels."""

    def __init__(self, dim, norm_layer=nn.LayerNorm, postnorm=True):
        super().__init__()
        self.dim = dim
        self.postnorm = postnorm
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(2 * dim) if postnorm else norm_layer(4 * dim)

    def forward(self, x, H, W):
        B, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        x = ops.patch_merge(x.view(B, H, W, C))          # pad to even + 2x2 gather in one kernel: [B, H2*W2, 4C]
        if self.postnorm:
            x = ops.linear(x, self.reduction.weight, None)
            if ops.stream32_supported(x.dtype, 2 * C):       # opens the fp32 residual stream of the next stage
                return ops.attach_stream(*ops.layer_norm_residual(x, self.norm.weight, self.norm.bias, self.norm.eps,
                                                                  stream32=True))
            return ops.layer_norm_residual(x, self.norm.weight, self.norm.bias, self.norm.eps)
        x = ops.layer_norm_residual(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return ops.linear(x, self.reduction.weight, None)


class BasicLayer(nn.Module):
    """One Swin stage: `depth` blocks alternating plain / shifted windows, then patch merging
    (reference :750-915).  forward -> (x, H, W, x_down, Wh, Ww)."""

    def __init__(self, dim, depth, num_heads, window_size=7, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False,
                 checkpoint_blocks=255, init_values=None, endnorm_interval=-1, use_mlp_norm=False, use_shift=True,
                 relative_coords_table_type='norm8_log', rpe_hidden_dim=512, rpe_output_type='normal',
                 attn_type='normal', mlp_type='normal', mlpfp32_blocks=[-1], postnorm=True, pretrain_window_size=-1):
        super().__init__()
        self.window_size = window_size
        self.shift_size = window_size // 2
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.checkpoint_blocks = checkpoint_blocks
        self.init_values = init_values if init_values is not None else 0.0
        self.endnorm_interval = endnorm_interval
        self.mlpfp32_blocks = mlpfp32_blocks
        self.postnorm = postnorm
        if endnorm_interval > 0 or any(i in mlpfp32_blocks for i in range(depth)):
            raise NotImplementedError("b200swin.BasicLayer: endnorm / mlpfp32 blocks are not built")
        common = dict(dim=dim, num_heads=num_heads, window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                      qk_scale=qk_scale, drop=drop, attn_drop=attn_drop, norm_layer=norm_layer,
                      use_mlp_norm=use_mlp_norm, relative_coords_table_type=relative_coords_table_type,
                      rpe_hidden_dim=rpe_hidden_dim, rpe_output_type=rpe_output_type, attn_type=attn_type,
                      mlp_type=mlp_type, pretrain_window_size=pretrain_window_size)
        blocks = []
        for i in range(depth):
            kw = dict(common, shift_size=0 if (i % 2 == 0) or (not use_shift) else window_size // 2,
                      drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path)
            blocks.append(SwinTransformerBlockPost(**kw) if postnorm
                          else SwinTransformerBlockPre(init_values=init_values, **kw))
        self.blocks = nn.ModuleList(blocks)
        self.downsample = downsample(dim=dim, norm_layer=norm_layer, postnorm=postnorm) if downsample is not None else None

    def forward(self, x, H, W):
        # the reference rebuilds a [nW,N,N] mask tensor here on every call (:874-892); the kernels derive the
        # same {0,-100} values from token coordinates, so only a handle travels to the blocks.
        attn_mask = ShiftMask(H, W, self.window_size, self.shift_size, x.device)
        staged = self._stage_pad_queries(H, W)
        try:
            for blk in self.blocks:
                blk.H, blk.W = H, W
                if self.use_checkpoint and torch.is_grad_enabled():
                    x = checkpoint.checkpoint(blk, x, attn_mask, use_reentrant=False)
                else:
                    x = blk(x, attn_mask)
        finally:
            for a in staged:
                a._qpad_pre = None
        if self.downsample is not None:
            x_down = self.downsample(x, H, W)
            return x, H, W, x_down, (H + 1) // 2, (W + 1) // 2
        return x, H, W, x, H, W

    def _stage_pad_queries(self, H, W):
        """Padded grids: the query of a padding token is normalize(q_bias) per head (the reference pads x with zeros
        before the qkv Linear, :446-452 / :283-291).  One batched normalisation for all blocks of the stage instead
        of three small launches per block."""
        attns = [blk.attn for blk in self.blocks if getattr(blk.attn, 'q_bias', None) is not None]
        if len(attns) < 2 or (self.use_checkpoint and torch.is_grad_enabled()):
            return []
        ws = attns[0].window_size[0]
        if (H % ws == 0 and W % ws == 0) or any(a.window_size[0] != ws or a.q_bias.shape != attns[0].q_bias.shape
                                                 for a in attns):
            return []
        with torch.no_grad():
            q = torch.stack([a.q_bias for a in attns]).float().view(len(attns), attns[0].num_heads, -1)
            q = F.normalize(q, dim=-1).view(len(attns), -1)
        for a, row in zip(attns, q):
            a._qpad_pre = row
        return attns

    def _init_block_norm_weights(self):
        for blk in self.blocks:
            nn.init.constant_(blk.norm1.bias, 0)
            nn.init.constant_(blk.norm1.weight, self.init_values)
            nn.init.constant_(blk.norm2.bias, 0)
            nn.init.constant_(blk.norm2.weight, self.init_values)


class PatchEmbed(nn.Module):
    """4x4 stride-4 conv patch embedding + LayerNorm (reference :918-957).  A conv whose stride equals its kernel is a
    GEMM over non-overlapping patches: one patchify kernel (NCHW image -> [patches, Cin*ph*pw]) feeds the b200swin GEMM
    with the conv weight viewed as [E, Cin*ph*pw]; the result is already in token layout for the norm and the blocks
    (no NCHW <-> NHWC transposes, no cuDNN)."""

    def __init__(self, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.patch_size = to_2tuple(patch_size)
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=self.patch_size, stride=self.patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward_tokens(self, x):
        """x[B,Cin,H,W] -> (tokens [B, Wh*Ww, E], Wh, Ww)."""
        B = x.shape[0]
        ph, pw = self.patch_size
        K = self.in_chans * ph * pw
        if K % 8 == 0 and self.embed_dim % 8 == 0 and not (x.requires_grad and torch.is_grad_enabled()):
            cols, Wh, Ww = ops.patchify(x, ph, pw, ops.compute_dtype(x))
            t = ops.linear(cols, self.proj.weight.view(self.embed_dim, K), self.proj.bias).view(B, Wh * Ww, self.embed_dim)
        else:
            # general case (input gradient wanted, or rows that are not 16-byte multiples): cuDNN conv + transpose
            _, _, H, W = x.size()
            if W % pw != 0:
                x = F.pad(x, (0, pw - W % pw))
            if H % ph != 0:
                x = F.pad(x, (0, 0, 0, ph - H % ph))
            y = self.proj(x)
            Wh, Ww = y.size(2), y.size(3)
            t = y.flatten(2).transpose(1, 2).contiguous()
        if self.norm is not None:
            if ops.stream32_supported(t.dtype, self.embed_dim):      # opens the fp32 residual stream of stage 0
                t = ops.attach_stream(*ops.layer_norm_residual(t, self.norm.weight, self.norm.bias, self.norm.eps,
                                                               stream32=True))
            else:
                t = ops.layer_norm_residual(t, self.norm.weight, self.norm.bias, self.norm.eps)
        return t, Wh, Ww

    def forward(self, x):
        """Reference contract: NCHW feature map [B, E, Wh, Ww]."""
        t, Wh, Ww = self.forward_tokens(x)
        return t.transpose(1, 2).reshape(-1, self.embed_dim, Wh, Ww)


class SwinTransformerV2(nn.Module):
    """Swin-V2 backbone with the reference's constructor and forward contract (reference :995-1282):
    forward(x[B,3,H,W]) -> list of fp32 NCHW feature maps for `out_indices`."""

    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], window_size=7, mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0.,
                 attn_drop_rate=0., drop_path_rate=0.1, norm_layer=partial(LayerNormFP32, eps=1e-6), ape=False,
                 patch_norm=True, use_checkpoint=False, init_values=1e-5, endnorm_interval=-1,
                 use_mlp_norm_layers=[], relative_coords_table_type='norm8_log_bylayer', rpe_hidden_dim=512,
                 attn_type='cosine_mh', rpe_output_type='sigmoid', rpe_wd=False, postnorm=True, mlp_type='normal',
                 patch_embed_type='normal', patch_merge_type='normal', strid16=False,
                 checkpoint_blocks=[255, 255, 255, 255], mlpfp32_layer_blocks=[[-1], [-1], [-1], [-1]],
                 out_indices=(3,), frozen_stages=-1, use_shift=True, rpe_interpolation='geo',
                 pretrain_window_size=[-1, -1, -1, -1], **kwargs):
        super().__init__()
        if ape or strid16 or patch_embed_type != 'normal' or patch_merge_type != 'normal' or use_mlp_norm_layers:
            raise NotImplementedError("b200swin.SwinTransformerV2: ape / strid16 / conv patch variants are not built")
        if drop_rate != 0. or attn_drop_rate != 0.:
            raise NotImplementedError("b200swin.SwinTransformerV2: dropout is not built (the reference uses 0)")
        # reference_rng=True: every DropPath call draws its own mask from the current generator, in the reference's call
        # order and with timm's call (one bernoulli_ over the batch per branch), so a seeded training run sees the same
        # stochastic-depth masks as the reference.  Default: all masks of a forward drawn in one batched launch (same
        # distribution, different use of the generator), ~94 fewer launch-bound kernels per step.
        self.reference_rng = bool(kwargs.pop('reference_rng', False))
        self.pretrain_img_size = pretrain_img_size
        self.depths = depths
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.ape = ape
        self.patch_norm = patch_norm
        self.out_indices = out_indices
        self.frozen_stages = frozen_stages
        self.rpe_interpolation = rpe_interpolation
        self.mlp_ratio = mlp_ratio
        self.endnorm_interval = endnorm_interval
        self.use_mlp_norm_layers = use_mlp_norm_layers
        self.relative_coords_table_type = relative_coords_table_type
        self.rpe_hidden_dim = rpe_hidden_dim
        self.rpe_output_type = rpe_output_type
        self.rpe_wd = rpe_wd
        self.attn_type = attn_type
        self.postnorm = postnorm
        self.mlp_type = mlp_type
        self.strid16 = strid16

        def per_layer(v, typ, what):
            if isinstance(v, list):
                return v
            if isinstance(v, typ):
                return [v] * self.num_layers
            raise TypeError(f"We only support list or {typ.__name__} for {what}")

        window_size = per_layer(window_size, int, "window size")
        use_shift = per_layer(use_shift, bool, "use_shift")
        use_checkpoint = per_layer(use_checkpoint, bool, "use_checkpoint")

        self.patch_embed = PatchEmbed(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        num_features = []
        for i in range(self.num_layers):
            cur_dim = int(embed_dim * 2 ** i)
            num_features.append(cur_dim)
            self.layers.append(BasicLayer(
                dim=cur_dim, depth=depths[i], num_heads=num_heads[i], window_size=window_size[i], mlp_ratio=mlp_ratio,
                qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate,
                drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if i < self.num_layers - 1 else None, use_checkpoint=use_checkpoint[i],
                checkpoint_blocks=checkpoint_blocks[i], init_values=init_values, endnorm_interval=endnorm_interval,
                use_mlp_norm=False, use_shift=use_shift[i], relative_coords_table_type=relative_coords_table_type,
                rpe_hidden_dim=rpe_hidden_dim, rpe_output_type=rpe_output_type, attn_type=attn_type,
                mlp_type=mlp_type, mlpfp32_blocks=mlpfp32_layer_blocks[i], postnorm=postnorm,
                pretrain_window_size=pretrain_window_size[i]))
        self.num_features = num_features
        for i in out_indices[:self.num_layers]:
            self.add_module(f'norm{i}', norm_layer(num_features[i]))
        self._freeze_stages()

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for p in self.patch_embed.parameters():
                p.requires_grad = False
        if self.frozen_stages >= 2:
            self.pos_drop.eval()
            for i in range(0, self.frozen_stages - 1):
                m = self.layers[i]
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False

    def init_weights(self, pretrained=None):
        """trunc-normal(0.02) linears/convs, unit LayerNorms, block norms gamma=init_values (reference :1218-1249).
        `pretrained`: path to a state_dict checkpoint (optionally nested under 'model'/'state_dict' and prefixed
        with 'encoder.'/'backbone.'/'module.'); keys are the reference's, loaded non-strictly."""

        def _init(m):
            if isinstance(m, nn.Linear):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
            elif isinstance(m, nn.Conv2d):
                trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

        self.apply(_init)
        for layer in self.layers:
            layer._init_block_norm_weights()
        if isinstance(pretrained, str) and pretrained != '':
            ckpt = torch.load(pretrained, map_location='cpu')
            for k in ('model', 'state_dict', 'model_state_dict'):
                if isinstance(ckpt, dict) and k in ckpt:
                    ckpt = ckpt[k]
            sd = {}
            for k, v in ckpt.items():
                for pre in ('module.', 'backbone.', 'encoder.'):
                    if k.startswith(pre):
                        k = k[len(pre):]
                sd[k] = v
            own = self.state_dict()
            sd = {k: v for k, v in sd.items() if k in own and v.shape == own[k].shape}
            self.load_state_dict(sd, strict=False)
        elif pretrained is not None and pretrained != '':
            raise TypeError('pretrained must be a str or None')

    def _draw_drop_paths(self, B, device):
        """All stochastic-depth masks of one forward in two launches instead of two per DropPath call (each block draws
        twice: ~94 launch-bound kernels per step on Swin-V2-B).  Same distribution as the per-call draw,
        Bernoulli(keep) / keep per sample; skipped under activation checkpointing, whose recompute relies on replaying
        the RNG of per-call draws."""
        if not self.training or self.reference_rng or any(getattr(l, 'use_checkpoint', False) for l in self.layers):
            return []
        mods = [m for m in self.modules() if isinstance(m, DropPath) and m.training and 0.0 < m.drop_prob < 1.0]
        if not mods:
            return []
        key = (str(device), tuple(m.drop_prob for m in mods))
        if getattr(self, '_dp_keep_key', None) != key:
            keep = torch.tensor([1.0 - m.drop_prob for m in mods for _ in range(2)], dtype=torch.float32)
            self._dp_keep, self._dp_keep_key = keep.to(device).view(-1, 1), key
        masks = torch.bernoulli(self._dp_keep.expand(-1, B)).div_(self._dp_keep)
        for i, m in enumerate(mods):
            m._drawn = [masks[2 * i], masks[2 * i + 1]]
        return mods

    def forward(self, x):
        drawn = self._draw_drop_paths(x.shape[0], x.device)
        try:
            return self._forward(x)
        finally:
            for m in drawn:
                m._drawn = []

    def _forward(self, x):
        x, Wh, Ww = self.patch_embed.forward_tokens(x)
        outs = []
        for i in range(self.num_layers):
            x_out, H, W, x, Wh, Ww = self.layers[i](x, Wh, Ww)
            if i in self.out_indices:
                norm = getattr(self, f'norm{i}')
                x32 = ops.stream_of_tensor(x_out)
                with torch.autocast('cuda', enabled=False):
                    if x32 is not None and x_out.requires_grad:
                        # values from the fp32 stream, gradient through the bf16 tensor (straight-through on the rounding)
                        xin = x_out.float() + (x32 - x_out.detach().float())
                    else:
                        xin = x32 if x32 is not None else x_out.float()
                    y = ops.layer_norm_residual(xin, norm.weight, norm.bias, norm.eps)
                outs.append(y.view(-1, H, W, self.num_features[i]).permute(0, 3, 1, 2).contiguous())
        return outs

    def train(self, mode=True):
        """Train mode that keeps frozen stages frozen (reference :1279-1282; returns self, the reference's
        version returns None by omission)."""
        super().train(mode)
        self._freeze_stages()
        return self
