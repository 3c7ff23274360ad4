"""Drop-ins for the global-attention encoder layer of the reference's multimodal path (models/cnn_transformer.py):

    b200swin.cnn_transformer.MultiheadAttention   <->  torch.nn.MultiheadAttention as the reference constructs it at
                                                       models/cnn_transformer.py:192 (batch_first, no masks, no dropout)
    b200swin.cnn_transformer.Transformer_Encoder  <->  models/cnn_transformer.py:176-216

Same constructor arguments, attribute names and ``state_dict`` keys (``self_attn.in_proj_weight``,
``self_attn.in_proj_bias``, ``self_attn.out_proj.{weight,bias}``, ``norm1|norm2.*``, ``ffn1.0.*``, ``ffn2.0.*``), so the
reference's checkpoints load key for key.  The attention core runs on the b200swin global-attention kernels
(csrc/attn_global.cu: 1200 tokens, 8 heads of 64 for a 480x640 frame), the projections and the feed-forward on the tcgen05
GEMM, the LayerNorms on the bulk-copy LayerNorm kernel.  The rest of the reference's ``cnn_transformer`` (ResNet-50
trunk, sine position embedding) is outside SURVEY.md section 8 and runs as the reference's own modules.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class MultiheadAttention(nn.Module):
    """nn.MultiheadAttention for the configuration the reference uses: self- or cross-attention without masks or dropout,
    kdim = vdim = embed_dim.  forward returns (output, head-averaged weights or None) like the original."""

    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None,
                 vdim=None, batch_first=False, device=None, dtype=None):
        super().__init__()
        if dropout != 0. or add_bias_kv or add_zero_attn or (kdim not in (None, embed_dim)) or (vdim not in (None, embed_dim)):
            raise NotImplementedError("b200swin.MultiheadAttention: dropout / bias_kv / zero_attn / kdim / vdim variants are not built")
        if embed_dim % num_heads or embed_dim // num_heads not in (32, 64):
            raise NotImplementedError("b200swin.MultiheadAttention: head_dim must be 32 or 64")
        kw = {"device": device, "dtype": dtype}
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.kdim = self.vdim = embed_dim
        self.dropout, self.batch_first = dropout, batch_first
        self._qkv_same_embed_dim = True
        self.in_proj_weight = nn.Parameter(torch.empty((3 * embed_dim, embed_dim), **kw))
        if bias:
            self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim, **kw))
        else:
            self.register_parameter('in_proj_bias', None)
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias, **kw)
        self._reset_parameters()

    def _reset_parameters(self):
        nn.init.xavier_uniform_(self.in_proj_weight)
        if self.in_proj_bias is not None:
            nn.init.constant_(self.in_proj_bias, 0.)
            nn.init.constant_(self.out_proj.bias, 0.)

    def _proj(self, x, lo, hi):
        b = self.in_proj_bias[lo:hi] if self.in_proj_bias is not None else None
        if x.dtype == torch.float32 and ops.compute_dtype(x) == torch.bfloat16:
            t = ops.bf16_twin_of(x)                   # written by the previous layer's LayerNorm kernel: no cast pass
            if t is not None:
                x = t
        return ops.linear(x, self.in_proj_weight[lo:hi], b)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None,
                average_attn_weights=True, is_causal=False):
        if key_padding_mask is not None or attn_mask is not None or is_causal:
            raise NotImplementedError("b200swin.MultiheadAttention: masks are not built (the reference passes none)")
        if query.dim() != 3:
            raise ValueError("b200swin.MultiheadAttention: batched [B, N, E] / [N, B, E] inputs only")
        same_qk, same_kv = key is query, value is key
        if not self.batch_first:
            query = query.transpose(0, 1)
            key = query if same_qk else key.transpose(0, 1)
            value = key if same_kv else value.transpose(0, 1)
        E, nH = self.embed_dim, self.num_heads
        if same_qk and same_kv:
            qkv = self._proj(query, 0, 3 * E)
            out, lse = ops.mha_core(qkv, None, None, nH, packed='qkv')
            q, k = qkv[..., :E], qkv[..., E:2 * E]
        elif same_qk:                       # the reference: q = k = feat + pos, v = feat   (cnn_transformer.py:198-201)
            qk = self._proj(query, 0, 2 * E)
            v = self._proj(value, 2 * E, 3 * E)
            out, lse = ops.mha_core(qk, v, None, nH, packed='qk_v')
            q, k = qk[..., :E], qk[..., E:]
        else:
            q, k, v = self._proj(query, 0, E), self._proj(key, E, 2 * E), self._proj(value, 2 * E, 3 * E)
            out, lse = ops.mha_core(q, k, v, nH)
        y = ops.linear(out, self.out_proj.weight, self.out_proj.bias)
        w = None
        if need_weights:
            if not average_attn_weights:
                raise NotImplementedError("b200swin.MultiheadAttention: per-head weights are not built")
            w = ops.mha_avg_weights(q, k, lse, nH)
        if not self.batch_first:
            y = y.transpose(0, 1)
        return y, w


class Transformer_Encoder(nn.Module):
    """One post-norm encoder layer: global self-attention with the position embedding added to q and k only, then a ReLU
    feed-forward (reference models/cnn_transformer.py:176-216)."""

    def __init__(self, args, hidden_dim):
        super().__init__()
        self.args = args
        self.hidden_dim = hidden_dim
        self.dim_feedforward = args.transformer_ff_dim
        self.dropout = nn.Dropout(0.)
        if self.hidden_dim == 256:
            num_heads = 4
        elif self.hidden_dim == 512:
            num_heads = 8
        else:
            raise ValueError("Transformer_Encoder: hidden_dim must be 256 or 512 (reference :187-190)")
        self.self_attn = MultiheadAttention(self.hidden_dim, num_heads=num_heads, batch_first=True)
        self.norm1 = nn.LayerNorm(self.hidden_dim)
        self.ffn1 = nn.Sequential(nn.Linear(self.hidden_dim, self.dim_feedforward), nn.ReLU())
        self.ffn2 = nn.Sequential(nn.Linear(self.dim_feedforward, self.hidden_dim))
        self.norm2 = nn.LayerNorm(self.hidden_dim)

    def forward(self, img_feat, img_pos):
        q = k = img_feat + img_pos
        v = img_feat
        x, _ = self.self_attn(q, k, v, need_weights=False)       # the reference computes the weights and drops them
        # x = norm1(v + x): add, LayerNorm and the bf16 copy for the next GEMM in one kernel; the stream stays fp32
        amp = ops.compute_dtype(img_feat) == torch.bfloat16
        x32, x16 = ops.layer_norm_sum(v, x, self.norm1.weight, self.norm1.bias, self.norm1.eps, want16=amp)
        # Linear -> ReLU -> Linear as two GEMMs, the ReLU (and its 0/1 derivative for the backward) in the first epilogue
        x2 = ops.mlp(x16 if amp else x32, self.ffn1[0].weight, self.ffn1[0].bias, self.ffn2[0].weight, self.ffn2[0].bias,
                     relu=True)
        y32, _ = ops.layer_norm_sum(x32, x2, self.norm2.weight, self.norm2.bias, self.norm2.eps, want16=amp)
        return y32                                               # carries its bf16 copy for the next layer's v projection
