"""CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py) for the global-attention encoder layer of the
reference's multimodal path: ``Transformer_Encoder`` in ``models/cnn_transformer.py:176-216`` and the
``torch.nn.MultiheadAttention`` it calls (``:192``, ``:201``).

``nn.MultiheadAttention`` is third-party code (PyTorch; the reference pins torch 1.8.0+cu111 in settings.sh, this image
has 2.11): its published algorithm for batch_first self-attention without masks is restated here --
    q, k, v = x_q W_q^T + b_q,  x_k W_k^T + b_k,  x_v W_v^T + b_v   (rows 0:E, E:2E, 2E:3E of in_proj_weight / in_proj_bias)
    per head h (head_dim = E / nH):  P_h = softmax(q_h k_h^T / sqrt(head_dim)),  o_h = P_h v_h
    y = concat_h(o_h) W_o^T + b_o,   weights = mean_h P_h   (need_weights=True, average_attn_weights=True)
-- and pinned two ways: against the golden vectors generated from the reference's own ``Transformer_Encoder``
(tests/golden/tenc_*.npz, tests/golden/make_golden.py) and against ``torch.nn.MultiheadAttention`` itself in
tests/test_oracle_golden.py.
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn.functional as F


def multihead_attention(query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, sd: Mapping[str, torch.Tensor],
                        num_heads: int, need_weights: bool = True):
    """nn.MultiheadAttention(E, num_heads, batch_first=True).forward(query, key, value) -> (y, mean-over-heads weights).
    query [B, Nq, E], key / value [B, Nk, E]; sd holds in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias."""
    B, Nq, E = query.shape
    Nk = key.shape[1]
    hd = E // num_heads
    W, b = sd["in_proj_weight"], sd["in_proj_bias"]
    q = F.linear(query, W[:E], b[:E]).view(B, Nq, num_heads, hd).transpose(1, 2)
    k = F.linear(key, W[E:2 * E], b[E:2 * E]).view(B, Nk, num_heads, hd).transpose(1, 2)
    v = F.linear(value, W[2 * E:], b[2 * E:]).view(B, Nk, num_heads, hd).transpose(1, 2)
    p = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(B, Nq, E)
    y = F.linear(o, sd["out_proj.weight"], sd["out_proj.bias"])
    return y, (p.mean(dim=1) if need_weights else None)


def transformer_encoder_layer(img_feat: torch.Tensor, img_pos: torch.Tensor, sd: Mapping[str, torch.Tensor],
                              num_heads: int, eps: float = 1e-5) -> torch.Tensor:
    """Transformer_Encoder.forward (models/cnn_transformer.py:197-216): position embedding on q and k only, post-norm
    residual blocks, ReLU feed-forward, dropout 0."""
    sub = {k[len("self_attn."):]: v for k, v in sd.items() if k.startswith("self_attn.")}
    qk = img_feat + img_pos                                                               # :198
    x, _ = multihead_attention(qk, qk, img_feat, sub, num_heads, need_weights=False)      # :199-201
    x = img_feat + x                                                                      # :202
    E = x.shape[-1]
    x = F.layer_norm(x, (E,), sd["norm1.weight"], sd["norm1.bias"], eps)                  # :203
    x2 = F.relu(F.linear(x, sd["ffn1.0.weight"], sd["ffn1.0.bias"]))                      # :206
    x2 = F.linear(x2, sd["ffn2.0.weight"], sd["ffn2.0.bias"])                             # :207
    x = x + x2                                                                            # :208
    return F.layer_norm(x, (E,), sd["norm2.weight"], sd["norm2.bias"], eps)               # :209
