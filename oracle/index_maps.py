"""Integer-exact index maps of the Swin-V2 window machinery (numpy, CPU).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function cites the reference
lines (relative to ``/root/reference``) whose behaviour it restates.  All maps are
expressed as *flat source index per destination element* so that a CUDA gather /
scatter can be compared bit-for-bit by pushing an ``arange`` tensor through it.
"""
from __future__ import annotations

import numpy as np


def padded_size(H: int, W: int, ws: int) -> tuple[int, int]:
    """Hp, Wp after right/bottom zero padding to a multiple of ``ws``.

    models/swin_transformer_v2.py:429-434 (block) and :874-875 (BasicLayer)."""
    return (H + ws - 1) // ws * ws, (W + ws - 1) // ws * ws


def partition_src_index(B: int, Hp: int, Wp: int, ws: int) -> np.ndarray:
    """``window_partition`` as a gather map.

    Returns int64 ``[B*nW, ws*ws]``: flat token index ``b*Hp*Wp + i*Wp + j`` of the
    source token for window ``w`` / in-window token ``t``.
    models/swin_transformer_v2.py:120-131: window id = b*nW + (i//ws)*(Wp//ws) + j//ws,
    token = (i%ws)*ws + j%ws."""
    assert Hp % ws == 0 and Wp % ws == 0
    nWh, nWw = Hp // ws, Wp // ws
    out = np.empty((B, nWh, nWw, ws, ws), dtype=np.int64)
    b = np.arange(B).reshape(B, 1, 1, 1, 1)
    wh = np.arange(nWh).reshape(1, nWh, 1, 1, 1)
    ww = np.arange(nWw).reshape(1, 1, nWw, 1, 1)
    r = np.arange(ws).reshape(1, 1, 1, ws, 1)
    c = np.arange(ws).reshape(1, 1, 1, 1, ws)
    out[...] = b * Hp * Wp + (wh * ws + r) * Wp + (ww * ws + c)
    return out.reshape(B * nWh * nWw, ws * ws)


def reverse_src_index(B: int, Hp: int, Wp: int, ws: int) -> np.ndarray:
    """``window_reverse`` as a gather map: int64 ``[B, Hp, Wp]`` holding the flat
    (window*N + token) index each image token is read from.
    models/swin_transformer_v2.py:134-147."""
    fwd = partition_src_index(B, Hp, Wp, ws).reshape(-1)
    inv = np.empty_like(fwd)
    inv[fwd] = np.arange(fwd.size, dtype=np.int64)
    return inv.reshape(B, Hp, Wp)


def roll_src_index(Hp: int, Wp: int, shift_h: int, shift_w: int) -> np.ndarray:
    """``torch.roll(x, shifts=(shift_h, shift_w), dims=(1, 2))`` as a gather map over
    one image: int64 ``[Hp, Wp]`` of flat source ``i*Wp + j``.
    out[i, j] = x[(i - shift_h) mod Hp, (j - shift_w) mod Wp]
    (models/swin_transformer_v2.py:438 uses shifts=(-s,-s), :458 uses (+s,+s))."""
    i = (np.arange(Hp).reshape(Hp, 1) - shift_h) % Hp
    j = (np.arange(Wp).reshape(1, Wp) - shift_w) % Wp
    return (i * Wp + j).astype(np.int64)


def fused_gather_index(B: int, H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """pad -> roll(-shift) -> partition, folded into ONE gather map.

    Returns int64 ``[B*nW, ws*ws]``; entry = flat index ``b*H*W + i*W + j`` into the
    UNPADDED ``[B, H, W]`` token grid, or ``-1`` where the source is a zero pad token.
    models/swin_transformer_v2.py:429-446."""
    Hp, Wp = padded_size(H, W, ws)
    part = partition_src_index(B, Hp, Wp, ws)              # indices into shifted padded grid
    b = part // (Hp * Wp)
    rem = part % (Hp * Wp)
    si, sj = rem // Wp, rem % Wp                           # coords in the shifted grid
    i = (si + shift) % Hp                                  # roll by -shift: shifted[i] = x[(i+s)%Hp]
    j = (sj + shift) % Wp
    src = b * H * W + i * W + j
    src[(i >= H) | (j >= W)] = -1
    return src


def fused_scatter_index(B: int, H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """reverse -> roll(+shift) -> crop, folded into one gather map for the OUTPUT side.

    Returns int64 ``[B, H, W]``: flat (window*N + token) index each kept output token is
    read from.  models/swin_transformer_v2.py:453-463."""
    g = fused_gather_index(B, H, W, ws, shift).reshape(-1)
    out = np.full(B * H * W, -1, dtype=np.int64)
    valid = g >= 0
    out[g[valid]] = np.nonzero(valid)[0]
    assert (out >= 0).all()
    return out.reshape(B, H, W)


def shift_region_ids(Hp: int, Wp: int, ws: int, shift: int) -> np.ndarray:
    """Region id 0..8 of every token of the SHIFTED padded grid, int64 ``[Hp, Wp]``.

    models/swin_transformer_v2.py:876-887: three h-slices x three w-slices
    ``[0,-ws) , [-ws,-shift) , [-shift, end)`` numbered row-major.  Implemented with the
    same slice semantics (python slices, including the degenerate ``shift == 0`` and
    ``Hp == ws`` cases) so the closed form used by the kernels can be tested against it."""
    img = np.zeros((Hp, Wp), dtype=np.int64)
    hs = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    cnt = 0
    for h in hs:
        for w in hs:
            img[h, w] = cnt
            cnt += 1
    return img


def shift_region_ids_closed_form(Hp: int, Wp: int, ws: int, shift: int) -> np.ndarray:
    """Closed form the CUDA kernels use: r(i, L) = [i >= L-ws] + [i >= L-shift];
    id = 3*r(i,Hp) + r(j,Wp).  Valid for 0 < shift < ws (the only case the blocks use)."""
    assert 0 < shift < ws
    i = np.arange(Hp).reshape(Hp, 1)
    j = np.arange(Wp).reshape(1, Wp)
    ri = (i >= Hp - ws).astype(np.int64) + (i >= Hp - shift).astype(np.int64)
    rj = (j >= Wp - ws).astype(np.int64) + (j >= Wp - shift).astype(np.int64)
    return 3 * ri + rj


def shift_attn_mask(H: int, W: int, ws: int, shift: int) -> np.ndarray:
    """The SW-MSA mask ``[nW, N, N]`` float32 in {0, -100} built by BasicLayer.forward
    (models/swin_transformer_v2.py:874-892).  NOTE: -100.0, not -inf."""
    Hp, Wp = padded_size(H, W, ws)
    ids = shift_region_ids(Hp, Wp, ws, shift).reshape(-1)
    part = partition_src_index(1, Hp, Wp, ws)              # [nW, N]
    mw = ids[part]                                         # region id per window token
    diff = mw[:, None, :] - mw[:, :, None]
    return np.where(diff != 0, np.float32(-100.0), np.float32(0.0)).astype(np.float32)


def relative_position_index(ws_h: int, ws_w: int) -> np.ndarray:
    """``relative_position_index`` buffer, int64 ``[N, N]``.
    models/swin_transformer_v2.py:249-259:
    idx[a, b] = (ya - yb + ws_h - 1) * (2*ws_w - 1) + (xa - xb + ws_w - 1)."""
    y = np.repeat(np.arange(ws_h), ws_w)
    x = np.tile(np.arange(ws_w), ws_h)
    dy = y[:, None] - y[None, :] + ws_h - 1
    dx = x[:, None] - x[None, :] + ws_w - 1
    return (dy * (2 * ws_w - 1) + dx).astype(np.int64)


def relative_coords_table(ws_h: int, ws_w: int, pretrain_ws: int) -> np.ndarray:
    """``relative_coords_table`` buffer for ``norm8_log_bylayer``: float32
    ``[1, 2*ws_h-1, 2*ws_w-1, 2]``.  models/swin_transformer_v2.py:190-194, 233-239.

    The reference computes it with float32 torch ops; this restatement follows the same
    op order in float32 (``/ (pre-1)``, ``* 8``, ``sign * log2(|.|+1) / log2(8)``) so the
    buffer matches to the last ulp on the values tested in tests/test_oracle_golden.py."""
    h = np.arange(-(ws_h - 1), ws_h, dtype=np.float32)
    w = np.arange(-(ws_w - 1), ws_w, dtype=np.float32)
    t = np.stack(np.meshgrid(h, w, indexing="ij"), axis=-1)[None].astype(np.float32)
    t = t / np.float32(pretrain_ws - 1)
    t = t * np.float32(8)
    t = np.sign(t) * np.log2(np.abs(t) + np.float32(1.0)) / np.float32(np.log2(8))
    return t.astype(np.float32)
