// Window geometry shared by the standalone window kernels and the attention kernels:
// pad (right/bottom, to multiples of ws) -> cyclic roll by -shift -> window partition, as address math.
// Reference: models/swin_transformer_v2.py:120-147, 429-463.
#pragma once
#include "common.cuh"

namespace b200swin {

struct WinGeom {
  int B, H, W, Hp, Wp, ws, shift, nWh, nWw;
  int64_t row_vecs;  // used by the copy kernels only
};

inline void make_geom(WinGeom* g, int B, int H, int W, int ws, int shift) {
  g->B = B; g->H = H; g->W = W; g->ws = ws; g->shift = shift;
  g->Hp = (H + ws - 1) / ws * ws;
  g->Wp = (W + ws - 1) / ws * ws;
  g->nWh = g->Hp / ws;
  g->nWw = g->Wp / ws;
  g->row_vecs = 0;
}

// (window index over B*nW, in-window token) -> source token (b, i, j) on the UNPADDED grid.
// Returns false when the slot is a zero pad token.
__device__ __forceinline__ bool win_token(const WinGeom& g, int64_t win, int tok, int& b, int& i, int& j,
                                          int& si, int& sj) {
  const int nW = g.nWh * g.nWw;
  b = (int)(win / nW);
  int w = (int)(win - (int64_t)b * nW);
  int wh = w / g.nWw, ww = w - wh * g.nWw;
  int r = tok / g.ws, c = tok - r * g.ws;
  si = wh * g.ws + r;                               // coordinates on the shifted padded grid
  sj = ww * g.ws + c;
  i = si + g.shift; if (i >= g.Hp) i -= g.Hp;       // shifted[si] = x[(si + shift) mod Hp]
  j = sj + g.shift; if (j >= g.Wp) j -= g.Wp;
  return i < g.H && j < g.W;
}

__device__ __forceinline__ bool slot_to_token(const WinGeom& g, int64_t slot, int& b, int& i, int& j) {
  const int N = g.ws * g.ws;
  int64_t win = slot / N;
  int tok = (int)(slot - win * N);
  int si, sj;
  return win_token(g, win, tok, b, i, j, si, sj);
}

}  // namespace b200swin
