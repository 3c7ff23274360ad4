// Window plumbing as pure index maps: pad + cyclic roll + window_partition folded into ONE gather,
// and window_reverse + roll back + crop folded into ONE scatter (its exact adjoint).
// Replaces models/swin_transformer_v2.py:120-147 (partition/reverse), :438/:458 (torch.roll),
// :429-434/:462-463 (pad/crop) and :874-892 (shift mask).  Integer work only -> bit exact.
//
// These standalone kernels back the drop-in `window_partition` / `window_reverse` functions and the
// index-map parity tests.  Inside the attention block the same address math is folded into the
// attention kernel's loads and stores, so no permuted copy touches HBM there.
#include "common.cuh"
#include "wingeom.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

template <typename V, bool kGather>
__global__ void __launch_bounds__(256)
window_move_kernel(const V* __restrict__ src, V* __restrict__ dst, WinGeom g, int64_t total_vecs) {
  // one thread per vector of the window-major tensor; gather: dst is window-major, scatter: src is.
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total_vecs;
       v += (int64_t)gridDim.x * blockDim.x) {
    int64_t slot = v / g.row_vecs;
    int64_t k = v - slot * g.row_vecs;
    int b, i, j;
    bool real = slot_to_token(g, slot, b, i, j);
    int64_t tok_off = (((int64_t)b * g.H + i) * g.W + j) * g.row_vecs + k;
    if (kGather) {
      V val;
      if (real) val = src[tok_off];
      else memset(&val, 0, sizeof(V));
      dst[v] = val;
    } else {
      if (real) dst[tok_off] = src[v];
    }
  }
}

__global__ void __launch_bounds__(256)
shift_mask_kernel(float* __restrict__ out, int Hp, int Wp, int ws, int shift, int nWw, int64_t total) {
  const int N = ws * ws;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    int bcol = (int)(e % N);
    int64_t t = e / N;
    int a = (int)(t % N);
    int w = (int)(t / N);
    int wh = w / nWw, ww = w - wh * nWw;
    int ia = wh * ws + a / ws, ja = ww * ws + a % ws;
    int ib = wh * ws + bcol / ws, jb = ww * ws + bcol % ws;
    int ida = 3 * region_1d(ia, Hp, ws, shift) + region_1d(ja, Wp, ws, shift);
    int idb = 3 * region_1d(ib, Hp, ws, shift) + region_1d(jb, Wp, ws, shift);
    out[e] = (ida != idb) ? -100.0f : 0.0f;
  }
}

static int check_geom(int B, int H, int W, int C, int ws, int shift, int elem_bytes, WinGeom* g, int* vec_bytes,
                      const void* p0, const void* p1) {
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && ws > 0, "window op: non-positive dimension");
  BSW_REQUIRE(shift >= 0 && shift < ws, "window op: shift %d must be in [0, ws=%d)", shift, ws);
  BSW_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "window op: elem_bytes %d", elem_bytes);
  int64_t row_bytes = (int64_t)C * elem_bytes;
  BSW_REQUIRE(row_bytes % 4 == 0, "window op: C*elem_bytes must be a multiple of 4");
  make_geom(g, B, H, W, ws, shift);
  bool al16 = ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1)) & 15) == 0;
  *vec_bytes = (row_bytes % 16 == 0 && al16) ? 16 : 4;
  g->row_vecs = row_bytes / *vec_bytes;
  return B200SWIN_OK;
}

template <bool kGather>
static int window_move(const void* src, void* dst, int B, int H, int W, int C, int ws, int shift, int elem_bytes,
                       void* stream) {
  BSW_REQUIRE(src && dst, "window op: null pointer");
  WinGeom g;
  int vb;
  int rc = check_geom(B, H, W, C, ws, shift, elem_bytes, &g, &vb, src, dst);
  if (rc) return rc;
  int64_t total = (int64_t)B * g.Hp * g.Wp * g.row_vecs;
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 16;
  int grid = (int)(blocks < cap ? blocks : cap);
  cudaStream_t st = (cudaStream_t)stream;
  if (vb == 16)
    window_move_kernel<uint4, kGather><<<grid, 256, 0, st>>>((const uint4*)src, (uint4*)dst, g, total);
  else
    window_move_kernel<uint32_t, kGather><<<grid, 256, 0, st>>>((const uint32_t*)src, (uint32_t*)dst, g, total);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_window_gather(const void* x, void* out, int B, int H, int W, int C, int ws, int shift,
                                      int elem_bytes, void* stream) {
  return window_move<true>(x, out, B, H, W, C, ws, shift, elem_bytes, stream);
}

extern "C" int b200swin_window_scatter(const void* win, void* out, int B, int H, int W, int C, int ws, int shift,
                                       int elem_bytes, void* stream) {
  return window_move<false>(win, out, B, H, W, C, ws, shift, elem_bytes, stream);
}

extern "C" int b200swin_shift_mask(float* out, int H, int W, int ws, int shift, void* stream) {
  BSW_REQUIRE(out, "shift_mask: null pointer");
  BSW_REQUIRE(H > 0 && W > 0 && ws > 0 && shift > 0 && shift < ws, "shift_mask: need 0 < shift < ws");
  int Hp = (H + ws - 1) / ws * ws, Wp = (W + ws - 1) / ws * ws;
  int nWw = Wp / ws;
  int64_t N = (int64_t)ws * ws;
  int64_t total = (int64_t)(Hp / ws) * nWw * N * N;
  int64_t blocks = (total + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 16;
  shift_mask_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(out, Hp, Wp, ws, shift,
                                                                                       nWw, total);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
