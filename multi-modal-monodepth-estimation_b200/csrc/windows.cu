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

// 2x2 patch merging as an index map (reference PatchMerging.forward, models/swin_transformer_v2.py:660-672: zero-pad H, W
// to even, cat([x[:,0::2,0::2], x[:,1::2,0::2], x[:,0::2,1::2], x[:,1::2,1::2]], -1)):
//   merged[b, i2*W2 + j2, k*C + c] = x[b, 2*i2 + (k & 1), 2*j2 + (k >> 1), c]      (zero beyond H, W)
// kGather: x -> merged;  !kGather: the adjoint (gradient of merged -> gradient of x; pad positions are dropped).
// One 16-byte vector per thread, indexed over the x-shaped tensor for the scatter and the merged one for the gather,
// so every store is coalesced and written exactly once.
template <bool kGather>
__global__ void __launch_bounds__(256)
patch_merge_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int B, int H, int W, int H2, int W2,
                   int row_vecs /* 16-byte vectors per C */, uint32_t total) {
  // 32-bit index arithmetic (the host checks total < 2^31): a 64-bit division is ~100 instructions per vector
  const uint32_t rv = (uint32_t)row_vecs;
  for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
    if (kGather) {
      // v over merged [B, H2, W2, 4, row_vecs]
      uint32_t t = v / rv;
      const uint32_t cv = v - t * rv;
      const int k = (int)(t & 3);
      t >>= 2;
      uint32_t t2 = t / (uint32_t)W2;
      const int j2 = (int)(t - t2 * (uint32_t)W2);
      const uint32_t b = t2 / (uint32_t)H2;
      const int i2 = (int)(t2 - b * (uint32_t)H2);
      const int i = 2 * i2 + (k & 1), j = 2 * j2 + (k >> 1);
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (i < H && j < W) val = src[(((int64_t)b * H + i) * W + j) * row_vecs + cv];
      dst[v] = val;
    } else {
      // v over x [B, H, W, row_vecs]
      uint32_t t = v / rv;
      const uint32_t cv = v - t * rv;
      uint32_t t2 = t / (uint32_t)W;
      const int j = (int)(t - t2 * (uint32_t)W);
      const uint32_t b = t2 / (uint32_t)H;
      const int i = (int)(t2 - b * (uint32_t)H);
      const int k = (i & 1) + 2 * (j & 1);
      dst[v] = src[((((int64_t)b * H2 + (i >> 1)) * W2 + (j >> 1)) * 4 + k) * row_vecs + cv];
    }
  }
}

// Patchify for the stride = kernel patch-embedding conv (reference PatchEmbed.forward, models/swin_transformer_v2.py:
// 941-957): the conv IS a GEMM over non-overlapping patches,
//   cols[(b, i, j), c*ph*pw + kh*pw + kw] = x[b, c, i*ph + kh, j*pw + kw]       (zero beyond H, W: the reference pads)
// with the conv weight [E, Cin, ph, pw] viewed as [E, Cin*ph*pw].  One thread per (patch, c, kh): pw contiguous pixels.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
patchify_kernel(const TI* __restrict__ x, TO* __restrict__ cols, int B, int Cin, int H, int W, int ph, int pw, int Hp,
                int Wp, int64_t total) {
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads -> consecutive patches j of the same (b, i, c, kh): coalesced reads of one image row
    int j = (int)(v % Wp);
    int64_t t = v / Wp;
    const int kh = (int)(t % ph);
    t /= ph;
    const int c = (int)(t % Cin);
    t /= Cin;
    const int i = (int)(t % Hp);
    const int b = (int)(t / Hp);
    const int y = i * ph + kh;
    const TI* src = x + (((int64_t)b * Cin + c) * H + y) * W + (int64_t)j * pw;
    TO* dst = cols + ((((int64_t)b * Hp + i) * Wp + j) * Cin + c) * (ph * pw) + kh * pw;
    for (int kw = 0; kw < pw; ++kw) {
      const float val = (y < H && j * pw + kw < W) ? Io<TI>::ld(src + kw) : 0.f;
      Io<TO>::st(dst + kw, val);
    }
  }
}

// The same map with one CTA per row of patches (b, i): the Cin*ph image rows it covers are read coalesced into shared
// memory and the Wp patches, contiguous in `cols`, are written coalesced (the per-thread version above writes pw elements
// at a stride of a whole patch: 4x the time on the 480x480 frames).  Row stride padded by pw: conflict-free reads for pw = 4.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
patchify_row_kernel(const TI* __restrict__ x, TO* __restrict__ cols, int Cin, int H, int W, int ph, int pw, int Hp, int Wp) {
  extern __shared__ float tile[];   // [Cin*ph][Wp*pw + pw]
  const int b = blockIdx.x / Hp, i = blockIdx.x - b * Hp;
  const int wpad = Wp * pw, stride = wpad + pw, rows = Cin * ph;
  for (int e = threadIdx.x; e < rows * wpad; e += 256) {
    const int r = e / wpad, col = e - r * wpad;
    const int c = r / ph, kh = r - c * ph, y = i * ph + kh;
    tile[r * stride + col] = (y < H && col < W) ? Io<TI>::ld(x + (((int64_t)b * Cin + c) * H + y) * W + col) : 0.f;
  }
  __syncthreads();
  const int K = rows * pw;          // elements of one patch: (c, kh, kw)
  TO* dst = cols + (int64_t)blockIdx.x * Wp * K;
  for (int o = threadIdx.x; o < Wp * K; o += 256) {
    const int j = o / K, q = o - j * K;
    const int r = q / pw, kw = q - r * pw;
    Io<TO>::st(dst + o, tile[r * stride + j * pw + kw]);
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

extern "C" int b200swin_patch_merge(const void* in, void* out, int B, int H, int W, int C, int elem_bytes, int backward,
                                    void* stream) {
  BSW_REQUIRE(in && out, "patch_merge: null pointer");
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "patch_merge: non-positive dimension");
  BSW_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "patch_merge: elem_bytes %d", elem_bytes);
  const int64_t row_bytes = (int64_t)C * elem_bytes;
  BSW_REQUIRE(row_bytes % 16 == 0, "patch_merge: C * elem_bytes = %lld must be a multiple of 16", (long long)row_bytes);
  BSW_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "patch_merge: pointers must be 16-byte aligned");
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const int row_vecs = (int)(row_bytes / 16);
  const int64_t total = backward ? (int64_t)B * H * W * row_vecs : (int64_t)B * H2 * W2 * 4 * row_vecs;
  BSW_REQUIRE(total < (1ll << 31), "patch_merge: tensor too large for 32-bit vector indices");
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(blocks < cap ? blocks : cap);
  cudaStream_t st = (cudaStream_t)stream;
  if (backward)
    patch_merge_kernel<false><<<grid, 256, 0, st>>>((const uint4*)in, (uint4*)out, B, H, W, H2, W2, row_vecs, (uint32_t)total);
  else
    patch_merge_kernel<true><<<grid, 256, 0, st>>>((const uint4*)in, (uint4*)out, B, H, W, H2, W2, row_vecs, (uint32_t)total);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

extern "C" int b200swin_patchify(const void* x, int x_dtype, void* cols, int cols_dtype, int B, int Cin, int H, int W,
                                 int ph, int pw, void* stream) {
  BSW_REQUIRE(x && cols, "patchify: null pointer");
  BSW_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && ph > 0 && pw > 0, "patchify: non-positive dimension");
  BSW_REQUIRE((x_dtype == B200SWIN_F32 || x_dtype == B200SWIN_BF16) && (cols_dtype == B200SWIN_F32 || cols_dtype == B200SWIN_BF16),
              "patchify: bad dtype");
  const int Hp = (H + ph - 1) / ph, Wp = (W + pw - 1) / pw;
  const int64_t total = (int64_t)B * Hp * Cin * ph * Wp;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(blocks < cap ? blocks : cap);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tile_bytes = (int64_t)Cin * ph * ((int64_t)Wp * pw + pw) * 4;
  const bool by_row = tile_bytes <= 48 * 1024 && (int64_t)B * Hp < (1ll << 31) && (int64_t)Cin * ph * Wp * pw < (1ll << 30);
#define PF(TI, TO)                                                                                                   \
  if (by_row)                                                                                                        \
    patchify_row_kernel<TI, TO><<<B * Hp, 256, (size_t)tile_bytes, st>>>((const TI*)x, (TO*)cols, Cin, H, W, ph, pw, Hp, Wp); \
  else                                                                                                               \
    patchify_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)x, (TO*)cols, B, Cin, H, W, ph, pw, Hp, Wp, total)
  if (x_dtype == B200SWIN_F32 && cols_dtype == B200SWIN_BF16) { PF(float, __nv_bfloat16); }
  else if (x_dtype == B200SWIN_F32) { PF(float, float); }
  else if (cols_dtype == B200SWIN_BF16) { PF(__nv_bfloat16, __nv_bfloat16); }
  else { PF(__nv_bfloat16, float); }
#undef PF
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
