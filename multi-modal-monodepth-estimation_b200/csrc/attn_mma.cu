// Register-resident attention core for the small windows (bf16 storage): one warp per 16 query rows, the whole
// softmax(QK^T) V chain of those rows in registers (mma.sync m16n8k16 -> HMMA), no TMEM round trips, no MMA-issuer thread.
//
// Why not tcgen05 here.  With head_dim 32 the contractions of a 12x12 window are ~400 cycles of tcgen05 work per (window,
// head) against >= 1300 cycles of exponentials: the tensor pipe is idle either way, and what the single-tile tcgen05
// kernels (attn_fwd_ws.cu / attn_bwd_ws.cu) pay for is the 128-lane accumulator tile -- a 144-row window is 128 + 16
// rows, the 16-row tail costs a full pass of latency on one TMEM lane quarter -- plus a TMEM load / pack / store and an
// mbarrier round trip per logit block.  144 rows are exactly nine 16-row tiles of the warp-level MMA, whose accumulator
// layout IS the A-operand layout of the next contraction (P goes from the softmax straight into P V).  The legacy
// tensor path sustains 1890 FLOP/clk/SM on B200 (tools/probe/hmma_rate.cu), ~4x what this kernel needs.
// The GEMMs and the large windows (KV-blocked kernels, attn_flash.cu) stay on tcgen05.
//
// Same math as the other attention kernels: models/swin_transformer_v2.py:295-328 with the pad / roll / partition /
// reverse / crop of :429-463 and the shift mask of :874-892 as address math; backward per SURVEY.md appendix A.
#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include "common.cuh"
#include "wingeom.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

namespace {
constexpr int HD = 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMaskLog2 = -100.0f * 1.4426950408889634f;
constexpr float kLazy = 8.0f;

template <int WS>
struct MCfg {
  static constexpr int N = WS * WS;
  static constexpr int NT = (N + 15) / 16;           // 16-row tiles = warps of a CTA
  static constexpr int NP = NT * 16;                 // rows / keys incl. the zero rows of a ragged last tile
  static constexpr bool RAGGED = NP != N;
  static constexpr int TW = 2 * WS - 1, NTAB = TW * TW;
  static constexpr int THREADS = NT * 32;
  static constexpr uint32_t TILE = NP * 64;          // one operand tile: NP rows of 64 bytes, 64B-swizzled
  static constexpr int NTILES8 = NP / 8;             // 8-key column tiles of the logits
  static constexpr int CHN = (NTILES8 % 6 == 0) ? 6 : ((NTILES8 % 4 == 0) ? 4 : 2);   // column tiles per softmax chunk
  static constexpr int NCH = NTILES8 / CHN;
  // resident CTAs per SM the kernels are compiled for: ~18 warps forward (<= 96 registers), ~12 backward (<= 168)
  static constexpr int FWD_CTAS = NT >= 9 ? 2 : (18 / NT > 8 ? 8 : 18 / NT);
  static constexpr int BWD_CTAS = NT >= 9 ? 1 : (12 / NT > 8 ? 8 : 12 / NT);
};

struct MmaArgs {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* dout;
  __nv_bfloat16* out;
  __nv_bfloat16* out_lo;
  __nv_bfloat16* dqkv;
  float* lse;
  const float* dvec;
  const float* inv_norm;
  const float* table16;
  const float* scale;
  const float* qpad;
  const float* vpad;
  float* dtable16;
  float* dscale;
  float* dvpad;
  float* dcol;                // optional [3C]: column sums of dq (first C) and dv (last C) = q_bias / v_bias gradients
  WinGeom g;
  int C, nH;
  int64_t nwin, nitems;       // item = head * nwin + window
};

__device__ __forceinline__ uint32_t sw64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// 16 bytes of one row of an operand tile: global (a real token), a pad value (fp32 -> bf16) or zeros
__device__ __forceinline__ void put16(unsigned char* tile, uint32_t tile_s, uint32_t off, const __nv_bfloat16* src,
                                      const float* padv) {
  if (src) {
    ptx::cp_async_16(tile_s + off, src);
  } else {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (padv) v = make_uint4(pack2(padv[0], padv[1]), pack2(padv[2], padv[3]), pack2(padv[4], padv[5]), pack2(padv[6], padv[7]));
    *reinterpret_cast<uint4*>(tile + off) = v;
  }
}

// Byte offset 4 (y TW + x) into the bias table of in-window token q0 + 2 tq, q0 a compile-time multiple of 8 (even windows:
// the pair (q, q + 1) shares a window row, the second entry is the first + 4 bytes).  Replaces a shared-memory table
// read per column tile: the kernels are bound by shared-memory wavefronts, not by integer issue slots.
template <int WS>
__device__ __forceinline__ int kof_pair(int q0, int tq) {
  constexpr int TW = 2 * WS - 1;
  int y = q0 / WS, x = q0 % WS + 2 * tq;          // q0 / WS and q0 % WS fold at compile time when q0 does
  if (x >= WS) { x -= WS; ++y; }
  return 4 * (y * TW + x);
}

// window of an item and the source token of in-window row r: >= 0 flat token, -1 pad token, -2 row beyond the window
struct ItemPos {
  int h, b, wh, ww;
  int64_t win;
};
__device__ __forceinline__ ItemPos item_pos(const MmaArgs& a, int64_t item) {
  ItemPos p;
  p.h = (int)(item / a.nwin);
  p.win = item - (int64_t)p.h * a.nwin;
  const int nW = a.g.nWh * a.g.nWw;
  p.b = (int)(p.win / nW);
  const int w = (int)(p.win - (int64_t)p.b * nW);
  p.wh = w / a.g.nWw;
  p.ww = w - p.wh * a.g.nWw;
  return p;
}
// the next item of a CTA's contiguous head-major range, without the 64-bit divisions of item_pos
__device__ __forceinline__ void item_next(const MmaArgs& a, ItemPos& p) {
  ++p.win;
  if (++p.ww == a.g.nWw) {
    p.ww = 0;
    if (++p.wh == a.g.nWh) { p.wh = 0; ++p.b; }
  }
  if (p.win == a.nwin) { p.win = 0; p.b = 0; ++p.h; }
}
// in-window row (y, x) of the item's window -> source token (>= 0 flat token, -1 pad token) and shift-mask region
template <int WS>
__device__ __forceinline__ int row_token(const WinGeom& g, const ItemPos& p, int y, int x, int* region) {
  const int si = p.wh * WS + y, sj = p.ww * WS + x;
  int i = si + g.shift; if (i >= g.Hp) i -= g.Hp;
  int j = sj + g.shift; if (j >= g.Wp) j -= g.Wp;
  *region = g.shift > 0 ? 3 * region_1d(si, g.Hp, WS, g.shift) + region_1d(sj, g.Wp, WS, g.shift) : 0;
  return (i < g.H && j < g.W) ? (p.b * g.H + i) * g.W + j : -1;
}

// ------------------------------------------------------------------------------------------------------- forward
// CTA = NT warps = one (window, head) item at a time, a contiguous head-major range of items per CTA; the q / k / v tiles
// of the next two items are in flight (three stages) while the warps work on the current one, one CTA barrier per item.
// Two CTAs per SM.
// Staging of a window's tiles: a window that does not wrap around the map under the cyclic shift is a ws x ws x 32-column
// BOX of the [B, H, W, 3C] qkv tensor: ONE elected thread issues three 4-D TMA loads (cp.async.bulk.tensor, SWIZZLE_64B =
// the tile layout sw64() of the ldmatrix reads, completing on the stage's mbarrier); rows beyond H / W arrive as the
// hardware's zero fill, which IS the k of a pad token, and the threads owning pad rows then write q = q_bias-hat and
// v = v_bias over them.  Windows on the roll seam (last window row / column of a shifted block: up to four rectangles)
// and the ragged window sizes keep the per-row 16-byte cp.async gather.
constexpr int kFwdStages = 3;

template <int WS>
__global__ void __launch_bounds__(MCfg<WS>::THREADS, MCfg<WS>::FWD_CTAS)
attn_mma_fwd_kernel(const __grid_constant__ MmaArgs a, const __grid_constant__ CUtensorMap tmap, const int use_tma) {
  using Cf = MCfg<WS>;
  constexpr int N = Cf::N, NP = Cf::NP, TW = Cf::TW, CHN = Cf::CHN;
  constexpr uint32_t TILE = Cf::TILE, STAGE = 3 * TILE;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;  // 512-byte period of the TMA swizzle pattern
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  unsigned char* tiles = sm;                                              // [stages][q | k | v]
  int* tokm = reinterpret_cast<int*>(tiles + kFwdStages * STAGE);         // [stages][NP] source token of a row
  int* ridm = tokm + kFwdStages * NP;                                     // [stages][NP] shift-mask region id
  int* kofk = ridm + kFwdStages * NP;                                     // [NP] byte offset 4 (y TW + x) of a window row
  float* tab = reinterpret_cast<float*>(kofk + NP);                       // [NTAB] bias table of the head, log2 units, minus the softmax offset
  float* tred = tab + Cf::NTAB;                                           // [2 NT] per-warp max / min of the table
  uint64_t* full = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(tred + 2 * Cf::NT) + 7) & ~uintptr_t(7));   // [stages] "TMA tiles landed"
  const uint32_t tiles_s = base_u32;

  const WinGeom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int64_t per = a.nitems / gridDim.x, rem = a.nitems % gridDim.x;
  const int64_t it0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int nit = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));
  const int C3 = 3 * a.C;
  const bool tma_on = !Cf::RAGGED && use_tma != 0;
  // does the item's window lie in one piece on the map (no wrap under the roll)?  CTA-uniform
  auto in_one_piece = [&](const ItemPos& q) { return tma_on && (g.shift == 0 || (q.wh < g.nWh - 1 && q.ww < g.nWw - 1)); };

  for (int r = tid; r < NP; r += Cf::THREADS) kofk[r] = r < N ? 4 * ((r / WS) * TW + (r % WS)) : 0;
  if (tma_on) {
    if (tid == 0) {
      ptx::prefetch_tmap(&tmap);
      for (int s_ = 0; s_ < kFwdStages; ++s_) ptx::mbar_init(&full[s_], 1);
      ptx::fence_mbar_init();
    }
    __syncthreads();
  }

  // a thread copies the same half row (32 bytes of q, k and v) of every item: row tid / 2, dims 16 (tid & 1) ...
  static_assert(NP * 2 == Cf::THREADS, "two threads per row");
  const int prow = tid >> 1, py = prow / WS, px = prow - py * WS;
  const uint32_t poff = sw64(prow, (tid & 1) * 2);               // second 16 bytes: poff ^ 16
  ItemPos pp = item_pos(a, it0);                 // cursor of the prefetch stream
  auto prefetch = [&](int i) {
    const int stage = i % kFwdStages;
    unsigned char* q0 = tiles + (size_t)stage * STAGE;
    const uint32_t q0_s = tiles_s + (uint32_t)stage * STAGE;
    int region = 0, t = -2;
    if (!Cf::RAGGED || prow < N) t = row_token<WS>(g, pp, py, px, &region);
    if ((tid & 1) == 0) {
      tokm[stage * NP + prow] = t;
      ridm[stage * NP + prow] = region;
    }
    if (in_one_piece(pp)) {
      if (tid == 0) {
        ptx::fence_proxy_async_smem();            // the stage was last written / read through the generic proxy
        ptx::mbar_arrive_expect_tx(&full[stage], 3u * TILE);
#pragma unroll
        for (int k = 0; k < 3; ++k)
          ptx::tma_load_4d(q0 + k * TILE, &tmap, &full[stage], k * a.C + pp.h * HD, pp.ww * WS + g.shift,
                           pp.wh * WS + g.shift, pp.b);
      }
    } else {
      if (tma_on && tid == 0) ptx::mbar_arrive(&full[stage]);      // a gathered item: the phase completes at once
      if (t >= 0) {
        const __nv_bfloat16* src = a.qkv + (int64_t)t * C3 + pp.h * HD + (tid & 1) * 16;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          ptx::cp_async_16(q0_s + k * TILE + poff, src + k * a.C);
          ptx::cp_async_16(q0_s + k * TILE + (poff ^ 16u), src + k * a.C + 8);
        }
      } else {
        const float* qp = (t == -1 && a.qpad) ? a.qpad + pp.h * HD + (tid & 1) * 16 : nullptr;
        const float* vp = (t == -1 && a.vpad) ? a.vpad + pp.h * HD + (tid & 1) * 16 : nullptr;
        put16(q0, q0_s, poff, nullptr, qp);
        put16(q0, q0_s, poff ^ 16u, nullptr, qp ? qp + 8 : nullptr);
        put16(q0 + TILE, q0_s + TILE, poff, nullptr, nullptr);
        put16(q0 + TILE, q0_s + TILE, poff ^ 16u, nullptr, nullptr);
        put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff, nullptr, vp);
        put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff ^ 16u, nullptr, vp ? vp + 8 : nullptr);
      }
    }
    item_next(a, pp);
  };

  if (nit > 0) prefetch(0);
  ptx::cp_async_commit();
  if (nit > 1) prefetch(1);
  ptx::cp_async_commit();

  int cur_h = -1;
  float scale2 = 0.f, boff = 0.f;
  bool fixed_off = false;
  const int rA = warp * 16 + gq, rB = rA + 8;
  ItemPos p = item_pos(a, it0);                  // cursor of the compute stream
  // lane-constant parts of the ldmatrix addresses (the swizzle term only sees the low row bits)
  const uint32_t lq_off = sw64(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);      // q rows, dims 0-15 (+32 B: 16-31)
  const uint32_t lk_off = sw64(lane & 7, lane >> 3);                                            // 8 keys x 4 dim chunks
  const uint32_t lv_off0 = sw64((lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);                 // 16 keys x dim chunks 0, 1
  const uint32_t lv_off2 = sw64((lane & 7) + ((lane >> 3) & 1) * 8, 2 + (lane >> 4));           //           dim chunks 2, 3

#pragma unroll 1
  for (int i = 0; i < nit; ++i, item_next(a, p)) {
    const int stage = i % kFwdStages;
    ptx::cp_async_wait<1>();                      // this thread's copies of item i have landed
    if (tma_on) {
      // ... or the item's three TMA boxes.  Every use of a stage completes one phase of its mbarrier (gathered items by
      // a plain arrive), so the parity is that of the use count
      ptx::mbar_wait(&full[stage], (uint32_t)(i / kFwdStages) & 1u);
      // pad tokens of a window that overhangs the map: the zero fill is their k; q and v are the bias rows
      if (in_one_piece(p) && (p.wh * WS + g.shift + WS > g.H || p.ww * WS + g.shift + WS > g.W)) {
        if (tokm[stage * NP + prow] == -1) {
          unsigned char* q0 = tiles + (size_t)stage * STAGE;
          const uint32_t q0_s = tiles_s + (uint32_t)stage * STAGE;
          const float* qp = a.qpad ? a.qpad + p.h * HD + (tid & 1) * 16 : nullptr;
          const float* vp = a.vpad ? a.vpad + p.h * HD + (tid & 1) * 16 : nullptr;
          if (qp) { put16(q0, q0_s, poff, nullptr, qp); put16(q0, q0_s, poff ^ 16u, nullptr, qp + 8); }
          if (vp) {
            put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff, nullptr, vp);
            put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff ^ 16u, nullptr, vp + 8);
          }
        }
      }
    }
    __syncthreads();                              // ... everybody's; and every warp is done with item i - 1
    if (i + 2 < nit) prefetch(i + 2);             // into the stage item i - 1 has just released
    ptx::cp_async_commit();
    if (p.h != cur_h) {                           // CTA-uniform
      // Bias table of the head and the softmax offset.  Logits are cos * scale + bias with |cos| <= 1 (+ bf16 rounding):
      // when 2 scale + (range of the table) stays below 2^100 in the exponent, a FIXED offset scale + max(bias) makes
      // every P <= ~1 and no row sum can underflow -- the row maximum, the rescale and the subtraction per logit
      // disappear.  Heads whose temperature is too large for that (the clamp allows 100) take the online-softmax path.
      float tmx = -INFINITY, tmn = INFINITY;
      for (int t = tid; t < Cf::NTAB; t += Cf::THREADS) {
        const float v = a.table16[(int64_t)t * a.nH + p.h] * kLog2e;
        tmx = fmaxf(tmx, v);
        tmn = fminf(tmn, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        tmx = fmaxf(tmx, __shfl_xor_sync(0xffffffffu, tmx, o));
        tmn = fminf(tmn, __shfl_xor_sync(0xffffffffu, tmn, o));
      }
      if (lane == 0) { tred[2 * warp] = tmx; tred[2 * warp + 1] = tmn; }
      __syncthreads();
      for (int w = 0; w < Cf::NT; ++w) { tmx = fmaxf(tmx, tred[2 * w]); tmn = fminf(tmn, tred[2 * w + 1]); }
      scale2 = a.scale[p.h] * kLog2e;
      fixed_off = 2.02f * scale2 + (tmx - tmn) < 100.f;
      boff = fixed_off ? 1.01f * scale2 + tmx : 0.f;
      for (int t = tid; t < Cf::NTAB; t += Cf::THREADS) tab[t] = a.table16[(int64_t)t * a.nH + p.h] * kLog2e - boff;
      cur_h = p.h;
      __syncthreads();
    }
    const bool need_mask = g.shift > 0 && (p.wh == g.nWh - 1 || p.ww == g.nWw - 1);
    const uint32_t q_s = tiles_s + (uint32_t)stage * STAGE, k_s = q_s + TILE, v_s = k_s + TILE;
    const int* tokS = tokm + stage * NP;
    const int* ridS = ridm + stage * NP;
    const uint32_t tab_s = ptx::smem_u32(tab) + 4u * (uint32_t)((WS - 1) * (TW + 1));
    const uint32_t tabA = tab_s + (uint32_t)kofk[rA], tabB = tab_s + (uint32_t)kofk[rB];
    const int ridA = ridS[rA], ridB = ridS[rB];

    // A fragments of the warp's 16 query rows (two k-steps of 16 dims)
    uint32_t qa[2][4];
    ldsm4(qa[0], q_s + lq_off);
    ldsm4(qa[1], q_s + (lq_off ^ 32u));           // dim chunks 2, 3: bit 1 of the (swizzled) chunk index
    // a tile of 16 pad queries (right / bottom edge of a padded map) has no output: nothing to do.  Its rows get
    // lse = +inf so that any backward sees P = 0 for them.
    if (__all_sync(0xffffffffu, tokS[rA] < 0 && tokS[rB] < 0)) {
      float* lse_pad = a.lse + (p.win * a.nH + p.h) * N;
      if (tq == 0) {
        if (rA < N) lse_pad[rA] = INFINITY;
        if (rB < N) lse_pad[rB] = INFINITY;
      }
      continue;
    }
    float o[4][4], ol[4] = {0.f, 0.f, 0.f, 0.f};  // O and, as a fifth column tile against a ones operand, the row sums of P
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    float mA = fixed_off ? 0.f : -INFINITY, mB = mA;           // reference maximum of the rows, relative to boff
    constexpr uint32_t kOnes = 0x3f803f80u;       // bf16 (1, 1)

    auto chunk = [&](int c, auto mask_c, auto fixed_c) {
      constexpr bool MASK = decltype(mask_c)::value, FIXED = decltype(fixed_c)::value;
      const int key0 = c * CHN * 8;
      float s[CHN][4];
#pragma unroll
      for (int n = 0; n < CHN; ++n) {
        s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
        uint32_t kb[4];
        ldsm4(kb, k_s + (uint32_t)(key0 + n * 8) * 64u + lk_off);
        mma16816(s[n], qa[0], kb[0], kb[1]);
        mma16816(s[n], qa[1], kb[2], kb[3]);
      }
      float mxA = -INFINITY, mxB = -INFINITY;
#pragma unroll
      for (int n = 0; n < CHN; ++n) {
        const int kcol = key0 + n * 8 + 2 * tq;
        const int2 kk = *reinterpret_cast<const int2*>(kofk + kcol);
        s[n][0] = fmaf(s[n][0], scale2, lds32(tabA - (uint32_t)kk.x));
        s[n][1] = fmaf(s[n][1], scale2, lds32(tabA - (uint32_t)kk.y));
        s[n][2] = fmaf(s[n][2], scale2, lds32(tabB - (uint32_t)kk.x));
        s[n][3] = fmaf(s[n][3], scale2, lds32(tabB - (uint32_t)kk.y));
        if (MASK) {
          const int2 rr = *reinterpret_cast<const int2*>(ridS + kcol);
          if (rr.x != ridA) s[n][0] += kMaskLog2;
          if (rr.y != ridA) s[n][1] += kMaskLog2;
          if (rr.x != ridB) s[n][2] += kMaskLog2;
          if (rr.y != ridB) s[n][3] += kMaskLog2;
        }
        if (Cf::RAGGED) {
          if (kcol >= N) s[n][0] = s[n][2] = -INFINITY;
          if (kcol + 1 >= N) s[n][1] = s[n][3] = -INFINITY;
        }
        if (!FIXED) {
          mxA = fmaxf(mxA, fmaxf(s[n][0], s[n][1]));
          mxB = fmaxf(mxB, fmaxf(s[n][2], s[n][3]));
        }
      }
      if (!FIXED) {
        // lazy rescale: the reference maximum of a row only moves when a chunk exceeds it by more than 2^8 (P <= 256 is
        // harmless in fp32 / bf16), so after the first chunk the accumulators are almost never touched
        mxA = quad_max(mxA);
        mxB = quad_max(mxB);
        const float mnA = mxA > mA + kLazy ? mxA : mA, mnB = mxB > mB + kLazy ? mxB : mB;
        if (__any_sync(0xffffffffu, mnA != mA || mnB != mB)) {
          const float cA = ex2f(mA - mnA), cB = ex2f(mB - mnB);
          ol[0] *= cA; ol[1] *= cA; ol[2] *= cB; ol[3] *= cB;
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            o[dn][0] *= cA; o[dn][1] *= cA; o[dn][2] *= cB; o[dn][3] *= cB;
          }
          mA = mnA; mB = mnB;
        }
      }
#pragma unroll
      for (int n = 0; n < CHN; ++n) {
        s[n][0] = ex2f(FIXED ? s[n][0] : s[n][0] - mA);
        s[n][1] = ex2f(FIXED ? s[n][1] : s[n][1] - mA);
        s[n][2] = ex2f(FIXED ? s[n][2] : s[n][2] - mB);
        s[n][3] = ex2f(FIXED ? s[n][3] : s[n][3] - mB);
      }
      // O += P V: the accumulator layout of two adjacent column tiles is the A layout of one 16-key step
#pragma unroll
      for (int j = 0; j < CHN / 2; ++j) {
        const uint32_t pa[4] = {pack2(s[2 * j][0], s[2 * j][1]), pack2(s[2 * j][2], s[2 * j][3]),
                                pack2(s[2 * j + 1][0], s[2 * j + 1][1]), pack2(s[2 * j + 1][2], s[2 * j + 1][3])};
        const uint32_t vrow_s = v_s + (uint32_t)(key0 + j * 16) * 64u;
        uint32_t vb[4];
        ldsm4t(vb, vrow_s + lv_off0);
        mma16816(o[0], pa, vb[0], vb[1]);
        mma16816(o[1], pa, vb[2], vb[3]);
        ldsm4t(vb, vrow_s + lv_off2);
        mma16816(o[2], pa, vb[0], vb[1]);
        mma16816(o[3], pa, vb[2], vb[3]);
        mma16816(ol, pa, kOnes, kOnes);           // row sums of the (rounded) P: the weights O is built from
      }
    };
    if (fixed_off) {
      if (need_mask) {
#pragma unroll 1
        for (int c = 0; c < Cf::NCH; ++c) chunk(c, std::true_type{}, std::true_type{});
      } else {
#pragma unroll 1
        for (int c = 0; c < Cf::NCH; ++c) chunk(c, std::false_type{}, std::true_type{});
      }
    } else {
      if (need_mask) {
#pragma unroll 1
        for (int c = 0; c < Cf::NCH; ++c) chunk(c, std::true_type{}, std::false_type{});
      } else {
#pragma unroll 1
        for (int c = 0; c < Cf::NCH; ++c) chunk(c, std::false_type{}, std::false_type{});
      }
    }

    // ---- epilogue: normalise, store O (+ its bf16 residual) and the row's log-sum-exp
    const float lA = ol[0], lB = ol[2];
    mA += boff;
    mB += boff;
    const int tokA = tokS[rA], tokB = tokS[rB];
    float* lse_it = a.lse + (p.win * a.nH + p.h) * N;
    if (tq == 0) {
      if (rA < N) lse_it[rA] = (mA + log2f(lA)) * kLn2;
      if (rB < N) lse_it[rB] = (mB + log2f(lB)) * kLn2;
    }
    const float iA = 1.0f / lA, iB = 1.0f / lB;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int t = half ? tokB : tokA;
      if (t < 0) continue;
      const float inv = half ? iB : iA;
      uint32_t* dst = reinterpret_cast<uint32_t*>(a.out + (int64_t)t * a.C + p.h * HD) + tq;
      uint32_t* dlo = a.out_lo ? reinterpret_cast<uint32_t*>(a.out_lo + (int64_t)t * a.C + p.h * HD) + tq : nullptr;
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        const float v0 = o[dn][2 * half] * inv, v1 = o[dn][2 * half + 1] * inv;
        const uint32_t hi = pack2(v0, v1);
        dst[dn * 4] = hi;
        if (dlo) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi));
          dlo[dn * 4] = pack2(v0 - f.x, v1 - f.y);
        }
      }
    }
  }
  ptx::cp_async_wait<0>();
}

template <int WS>
size_t mma_fwd_smem() {
  using Cf = MCfg<WS>;
  return 1024 + (size_t)kFwdStages * 3 * Cf::TILE + 2 * (size_t)kFwdStages * Cf::NP * 4 + (size_t)Cf::NTAB * 4 + (size_t)Cf::NP * 4 +
         2 * (size_t)Cf::NT * 4 + 16 + 8 * kFwdStages + 16;
}

// CTAs of a kernel that fit one SM: shared memory, threads and the per-sub-partition register file (warps of a CTA are
// dealt round-robin to the four sub-partitions of 16384 registers each).  Computed here because the occupancy query
// answers for the shared-memory carve-out of the moment (1 CTA before the first launch).
template <typename K>
int ctas_per_sm(K kernel, int threads, size_t smem) {
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) return 1;
  const int warps = threads / 32;
  int occ = (int)((227 * 1024) / (smem + 1024));
  if (occ > 2048 / threads) occ = 2048 / threads;
  const int regs_per_warp = ((fa.numRegs + 7) / 8 * 8) * 32;
  while (occ > 1 && ((occ * warps + 3) / 4) * regs_per_warp > 16384) --occ;
  if (occ > 8) occ = 8;
  return occ < 1 ? 1 : occ;
}

template <int WS>
int launch_mma_fwd(const MmaArgs& a, cudaStream_t st) {
  using Cf = MCfg<WS>;
  const size_t smem = mma_fwd_smem<WS>();
  BSW_CUDA(cudaFuncSetAttribute(attn_mma_fwd_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int occ = ctas_per_sm(attn_mma_fwd_kernel<WS>, Cf::THREADS, smem);
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > a.nitems) grid = a.nitems;
  // window boxes of the qkv tensor viewed as [B][H][W][3C]: {32 columns (one head of q, k or v), ws, ws, 1}
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  // (a shape the driver refuses to encode -- none known -- is served by the cp.async gather alone)
  const int use_tma = !Cf::RAGGED && make_tmap_window_bf16(&tmap, a.qkv, a.g.B, a.g.H, a.g.W, 3 * a.C, WS) == B200SWIN_OK;
  attn_mma_fwd_kernel<WS><<<(unsigned)grid, Cf::THREADS, smem, st>>>(a, tmap, use_tma);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

// ------------------------------------------------------------------------------------------------------ backward
// One CTA (NT warps) per (window, head) item, ONE pass, two phases per item:
//   phase 1  warp j owns 16 KEYS: S^T = K_j Q^T and dP^T = V_j dO^T (rows = keys, columns = queries) -> P^T, dS^T in
//            registers, which ARE the A operands of dV_j += P^T dO and dK_j += dS^T Q (no transposes); dS^T also goes to a
//            bf16 panel [key][query] in shared memory; the bias-table gradient is summed in registers across all windows
//            of a head (a thread sees the same (key, query) positions in every item);
//   phase 2  warp i owns 16 QUERIES: dQ_i = dS_i K with dS_i read from the panel (ldmatrix.trans).
// Two CTA barriers per item; the q / k / v / dO tiles, lse and D = <dO, O> of the next item are in flight meanwhile.
constexpr int kBwdStages = 2;

template <int WS>
struct BCfg {
  using Cf = MCfg<WS>;
  static constexpr uint32_t STAGE_TILES = 4 * Cf::TILE;                       // q | k | v | dO
  static constexpr uint32_t PSTRIDE = Cf::NP * 2 + 16;                        // bytes per key row of the dS panel
  static constexpr uint32_t OFF_QMETA = kBwdStages * STAGE_TILES;             // [stages][NP] {lse, D}
  static constexpr uint32_t OFF_INN = OFF_QMETA + kBwdStages * Cf::NP * 8;    // [stages][NP] {1 / ||q||, 1 / ||k||}
  static constexpr uint32_t OFF_TOK = OFF_INN + kBwdStages * Cf::NP * 8;      // [stages][NP]
  static constexpr uint32_t OFF_RID = OFF_TOK + kBwdStages * Cf::NP * 4;      // [stages][NP]
  static constexpr uint32_t OFF_KOF = OFF_RID + kBwdStages * Cf::NP * 4;      // [NP]
  static constexpr uint32_t OFF_TAB = OFF_KOF + Cf::NP * 4;                   // [NTAB]
  static constexpr uint32_t OFF_PANEL = (OFF_TAB + Cf::NTAB * 4 + 127) / 128 * 128;
  static constexpr uint32_t OFF_RED = OFF_PANEL + Cf::NP * PSTRIDE;           // [NT] floats (+ pad to 16 B)
  // per-thread gradient sums that do not fit the register file (168 registers per thread with nine warps: three warps on
  // one SM sub-partition): the bias-gradient sums of the last DBS column tiles and the v_bias gradient, [slot][thread]
  static constexpr int DBS = Cf::NTILES8 >= 18 ? 12 : 0;                       // column tiles (of 8 queries) kept in smem
  static constexpr uint32_t OFF_PADF = OFF_RED + (Cf::NT * 4 + 15) / 16 * 16; // [stages][NT] ints: tile of 16 rows all pad
  static constexpr uint32_t OFF_DBS = OFF_PADF + (kBwdStages * Cf::NT * 4 + 15) / 16 * 16;  // [DBS * 2][THREADS] float2
  static constexpr uint32_t OFF_DVP = OFF_DBS + DBS * 2 * Cf::THREADS * 8;    // [8][THREADS] float
  static constexpr uint32_t SMEM = OFF_DVP + 8 * Cf::THREADS * 4 + 128;
  static_assert((Cf::NP / 2) * Cf::NP * 4 <= Cf::NP * PSTRIDE, "the flush staging lives in the panel");
};

template <int WS>
__global__ void __launch_bounds__(MCfg<WS>::THREADS, MCfg<WS>::BWD_CTAS)
attn_mma_bwd_kernel(const __grid_constant__ MmaArgs a) {
  using Cf = MCfg<WS>;
  using Bc = BCfg<WS>;
  constexpr int N = Cf::N, NP = Cf::NP, TW = Cf::TW, NT8 = Cf::NTILES8;
  constexpr uint32_t TILE = Cf::TILE, PSTRIDE = Bc::PSTRIDE;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 127u) & ~127u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  float2* qmeta = reinterpret_cast<float2*>(sm + Bc::OFF_QMETA);
  float2* innorm = reinterpret_cast<float2*>(sm + Bc::OFF_INN);
  int* tokm = reinterpret_cast<int*>(sm + Bc::OFF_TOK);
  int* ridm = reinterpret_cast<int*>(sm + Bc::OFF_RID);
  int* kofk = reinterpret_cast<int*>(sm + Bc::OFF_KOF);
  float* tab = reinterpret_cast<float*>(sm + Bc::OFF_TAB);
  unsigned char* panel = sm + Bc::OFF_PANEL;
  float* red = reinterpret_cast<float*>(sm + Bc::OFF_RED);
  int* padf = reinterpret_cast<int*>(sm + Bc::OFF_PADF);
  float2* dbs = reinterpret_cast<float2*>(sm + Bc::OFF_DBS) + threadIdx.x;   // + slot * THREADS
  float* dvps = reinterpret_cast<float*>(sm + Bc::OFF_DVP) + threadIdx.x;    // + slot * THREADS
  const uint32_t panel_s = base_u32 + Bc::OFF_PANEL;
  constexpr int DBS = Bc::DBS, NREG = Cf::NTILES8 - DBS;                      // column tiles summed in registers

  const WinGeom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int64_t per = a.nitems / gridDim.x, rem = a.nitems % gridDim.x;
  const int64_t it0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int nit = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));
  const int C3 = 3 * a.C;

  for (int r = tid; r < NP; r += Cf::THREADS) kofk[r] = r < N ? 4 * ((r / WS) * TW + (r % WS)) : 0;

  static_assert(NP * 2 == Cf::THREADS, "two threads per row");
  const int prow = tid >> 1, py = prow / WS, px = prow - py * WS;
  const uint32_t poff = sw64(prow, (tid & 1) * 2);               // second 16 bytes: poff ^ 16
  ItemPos pp = item_pos(a, it0);
  auto prefetch = [&](int i) {
    const int stage = i % kBwdStages;
    unsigned char* q0 = sm + (size_t)stage * Bc::STAGE_TILES;
    const uint32_t q0_s = base_u32 + (uint32_t)stage * Bc::STAGE_TILES;
    int region = 0, t = -2;
    if (!Cf::RAGGED || prow < N) t = row_token<WS>(g, pp, py, px, &region);
    // a warp copies exactly one 16-row tile: is it pad tokens only?  (pad queries have dO = 0: their dS is zero)
    const bool tile_pad = __all_sync(0xffffffffu, t < 0);
    if (lane == 0) padf[stage * Cf::NT + warp] = tile_pad ? 1 : 0;
    if ((tid & 1) == 0) {
      // per-row scalars: {lse, D = <dO, O>} and {1 / ||q||, 1 / ||k||}
      tokm[stage * NP + prow] = t;
      ridm[stage * NP + prow] = region;
      float2* qm = qmeta + stage * NP + prow;
      float2* im = innorm + stage * NP + prow;
      const uint32_t qm_s = base_u32 + Bc::OFF_QMETA + (uint32_t)(stage * NP + prow) * 8u;
      const uint32_t im_s = base_u32 + Bc::OFF_INN + (uint32_t)(stage * NP + prow) * 8u;
      if (t >= 0) {
        ptx::cp_async_4(qm_s, a.lse + (pp.win * a.nH + pp.h) * N + prow);
        ptx::cp_async_4(qm_s + 4, a.dvec + (int64_t)t * a.nH + pp.h);
        ptx::cp_async_4(im_s, a.inv_norm + ((int64_t)t * 2 + 0) * a.nH + pp.h);
        ptx::cp_async_4(im_s + 4, a.inv_norm + ((int64_t)t * 2 + 1) * a.nH + pp.h);
      } else if (t == -1) {
        ptx::cp_async_4(qm_s, a.lse + (pp.win * a.nH + pp.h) * N + prow);
        qm->y = 0.f;
        *im = make_float2(0.f, 0.f);
      } else {
        *qm = make_float2(INFINITY, 0.f);            // beyond the window: P = 0
        *im = make_float2(0.f, 0.f);
      }
    }
    if (t >= 0) {
      const __nv_bfloat16* src = a.qkv + (int64_t)t * C3 + pp.h * HD + (tid & 1) * 16;
      const __nv_bfloat16* gsrc = a.dout + (int64_t)t * a.C + pp.h * HD + (tid & 1) * 16;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        ptx::cp_async_16(q0_s + k * TILE + poff, src + k * a.C);
        ptx::cp_async_16(q0_s + k * TILE + (poff ^ 16u), src + k * a.C + 8);
      }
      ptx::cp_async_16(q0_s + 3 * TILE + poff, gsrc);
      ptx::cp_async_16(q0_s + 3 * TILE + (poff ^ 16u), gsrc + 8);
    } else {
      const float* qp = (t == -1 && a.qpad) ? a.qpad + pp.h * HD + (tid & 1) * 16 : nullptr;
      const float* vp = (t == -1 && a.vpad) ? a.vpad + pp.h * HD + (tid & 1) * 16 : nullptr;
      put16(q0, q0_s, poff, nullptr, qp);
      put16(q0, q0_s, poff ^ 16u, nullptr, qp ? qp + 8 : nullptr);
      put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff, nullptr, vp);
      put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff ^ 16u, nullptr, vp ? vp + 8 : nullptr);
#pragma unroll
      for (int k = 1; k < 4; k += 2) {
        put16(q0 + k * TILE, q0_s + k * TILE, poff, nullptr, nullptr);
        put16(q0 + k * TILE, q0_s + k * TILE, poff ^ 16u, nullptr, nullptr);
      }
    }
    item_next(a, pp);
  };

  if (nit > 0) prefetch(0);
  ptx::cp_async_commit();

  int cur_h = -1;
  float sc = 0.f, scale2 = 0.f, dsc = 0.f;
  const int rA = warp * 16 + gq, rB = rA + 8;      // phase 1: keys; phase 2: queries
  ItemPos p = item_pos(a, it0);
  const uint32_t la_off = sw64(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);      // A fragments of the warp's 16 rows
  const uint32_t lk_off = sw64(lane & 7, lane >> 3);                                            // B fragments, K-major (8 rows x 4 chunks)
  const uint32_t lv_off0 = sw64((lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);                 // B fragments, MN-major (16 rows x chunks 0, 1)
  const uint32_t lv_off2 = sw64((lane & 7) + ((lane >> 3) & 1) * 8, 2 + (lane >> 4));
  // panel: phase 1 writes (key row, query pair); phase 2 reads 8 keys x 8 queries blocks transposed
  const uint32_t pw_offA = (uint32_t)rA * PSTRIDE + (uint32_t)tq * 4u, pw_offB = pw_offA + 8u * PSTRIDE;
  const uint32_t pr_off = (uint32_t)((lane & 7) + (lane >> 4) * 8) * PSTRIDE + (uint32_t)(warp * 16 + ((lane >> 3) & 1) * 8) * 2u;

  float db[NREG > 0 ? NREG : 1][4];                 // bias-table gradient of this thread's (key, query) positions
#pragma unroll
  for (int n = 0; n < NREG; ++n) db[n][0] = db[n][1] = db[n][2] = db[n][3] = 0.f;
#pragma unroll
  for (int n = 0; n < DBS * 2; ++n) dbs[n * Cf::THREADS] = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 8; ++e) dvps[e * Cf::THREADS] = 0.f;   // dV rows of pad keys = gradient of v_bias

  auto flush_head = [&](int h) {
    // every warp is past phase 2 of its last item (the caller sits behind a CTA barrier): the panel is free
    const float s = warp_sum(dsc);
    dsc = 0.f;
    if (lane == 0) red[warp] = s;
    if (a.dvpad) {
#pragma unroll
      for (int dn = 0; dn < 4; ++dn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float v = dvps[(dn * 2 + e) * Cf::THREADS];
          dvps[(dn * 2 + e) * Cf::THREADS] = 0.f;
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (gq == 0 && v != 0.f) atomicAdd(a.dvpad + h * HD + dn * 8 + 2 * tq + e, v);
        }
    }
    float* stg = reinterpret_cast<float*>(panel);   // [NP / 2][NP]
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int lo = half * (NP / 2);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = rr ? rB : rA;
        if (r >= lo && r < lo + NP / 2) {
#pragma unroll
          for (int n = 0; n < NREG; ++n)
            *reinterpret_cast<float2*>(stg + (r - lo) * NP + n * 8 + 2 * tq) = make_float2(db[n][2 * rr], db[n][2 * rr + 1]);
#pragma unroll
          for (int n = 0; n < DBS; ++n)
            *reinterpret_cast<float2*>(stg + (r - lo) * NP + (NREG + n) * 8 + 2 * tq) = dbs[(n * 2 + rr) * Cf::THREADS];
        }
      }
      __syncthreads();
      if (half == 0 && tid == 0) {
        float t = 0.f;
        for (int w = 0; w < Cf::NT; ++w) t += red[w];
        atomicAdd(a.dscale + h, t);
      }
      for (int r = tid; r < Cf::NTAB; r += Cf::THREADS) {
        const int dy = r / TW - (WS - 1), dx = r % TW - (WS - 1);
        float sum = 0.f;
        for (int key = lo; key < lo + NP / 2 && key < N; ++key) {
          const int yk = key / WS, xk = key - yk * WS;
          const int yq = yk + dy, xq = xk + dx;
          if (yq >= 0 && yq < WS && xq >= 0 && xq < WS) sum += stg[(key - lo) * NP + yq * WS + xq];
        }
        if (sum != 0.f) atomicAdd(a.dtable16 + (int64_t)r * a.nH + h, sum);
      }
      __syncthreads();
    }
#pragma unroll
    for (int n = 0; n < NREG; ++n) db[n][0] = db[n][1] = db[n][2] = db[n][3] = 0.f;
#pragma unroll
    for (int n = 0; n < DBS * 2; ++n) dbs[n * Cf::THREADS] = make_float2(0.f, 0.f);
  };

#pragma unroll 1
  for (int i = 0;; ++i, item_next(a, p)) {        // one extra round after the last item: the final flush (single call site)
    const int stage = i % kBwdStages;
    const bool done = i >= nit;
    ptx::cp_async_wait<0>();                      // this thread's copies of item i have landed
    __syncthreads();                              // (B1) everybody's; every warp is done with item i - 1 (panel, other stage)
    if (done || p.h != cur_h) {                   // CTA-uniform
      if (cur_h >= 0) flush_head(cur_h);
      if (done) break;
      for (int t = tid; t < Cf::NTAB; t += Cf::THREADS) tab[t] = a.table16[(int64_t)t * a.nH + p.h] * kLog2e;
      sc = a.scale[p.h];
      scale2 = sc * kLog2e;
      cur_h = p.h;
      __syncthreads();
    }
    if (i + 1 < nit) prefetch(i + 1);
    ptx::cp_async_commit();

    const bool need_mask = g.shift > 0 && (p.wh == g.nWh - 1 || p.ww == g.nWw - 1);
    uint32_t padmask = 0;                         // bit c: the 16 queries of step / tile c are all pad tokens (CTA-uniform)
#pragma unroll
    for (int w = 0; w < Cf::NT; ++w) padmask |= (uint32_t)padf[stage * Cf::NT + w] << w;
    const uint32_t q_s = base_u32 + (uint32_t)stage * Bc::STAGE_TILES, k_s = q_s + TILE, v_s = k_s + TILE, g_s = v_s + TILE;
    const int* tokS = tokm + stage * NP;
    const int* ridS = ridm + stage * NP;
    const float4* qm4 = reinterpret_cast<const float4*>(qmeta + stage * NP);
    const int tokA = tokS[rA], tokB = tokS[rB];

    // ------------------------------------------------------------------ phase 1: this warp's 16 keys x all queries
    {
      // bias entry of (query, key) = tab[kof(query) - kof(key) + (WS-1)(TW+1)]
      const uint32_t tab_s = ptx::smem_u32(tab) + 4u * (uint32_t)((WS - 1) * (TW + 1));
      const uint32_t tabA = tab_s - (uint32_t)kofk[rA], tabB = tab_s - (uint32_t)kofk[rB];
      const int ridA = ridS[rA], ridB = ridS[rB];
      uint32_t ka[2][4], va[2][4];
      ldsm4(ka[0], k_s + la_off);
      ldsm4(ka[1], k_s + (la_off ^ 32u));
      ldsm4(va[0], v_s + la_off);
      ldsm4(va[1], v_s + (la_off ^ 32u));
      float dk[4][4], dv[4][4];
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
        dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
      }
      // S^T and dP^T of 16 queries: both k-steps of every accumulator apart from each other (a dependent HMMA pair stalls)
      auto sdp = [&](int c, float (&st)[2][4], float (&dp)[2][4]) {
        uint32_t bq[2][4], bg[2][4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          ldsm4(bq[n], q_s + (uint32_t)(c * 16 + n * 8) * 64u + lk_off);
          ldsm4(bg[n], g_s + (uint32_t)(c * 16 + n * 8) * 64u + lk_off);
          st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
          dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            mma16816(st[n], ka[ks], bq[n][2 * ks], bq[n][2 * ks + 1]);
            mma16816(dp[n], va[ks], bg[n][2 * ks], bg[n][2 * ks + 1]);
          }
      };
      auto sweep = [&](auto mask_c) {
        constexpr bool MASK = decltype(mask_c)::value;
#pragma unroll
        for (int c = 0; c < NT8 / 2; ++c) {         // 16 queries per step
          if ((padmask >> c) & 1u) continue;        // pad queries only (CTA-uniform): dS = 0, nothing to add anywhere
          float st[2][4], dp[2][4];
          sdp(c, st, dp);
          // B operands of this step's dV / dK contractions: in flight during the element work
          const uint32_t row16 = (uint32_t)(c * 16) * 64u;
          uint32_t bg0[4], bg2[4], bq0[4], bq2[4];
          ldsm4t(bg0, g_s + row16 + lv_off0);
          ldsm4t(bg2, g_s + row16 + lv_off2);
          ldsm4t(bq0, q_s + row16 + lv_off0);
          ldsm4t(bq2, q_s + row16 + lv_off2);
          uint32_t aP[4], aD[4];
#pragma unroll
          for (int n = 0; n < 2; ++n) {
            const int qcol = c * 16 + n * 8 + 2 * tq;
            const float4 m = qm4[(c * 16 + n * 8) / 2 + tq];             // {lse, D} of queries qcol, qcol + 1
            int2 kq;
            if (WS % 2 == 0) kq.x = kof_pair<WS>(c * 16 + n * 8, tq), kq.y = kq.x + 4;
            else kq = *reinterpret_cast<const int2*>(kofk + qcol);
            const uint32_t aA0 = tabA + (uint32_t)kq.x, aB0 = tabB + (uint32_t)kq.x;
            const uint32_t aA1 = (WS % 2 == 0) ? aA0 + 4u : tabA + (uint32_t)kq.y;
            const uint32_t aB1 = (WS % 2 == 0) ? aB0 + 4u : tabB + (uint32_t)kq.y;
            float s2[4] = {fmaf(st[n][0], scale2, lds32(aA0)), fmaf(st[n][1], scale2, lds32(aA1)),
                           fmaf(st[n][2], scale2, lds32(aB0)), fmaf(st[n][3], scale2, lds32(aB1))};
            if (MASK) {
              const int2 rr = *reinterpret_cast<const int2*>(ridS + qcol);
              if (rr.x != ridA) s2[0] += kMaskLog2;
              if (rr.y != ridA) s2[1] += kMaskLog2;
              if (rr.x != ridB) s2[2] += kMaskLog2;
              if (rr.y != ridB) s2[3] += kMaskLog2;
            }
            float pv[4], ds[4];
            pv[0] = ex2f(fmaf(m.x, -kLog2e, s2[0]));
            pv[1] = ex2f(fmaf(m.z, -kLog2e, s2[1]));
            pv[2] = ex2f(fmaf(m.x, -kLog2e, s2[2]));
            pv[3] = ex2f(fmaf(m.z, -kLog2e, s2[3]));
            if (Cf::RAGGED) {                       // key rows beyond the window
              if (rA >= N) pv[0] = pv[1] = 0.f;
              if (rB >= N) pv[2] = pv[3] = 0.f;
            }
            ds[0] = pv[0] * (dp[n][0] - m.y);
            ds[1] = pv[1] * (dp[n][1] - m.w);
            ds[2] = pv[2] * (dp[n][2] - m.y);
            ds[3] = pv[3] * (dp[n][3] - m.w);
            if (c * 2 + n < NREG) {
#pragma unroll
              for (int e = 0; e < 4; ++e) db[c * 2 + n < NREG ? c * 2 + n : 0][e] += ds[e];
            } else {
              float2* slot = dbs + ((c * 2 + n - NREG) * 2) * Cf::THREADS;
              float2 u = slot[0], w = slot[Cf::THREADS];
              u.x += ds[0]; u.y += ds[1]; w.x += ds[2]; w.y += ds[3];
              slot[0] = u; slot[Cf::THREADS] = w;
            }
            aP[2 * n] = pack2(pv[0], pv[1]);
            aP[2 * n + 1] = pack2(pv[2], pv[3]);
            aD[2 * n] = pack2(ds[0], ds[1]);
            aD[2 * n + 1] = pack2(ds[2], ds[3]);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(panel_s + pw_offA + (uint32_t)(c * 16 + n * 8) * 2u), "r"(aD[2 * n]) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(panel_s + pw_offB + (uint32_t)(c * 16 + n * 8) * 2u), "r"(aD[2 * n + 1]) : "memory");
          }
          // dV += P^T dO, dK += dS^T Q over these 16 queries (eight independent accumulators)
          mma16816(dv[0], aP, bg0[0], bg0[1]);
          mma16816(dv[1], aP, bg0[2], bg0[3]);
          mma16816(dv[2], aP, bg2[0], bg2[1]);
          mma16816(dv[3], aP, bg2[2], bg2[3]);
          mma16816(dk[0], aD, bq0[0], bq0[1]);
          mma16816(dk[1], aD, bq0[2], bq0[3]);
          mma16816(dk[2], aD, bq2[0], bq2[1]);
          mma16816(dk[3], aD, bq2[2], bq2[3]);
        }
      };
      if (need_mask) sweep(std::true_type{}); else sweep(std::false_type{});

      // ---- dV, dK of the warp's 16 keys.  k_hat of row g / g + 8 sits in the A fragments at the accumulator's columns.
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int t = half ? tokB : tokA;
        // (the quad sum runs on every lane: a pad row's lanes must not skip a full-mask shuffle)
        float kh[4][2], dot = 0.f;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          const uint32_t w = ka[dn >> 1][(dn & 1) * 2 + half];
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
          kh[dn][0] = f.x; kh[dn][1] = f.y;
          dot = fmaf(dk[dn][2 * half], f.x, dot);
          dot = fmaf(dk[dn][2 * half + 1], f.y, dot);
        }
        dot = quad_sum(dot) * sc;
        if (t >= 0) {
          const float invn = innorm[stage * NP + (half ? rB : rA)].y;          // 1 / ||k||
          uint32_t* dkd = reinterpret_cast<uint32_t*>(a.dqkv + (int64_t)t * C3 + a.C + p.h * HD) + tq;
          uint32_t* dvd = reinterpret_cast<uint32_t*>(a.dqkv + (int64_t)t * C3 + 2 * a.C + p.h * HD) + tq;
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            dkd[dn * 4] = pack2((dk[dn][2 * half] * sc - kh[dn][0] * dot) * invn, (dk[dn][2 * half + 1] * sc - kh[dn][1] * dot) * invn);
            dvd[dn * 4] = pack2(dv[dn][2 * half], dv[dn][2 * half + 1]);
          }
        } else if (t == -1) {
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            dvps[(dn * 2) * Cf::THREADS] += dv[dn][2 * half];
            dvps[(dn * 2 + 1) * Cf::THREADS] += dv[dn][2 * half + 1];
          }
        }
      }
    }
    __syncthreads();                              // (B2) the dS panel is complete

    // ------------------------------------------------------------------ phase 2: this warp's 16 queries, dQ = dS K
    if (!((padmask >> warp) & 1u)) {
      float dq[4][4];
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) dq[dn][0] = dq[dn][1] = dq[dn][2] = dq[dn][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < NP / 16; ++ks) {
        uint32_t ad[4], b[4];
        ldsm4t(ad, panel_s + (uint32_t)(ks * 16) * PSTRIDE + pr_off);
        ldsm4t(b, k_s + (uint32_t)(ks * 16) * 64u + lv_off0);
        mma16816(dq[0], ad, b[0], b[1]);
        mma16816(dq[1], ad, b[2], b[3]);
        ldsm4t(b, k_s + (uint32_t)(ks * 16) * 64u + lv_off2);
        mma16816(dq[2], ad, b[0], b[1]);
        mma16816(dq[3], ad, b[2], b[3]);
      }
      uint32_t qa[2][4];
      ldsm4(qa[0], q_s + la_off);
      ldsm4(qa[1], q_s + (la_off ^ 32u));
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int t = half ? tokB : tokA;
        float qh[4][2], dot = 0.f;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          const uint32_t w = qa[dn >> 1][(dn & 1) * 2 + half];
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
          qh[dn][0] = f.x; qh[dn][1] = f.y;
          dot = fmaf(dq[dn][2 * half], f.x, dot);
          dot = fmaf(dq[dn][2 * half + 1], f.y, dot);
        }
        dot = quad_sum(dot);
        // gradient of the temperature: sum_k dS[q, k] cos[q, k] = <q_hat, sum_k dS[q, k] k_hat> = this very dot product
        if (tq == 0) dsc += dot;
        dot *= sc;
        if (t < 0) continue;
        const float invn = innorm[stage * NP + (half ? rB : rA)].x;            // 1 / ||q||
        uint32_t* dqd = reinterpret_cast<uint32_t*>(a.dqkv + (int64_t)t * C3 + p.h * HD) + tq;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
          dqd[dn * 4] = pack2((dq[dn][2 * half] * sc - qh[dn][0] * dot) * invn, (dq[dn][2 * half + 1] * sc - qh[dn][1] * dot) * invn);
      }
    }
  }
}

template <int WS>
int launch_mma_bwd(const MmaArgs& a, cudaStream_t st) {
  using Cf = MCfg<WS>;
  const size_t smem = BCfg<WS>::SMEM;
  BSW_REQUIRE(smem <= 227 * 1024, "attn_bwd(mma): window %dx%d needs %zu bytes of shared memory", WS, WS, smem);
  BSW_CUDA(cudaFuncSetAttribute(attn_mma_bwd_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // (no carve-out preference: the driver sizes shared memory to the request and the rest stays L1 for the few spills)
  int64_t grid = (int64_t)sm_count() * ctas_per_sm(attn_mma_bwd_kernel<WS>, Cf::THREADS, smem);
  if (grid > a.nitems) grid = a.nitems;
  attn_mma_bwd_kernel<WS><<<(unsigned)grid, Cf::THREADS, smem, st>>>(a);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

// ------------------------------------------------------------------------------ backward, warp-specialised (12x12)
// Same two phases, but with nine warps per SM the CTA barriers between the phases and the item prologue cost a third
// of the time (nothing else is resident to fill them).  Here the nine KEY warps run phase 1 of item after item without
// ever meeting a CTA barrier; three HELPER warps (one on each of the SM sub-partitions that hold only two key warps)
// gather the operands (cp.async, completion on an mbarrier per stage) and run phase 2 -- dQ = dS K from the dS panel
// of the item, its normalisation and the temperature gradient -- one item behind.  The panel is double-buffered.
//   full[s]   helpers -> key warps   operands of the item in stage s have landed
//   pfull[s]  key warps -> helpers   panel s complete, stage s no longer read by phase 1
//   pfree[s]  helpers -> key warps   panel s consumed
template <int WS>
struct SCfg {
  using Cf = MCfg<WS>;
  static constexpr int NH = 3;                                                // helper warps
  static constexpr int KTHREADS = Cf::THREADS, HTHREADS = NH * 32, THREADS = KTHREADS + HTHREADS;
  static constexpr uint32_t STAGE_TILES = 4 * Cf::TILE;
  static constexpr uint32_t PSTRIDE = Cf::NP * 2 + 16;
  static constexpr uint32_t PANEL = Cf::NP * PSTRIDE;
  static constexpr uint32_t OFF_QMETA = 2 * STAGE_TILES;
  static constexpr uint32_t OFF_INN = OFF_QMETA + 2 * Cf::NP * 8;
  static constexpr uint32_t OFF_TOK = OFF_INN + 2 * Cf::NP * 8;
  static constexpr uint32_t OFF_RID = OFF_TOK + 2 * Cf::NP * 4;
  static constexpr uint32_t OFF_KOF = OFF_RID + 2 * Cf::NP * 4;
  static constexpr uint32_t OFF_TAB = OFF_KOF + Cf::NP * 4;
  static constexpr uint32_t OFF_PANEL = (OFF_TAB + Cf::NTAB * 4 + 127) / 128 * 128;   // two panels
  static constexpr uint32_t OFF_RED = OFF_PANEL + 2 * PANEL;                  // [NT] floats
  static constexpr uint32_t OFF_PADF = OFF_RED + (Cf::NT * 4 + 15) / 16 * 16; // [2][NT] ints
  static constexpr uint32_t OFF_BARS = OFF_PADF + (2 * Cf::NT * 4 + 15) / 16 * 16;    // 6 mbarriers
  static constexpr int DBS = 8;                                               // column tiles whose sums live in smem
  static constexpr uint32_t OFF_DBS = OFF_BARS + 64;
  static constexpr uint32_t OFF_DVP = OFF_DBS + DBS * 2 * KTHREADS * 8;
  static constexpr uint32_t OFF_DVS = OFF_DVP + 8 * KTHREADS * 4;              // [8][KTHREADS] float: column sums of dV
  static constexpr uint32_t SMEM = OFF_DVS + 8 * KTHREADS * 4 + 1024;        // + alignment of the base to 1 KB (TMA swizzle period)
  static_assert((Cf::NP / 2) * Cf::NP * 4 <= PANEL, "the flush staging lives in a panel");
  static_assert(Cf::NT % NH == 0, "tiles per helper warp");
};

// Operand staging as in the forward: a window in one piece on the map = four 4-D TMA boxes (q, k, v of the qkv tensor, dO
// of the dout tensor) issued by one helper thread and completing on full[s] together with the helpers' arrivals; the zero
// fill of rows beyond H / W is the k and the dO of a pad token, a pad QUERY contributes nothing (lse = +inf), and the v =
// v_bias of pad KEYS is written by the key warp that owns those 16 rows (nobody else reads them) after the wait.
template <int WS>
__global__ void __launch_bounds__(SCfg<WS>::THREADS, 1)
attn_mma_bwd_spec_kernel(const __grid_constant__ MmaArgs a, const __grid_constant__ CUtensorMap tm_qkv,
                         const __grid_constant__ CUtensorMap tm_dout, const int use_tma) {
  using Cf = MCfg<WS>;
  using Sc = SCfg<WS>;
  constexpr int N = Cf::N, NP = Cf::NP, TW = Cf::TW, NT8 = Cf::NTILES8, NT = Cf::NT;
  constexpr int KTH = Sc::KTHREADS, HTH = Sc::HTHREADS;
  constexpr uint32_t TILE = Cf::TILE, PSTRIDE = Sc::PSTRIDE;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  float2* qmeta = reinterpret_cast<float2*>(sm + Sc::OFF_QMETA);
  float2* innorm = reinterpret_cast<float2*>(sm + Sc::OFF_INN);
  int* tokm = reinterpret_cast<int*>(sm + Sc::OFF_TOK);
  int* ridm = reinterpret_cast<int*>(sm + Sc::OFF_RID);
  int* kofk = reinterpret_cast<int*>(sm + Sc::OFF_KOF);
  float* tab = reinterpret_cast<float*>(sm + Sc::OFF_TAB);
  float* red = reinterpret_cast<float*>(sm + Sc::OFF_RED);
  int* padf = reinterpret_cast<int*>(sm + Sc::OFF_PADF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + Sc::OFF_BARS);
  uint64_t* full = bars;          // [2]
  uint64_t* pfull = bars + 2;     // [2]
  uint64_t* pfree = bars + 4;     // [2]

  const WinGeom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int64_t per = a.nitems / gridDim.x, rem = a.nitems % gridDim.x;
  const int64_t it0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int nit = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));
  const int C3 = 3 * a.C;
  const bool tma_on = !Cf::RAGGED && use_tma != 0;
  auto in_one_piece = [&](const ItemPos& q) { return tma_on && (g.shift == 0 || (q.wh < g.nWh - 1 && q.ww < g.nWw - 1)); };

  for (int r = tid; r < NP; r += Sc::THREADS) kofk[r] = r < N ? 4 * ((r / WS) * TW + (r % WS)) : 0;
  if (tid == 0) {
    if (tma_on) { ptx::prefetch_tmap(&tm_qkv); ptx::prefetch_tmap(&tm_dout); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&full[s], 2 * HTH);          // per helper thread: its cp.async have landed + its plain stores are released
      ptx::mbar_init(&pfull[s], NT);              // lane 0 of every key warp
      ptx::mbar_init(&pfree[s], Sc::NH);          // lane 0 of every helper warp
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();

  const uint32_t lk_off = sw64(lane & 7, lane >> 3);                                            // B fragments, K-major (8 rows x 4 chunks)
  const uint32_t lv_off0 = sw64((lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);                 // B fragments, MN-major (16 rows x chunks 0, 1)
  const uint32_t lv_off2 = sw64((lane & 7) + ((lane >> 3) & 1) * 8, 2 + (lane >> 4));

  if (warp >= NT) {
    // ============================================================================================== helper warps
    const int hw = warp - NT, hid = tid - KTH;
    ItemPos pp = item_pos(a, it0);               // cursor of the gather stream
    auto gather = [&](int i) {
      const int stage = i & 1;
      unsigned char* q0 = sm + (size_t)stage * Sc::STAGE_TILES;
      const uint32_t q0_s = base_u32 + (uint32_t)stage * Sc::STAGE_TILES;
      const bool one = in_one_piece(pp);
#pragma unroll 1
      for (int k = 0; k < NT / Sc::NH; ++k) {     // a helper warp copies the tile 3 k + hw in round k
        const int task = k * HTH + hid, prow = task >> 1, half = task & 1;
        const int py = prow / WS, px = prow - py * WS;
        const uint32_t poff = sw64(prow, half * 2);
        int region = 0, t = -2;
        if (!Cf::RAGGED || prow < N) t = row_token<WS>(g, pp, py, px, &region);
        const bool tile_pad = __all_sync(0xffffffffu, t < 0);
        if (lane == 0) padf[stage * NT + k * Sc::NH + hw] = tile_pad ? 1 : 0;
        if (half == 0) {
          tokm[stage * NP + prow] = t;
          ridm[stage * NP + prow] = region;
          float2* qm = qmeta + stage * NP + prow;
          float2* im = innorm + stage * NP + prow;
          const uint32_t qm_s = base_u32 + Sc::OFF_QMETA + (uint32_t)(stage * NP + prow) * 8u;
          const uint32_t im_s = base_u32 + Sc::OFF_INN + (uint32_t)(stage * NP + prow) * 8u;
          if (t >= 0) {
            ptx::cp_async_4(qm_s, a.lse + (pp.win * a.nH + pp.h) * N + prow);
            ptx::cp_async_4(qm_s + 4, a.dvec + (int64_t)t * a.nH + pp.h);
            ptx::cp_async_4(im_s, a.inv_norm + ((int64_t)t * 2 + 0) * a.nH + pp.h);
            ptx::cp_async_4(im_s + 4, a.inv_norm + ((int64_t)t * 2 + 1) * a.nH + pp.h);
          } else {
            *qm = make_float2(INFINITY, 0.f);        // pad query (dO = 0) or beyond the window: P = 0, dS = 0
            *im = make_float2(0.f, 0.f);
          }
        }
        if (one) continue;                        // the tiles come as TMA boxes (below)
        if (t >= 0) {
          const __nv_bfloat16* src = a.qkv + (int64_t)t * C3 + pp.h * HD + half * 16;
          const __nv_bfloat16* gsrc = a.dout + (int64_t)t * a.C + pp.h * HD + half * 16;
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            ptx::cp_async_16(q0_s + kk * TILE + poff, src + kk * a.C);
            ptx::cp_async_16(q0_s + kk * TILE + (poff ^ 16u), src + kk * a.C + 8);
          }
          ptx::cp_async_16(q0_s + 3 * TILE + poff, gsrc);
          ptx::cp_async_16(q0_s + 3 * TILE + (poff ^ 16u), gsrc + 8);
        } else {
          const float* qp = (t == -1 && a.qpad) ? a.qpad + pp.h * HD + half * 16 : nullptr;
          const float* vp = (t == -1 && a.vpad) ? a.vpad + pp.h * HD + half * 16 : nullptr;
          put16(q0, q0_s, poff, nullptr, qp);
          put16(q0, q0_s, poff ^ 16u, nullptr, qp ? qp + 8 : nullptr);
          put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff, nullptr, vp);
          put16(q0 + 2 * TILE, q0_s + 2 * TILE, poff ^ 16u, nullptr, vp ? vp + 8 : nullptr);
#pragma unroll
          for (int kk = 1; kk < 4; kk += 2) {
            put16(q0 + kk * TILE, q0_s + kk * TILE, poff, nullptr, nullptr);
            put16(q0 + kk * TILE, q0_s + kk * TILE, poff ^ 16u, nullptr, nullptr);
          }
        }
      }
      ptx::cp_async_mbar_arrive_noinc(&full[stage]);
      if (one && hid == 0) {
        ptx::fence_proxy_async_smem();            // the stage was last read / written through the generic proxy
        ptx::mbar_arrive_expect_tx(&full[stage], 4u * TILE);
        const int x0 = pp.ww * WS + g.shift, y0 = pp.wh * WS + g.shift;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) ptx::tma_load_4d(q0 + kk * TILE, &tm_qkv, &full[stage], kk * a.C + pp.h * HD, x0, y0, pp.b);
        ptx::tma_load_4d(q0 + 3 * TILE, &tm_dout, &full[stage], pp.h * HD, x0, y0, pp.b);
      } else {
        ptx::mbar_arrive(&full[stage]);
      }
      item_next(a, pp);
    };
    if (nit > 0) gather(0);
    if (nit > 1) gather(1);

    ItemPos p = item_pos(a, it0);
    int cur_h = -1;
    float dsc = 0.f;
    float dqs[4][2];                              // column sums of dq over this warp's rows (gradient of q_bias)
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) dqs[dn][0] = dqs[dn][1] = 0.f;
    auto flush_dsc = [&](int h) {
      const float s = warp_sum(dsc);
      dsc = 0.f;
      if (lane == 0 && s != 0.f) atomicAdd(a.dscale + h, s);
      if (a.dcol) {
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float v = dqs[dn][e];
            dqs[dn][e] = 0.f;
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (gq == 0 && v != 0.f) atomicAdd(a.dcol + h * HD + dn * 8 + 2 * tq + e, v);
          }
      }
    };
#pragma unroll 1
    for (int i = 0; i < nit; ++i, item_next(a, p)) {
      const int stage = i & 1;
      const uint32_t par = (uint32_t)(i >> 1) & 1u;
      if (p.h != cur_h) {
        if (cur_h >= 0) flush_dsc(cur_h);
        cur_h = p.h;
      }
      const float sc = a.scale[p.h];
      ptx::mbar_wait(&pfull[stage], par);          // panel complete (implies: this stage's operands landed long ago)
      const uint32_t q_s = base_u32 + (uint32_t)stage * Sc::STAGE_TILES, k_s = q_s + TILE;
      const uint32_t panel_s = base_u32 + Sc::OFF_PANEL + (uint32_t)stage * Sc::PANEL;
      const int* tokS = tokm + stage * NP;
      {
        constexpr int TPH = NT / Sc::NH;           // query tiles of this helper warp: hw, hw + NH, ...
        bool live[TPH];
        uint32_t pr_off[TPH];
        float dq[TPH][4][4];
#pragma unroll
        for (int k = 0; k < TPH; ++k) {
          const int tile = k * Sc::NH + hw;
          live[k] = !padf[stage * NT + tile];      // pad queries only: dQ rows of tokens that do not exist
          pr_off[k] = (uint32_t)((lane & 7) + (lane >> 4) * 8) * PSTRIDE + (uint32_t)(tile * 16 + ((lane >> 3) & 1) * 8) * 2u;
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) dq[k][dn][0] = dq[k][dn][1] = dq[k][dn][2] = dq[k][dn][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < NP / 16; ++ks) {     // one load of the K fragments of 16 keys serves all tiles
          uint32_t b0[4], b2[4];
          ldsm4t(b0, k_s + (uint32_t)(ks * 16) * 64u + lv_off0);
          ldsm4t(b2, k_s + (uint32_t)(ks * 16) * 64u + lv_off2);
#pragma unroll
          for (int k = 0; k < TPH; ++k) {
            if (!live[k]) continue;
            uint32_t ad[4];
            ldsm4t(ad, panel_s + (uint32_t)(ks * 16) * PSTRIDE + pr_off[k]);
            mma16816(dq[k][0], ad, b0[0], b0[1]);
            mma16816(dq[k][1], ad, b0[2], b0[3]);
            mma16816(dq[k][2], ad, b2[0], b2[1]);
            mma16816(dq[k][3], ad, b2[2], b2[3]);
          }
        }
#pragma unroll
        for (int k = 0; k < TPH; ++k) {
          if (!live[k]) continue;
          const int tile = k * Sc::NH + hw;
          const int rA = tile * 16 + gq, rB = rA + 8;
          const uint32_t la_off = sw64(tile * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);
          uint32_t qa[2][4];
          ldsm4(qa[0], q_s + la_off);
          ldsm4(qa[1], q_s + (la_off ^ 32u));
          const int tokA = tokS[rA], tokB = tokS[rB];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int t = half ? tokB : tokA;
            float qh[4][2], dot = 0.f;
#pragma unroll
            for (int dn = 0; dn < 4; ++dn) {
              const uint32_t w = qa[dn >> 1][(dn & 1) * 2 + half];
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
              qh[dn][0] = f.x; qh[dn][1] = f.y;
              dot = fmaf(dq[k][dn][2 * half], f.x, dot);
              dot = fmaf(dq[k][dn][2 * half + 1], f.y, dot);
            }
            dot = quad_sum(dot);
            // gradient of the temperature: sum_k dS[q, k] cos[q, k] = <q_hat, sum_k dS[q, k] k_hat> = this very dot product
            if (tq == 0) dsc += dot;
            dot *= sc;
            if (t < 0) continue;
            const float invn = innorm[stage * NP + (half ? rB : rA)].x;            // 1 / ||q||
            uint32_t* dqd = reinterpret_cast<uint32_t*>(a.dqkv + (int64_t)t * C3 + p.h * HD) + tq;
#pragma unroll
            for (int dn = 0; dn < 4; ++dn) {
              const float g0 = (dq[k][dn][2 * half] * sc - qh[dn][0] * dot) * invn;
              const float g1 = (dq[k][dn][2 * half + 1] * sc - qh[dn][1] * dot) * invn;
              dqd[dn * 4] = pack2(g0, g1);
              dqs[dn][0] += g0;
              dqs[dn][1] += g1;
            }
          }
        }
      }
      // every helper is done with panel / stage `stage`: hand the panel back and refill the stage
      ptx::named_bar_sync(2, HTH);
      if (lane == 0) ptx::mbar_arrive(&pfree[stage]);
      if (i + 2 < nit) gather(i + 2);
    }
    if (cur_h >= 0) flush_dsc(cur_h);
    ptx::cp_async_wait<0>();
    return;
  }

  // ================================================================================================== key warps
  float2* dbs = reinterpret_cast<float2*>(sm + Sc::OFF_DBS) + tid;           // + slot * KTH
  float* dvps = reinterpret_cast<float*>(sm + Sc::OFF_DVP) + tid;            // + slot * KTH
  float* dvs = reinterpret_cast<float*>(sm + Sc::OFF_DVS) + tid;             // + slot * KTH
  constexpr int DBS = Sc::DBS, NREG = NT8 - DBS;
  int cur_h = -1;
  float sc = 0.f, scale2 = 0.f;
  const int rA = warp * 16 + gq, rB = rA + 8;
  ItemPos p = item_pos(a, it0);
  const uint32_t la_off = sw64(warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, lane >> 4);
  const uint32_t pw_offA = (uint32_t)rA * PSTRIDE + (uint32_t)tq * 4u, pw_offB = pw_offA + 8u * PSTRIDE;

  float db[NREG][4];
#pragma unroll
  for (int n = 0; n < NREG; ++n) db[n][0] = db[n][1] = db[n][2] = db[n][3] = 0.f;
#pragma unroll
  for (int n = 0; n < DBS * 2; ++n) dbs[n * KTH] = make_float2(0.f, 0.f);
#pragma unroll
  for (int e = 0; e < 8; ++e) dvps[e * KTH] = dvs[e * KTH] = 0.f;

  auto flush_head = [&](int h, unsigned char* panel) {
    // key warps only (named barrier 1); `panel` is a panel the helpers have handed back
    if (a.dcol) {
#pragma unroll
      for (int dn = 0; dn < 4; ++dn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float v = dvs[(dn * 2 + e) * KTH];
          dvs[(dn * 2 + e) * KTH] = 0.f;
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (gq == 0 && v != 0.f) atomicAdd(a.dcol + 2 * a.C + h * HD + dn * 8 + 2 * tq + e, v);
        }
    }
    if (a.dvpad) {
#pragma unroll
      for (int dn = 0; dn < 4; ++dn)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float v = dvps[(dn * 2 + e) * KTH];
          dvps[(dn * 2 + e) * KTH] = 0.f;
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (gq == 0 && v != 0.f) atomicAdd(a.dvpad + h * HD + dn * 8 + 2 * tq + e, v);
        }
    }
    float* stg = reinterpret_cast<float*>(panel);   // [NP / 2][NP]
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int lo = half * (NP / 2);
      ptx::named_bar_sync(1, KTH);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = rr ? rB : rA;
        if (r >= lo && r < lo + NP / 2) {
#pragma unroll
          for (int n = 0; n < NREG; ++n)
            *reinterpret_cast<float2*>(stg + (r - lo) * NP + n * 8 + 2 * tq) = make_float2(db[n][2 * rr], db[n][2 * rr + 1]);
#pragma unroll
          for (int n = 0; n < DBS; ++n)
            *reinterpret_cast<float2*>(stg + (r - lo) * NP + (NREG + n) * 8 + 2 * tq) = dbs[(n * 2 + rr) * KTH];
        }
      }
      ptx::named_bar_sync(1, KTH);
      for (int r = tid; r < Cf::NTAB; r += KTH) {
        const int dy = r / TW - (WS - 1), dx = r % TW - (WS - 1);
        float sum = 0.f;
        for (int key = lo; key < lo + NP / 2 && key < N; ++key) {
          const int yk = key / WS, xk = key - yk * WS;
          const int yq = yk + dy, xq = xk + dx;
          if (yq >= 0 && yq < WS && xq >= 0 && xq < WS) sum += stg[(key - lo) * NP + yq * WS + xq];
        }
        if (sum != 0.f) atomicAdd(a.dtable16 + (int64_t)r * a.nH + h, sum);
      }
    }
    ptx::named_bar_sync(1, KTH);
#pragma unroll
    for (int n = 0; n < NREG; ++n) db[n][0] = db[n][1] = db[n][2] = db[n][3] = 0.f;
#pragma unroll
    for (int n = 0; n < DBS * 2; ++n) dbs[n * KTH] = make_float2(0.f, 0.f);
  };

#pragma unroll 1
  for (int i = 0;; ++i, item_next(a, p)) {        // one extra round after the last item: the final flush (single call site)
    const int stage = i & 1;
    const uint32_t par = (uint32_t)(i >> 1) & 1u;
    const bool done = i >= nit;
    // panel `stage` is free once the helpers have consumed item i - 2 (a fresh barrier passes the wait on parity 1)
    ptx::mbar_wait(&pfree[stage], par ^ 1u);
    unsigned char* panel = sm + Sc::OFF_PANEL + (size_t)stage * Sc::PANEL;
    const uint32_t panel_s = base_u32 + Sc::OFF_PANEL + (uint32_t)stage * Sc::PANEL;
    if (done || p.h != cur_h) {                   // uniform over the key warps
      if (cur_h >= 0) flush_head(cur_h, panel);
      if (done) break;
      ptx::named_bar_sync(1, KTH);                // nobody still reads the old table
      for (int t = tid; t < Cf::NTAB; t += KTH) tab[t] = a.table16[(int64_t)t * a.nH + p.h] * kLog2e;
      sc = a.scale[p.h];
      scale2 = sc * kLog2e;
      cur_h = p.h;
      ptx::named_bar_sync(1, KTH);
    }
    ptx::mbar_wait(&full[stage], par);            // the operands of item i have landed

    const bool need_mask = g.shift > 0 && (p.wh == g.nWh - 1 || p.ww == g.nWw - 1);
    uint32_t padmask = 0;                         // bit c: the 16 queries of step / tile c are all pad tokens
#pragma unroll
    for (int w = 0; w < NT; ++w) padmask |= (uint32_t)padf[stage * NT + w] << w;
    const uint32_t q_s = base_u32 + (uint32_t)stage * Sc::STAGE_TILES, k_s = q_s + TILE, v_s = k_s + TILE, g_s = v_s + TILE;
    const int* tokS = tokm + stage * NP;
    const int* ridS = ridm + stage * NP;
    const float4* qm4 = reinterpret_cast<const float4*>(qmeta + stage * NP);
    const int tokA = tokS[rA], tokB = tokS[rB];
    // bias entry of (query, key) = tab[kof(query) - kof(key) + (WS-1)(TW+1)]
    const uint32_t tab_s = ptx::smem_u32(tab) + 4u * (uint32_t)((WS - 1) * (TW + 1));
    const uint32_t tabA = tab_s - (uint32_t)kofk[rA], tabB = tab_s - (uint32_t)kofk[rB];
    const int ridA = ridS[rA], ridB = ridS[rB];
    if (in_one_piece(p) && (p.wh * WS + g.shift + WS > g.H || p.ww * WS + g.shift + WS > g.W)) {
      // TMA-staged window that overhangs the map: this warp's pad keys get v = v_bias (their k is the zero fill)
      const int row = warp * 16 + (lane >> 1);
      if (a.vpad && tokS[row] == -1) {
        unsigned char* vt = sm + (size_t)stage * Sc::STAGE_TILES + 2 * TILE;
        const float* vp = a.vpad + p.h * HD + (lane & 1) * 16;
        const uint32_t off = sw64(row, (lane & 1) * 2);
        put16(vt, v_s, off, nullptr, vp);
        put16(vt, v_s, off ^ 16u, nullptr, vp + 8);
      }
      __syncwarp();
    }
    uint32_t ka[2][4], va[2][4];
    ldsm4(ka[0], k_s + la_off);
    ldsm4(ka[1], k_s + (la_off ^ 32u));
    ldsm4(va[0], v_s + la_off);
    ldsm4(va[1], v_s + (la_off ^ 32u));
    float dk[4][4], dv[4][4];
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
      dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
    }
    auto sdp = [&](int c, float (&st)[2][4], float (&dp)[2][4]) {
      uint32_t bq[2][4], bg[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        ldsm4(bq[n], q_s + (uint32_t)(c * 16 + n * 8) * 64u + lk_off);
        ldsm4(bg[n], g_s + (uint32_t)(c * 16 + n * 8) * 64u + lk_off);
        st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f;
        dp[n][0] = dp[n][1] = dp[n][2] = dp[n][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          mma16816(st[n], ka[ks], bq[n][2 * ks], bq[n][2 * ks + 1]);
          mma16816(dp[n], va[ks], bg[n][2 * ks], bg[n][2 * ks + 1]);
        }
    };
    auto sweep = [&](auto mask_c) {
      constexpr bool MASK = decltype(mask_c)::value;
#pragma unroll
      for (int c = 0; c < NT8 / 2; ++c) {         // 16 queries per step
        if ((padmask >> c) & 1u) continue;        // pad queries only: dS = 0, nothing to add anywhere
        float st[2][4], dp[2][4];
        sdp(c, st, dp);
        const uint32_t row16 = (uint32_t)(c * 16) * 64u;
        uint32_t bg0[4], bg2[4], bq0[4], bq2[4];
        ldsm4t(bg0, g_s + row16 + lv_off0);
        ldsm4t(bg2, g_s + row16 + lv_off2);
        ldsm4t(bq0, q_s + row16 + lv_off0);
        ldsm4t(bq2, q_s + row16 + lv_off2);
        uint32_t aP[4], aD[4];
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int qcol = c * 16 + n * 8 + 2 * tq;
          const float4 m = qm4[(c * 16 + n * 8) / 2 + tq];             // {lse, D} of queries qcol, qcol + 1
          int2 kq;
          if (WS % 2 == 0) kq.x = kof_pair<WS>(c * 16 + n * 8, tq), kq.y = kq.x + 4;
          else kq = *reinterpret_cast<const int2*>(kofk + qcol);
          const uint32_t aA0 = tabA + (uint32_t)kq.x, aB0 = tabB + (uint32_t)kq.x;
          const uint32_t aA1 = (WS % 2 == 0) ? aA0 + 4u : tabA + (uint32_t)kq.y;
          const uint32_t aB1 = (WS % 2 == 0) ? aB0 + 4u : tabB + (uint32_t)kq.y;
          float s2[4] = {fmaf(st[n][0], scale2, lds32(aA0)), fmaf(st[n][1], scale2, lds32(aA1)),
                         fmaf(st[n][2], scale2, lds32(aB0)), fmaf(st[n][3], scale2, lds32(aB1))};
          if (MASK) {
            const int2 rr = *reinterpret_cast<const int2*>(ridS + qcol);
            if (rr.x != ridA) s2[0] += kMaskLog2;
            if (rr.y != ridA) s2[1] += kMaskLog2;
            if (rr.x != ridB) s2[2] += kMaskLog2;
            if (rr.y != ridB) s2[3] += kMaskLog2;
          }
          float pv[4], ds[4];
          pv[0] = ex2f(fmaf(m.x, -kLog2e, s2[0]));
          pv[1] = ex2f(fmaf(m.z, -kLog2e, s2[1]));
          pv[2] = ex2f(fmaf(m.x, -kLog2e, s2[2]));
          pv[3] = ex2f(fmaf(m.z, -kLog2e, s2[3]));
          if (Cf::RAGGED) {
            if (rA >= N) pv[0] = pv[1] = 0.f;
            if (rB >= N) pv[2] = pv[3] = 0.f;
          }
          ds[0] = pv[0] * (dp[n][0] - m.y);
          ds[1] = pv[1] * (dp[n][1] - m.w);
          ds[2] = pv[2] * (dp[n][2] - m.y);
          ds[3] = pv[3] * (dp[n][3] - m.w);
          if (c * 2 + n < NREG) {
#pragma unroll
            for (int e = 0; e < 4; ++e) db[c * 2 + n < NREG ? c * 2 + n : 0][e] += ds[e];
          } else {
            float2* slot = dbs + ((c * 2 + n - NREG) * 2) * KTH;
            float2 u = slot[0], w = slot[KTH];
            u.x += ds[0]; u.y += ds[1]; w.x += ds[2]; w.y += ds[3];
            slot[0] = u; slot[KTH] = w;
          }
          aP[2 * n] = pack2(pv[0], pv[1]);
          aP[2 * n + 1] = pack2(pv[2], pv[3]);
          aD[2 * n] = pack2(ds[0], ds[1]);
          aD[2 * n + 1] = pack2(ds[2], ds[3]);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(panel_s + pw_offA + (uint32_t)(c * 16 + n * 8) * 2u), "r"(aD[2 * n]) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(panel_s + pw_offB + (uint32_t)(c * 16 + n * 8) * 2u), "r"(aD[2 * n + 1]) : "memory");
        }
        mma16816(dv[0], aP, bg0[0], bg0[1]);
        mma16816(dv[1], aP, bg0[2], bg0[3]);
        mma16816(dv[2], aP, bg2[0], bg2[1]);
        mma16816(dv[3], aP, bg2[2], bg2[3]);
        mma16816(dk[0], aD, bq0[0], bq0[1]);
        mma16816(dk[1], aD, bq0[2], bq0[3]);
        mma16816(dk[2], aD, bq2[0], bq2[1]);
        mma16816(dk[3], aD, bq2[2], bq2[3]);
      }
    };
    if (need_mask) sweep(std::true_type{}); else sweep(std::false_type{});

    // ---- dV, dK of the warp's 16 keys.  k_hat of row g / g + 8 sits in the A fragments at the accumulator's columns.
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int t = half ? tokB : tokA;
      float kh[4][2], dot = 0.f;
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        const uint32_t w = ka[dn >> 1][(dn & 1) * 2 + half];
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
        kh[dn][0] = f.x; kh[dn][1] = f.y;
        dot = fmaf(dk[dn][2 * half], f.x, dot);
        dot = fmaf(dk[dn][2 * half + 1], f.y, dot);
      }
      dot = quad_sum(dot) * sc;
      if (t >= 0) {
        const float invn = innorm[stage * NP + (half ? rB : rA)].y;          // 1 / ||k||
        uint32_t* dkd = reinterpret_cast<uint32_t*>(a.dqkv + (int64_t)t * C3 + a.C + p.h * HD) + tq;
        uint32_t* dvd = reinterpret_cast<uint32_t*>(a.dqkv + (int64_t)t * C3 + 2 * a.C + p.h * HD) + tq;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          dkd[dn * 4] = pack2((dk[dn][2 * half] * sc - kh[dn][0] * dot) * invn, (dk[dn][2 * half + 1] * sc - kh[dn][1] * dot) * invn);
          dvd[dn * 4] = pack2(dv[dn][2 * half], dv[dn][2 * half + 1]);
        }
        if (a.dcol) {
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            dvs[(dn * 2) * KTH] += dv[dn][2 * half];
            dvs[(dn * 2 + 1) * KTH] += dv[dn][2 * half + 1];
          }
        }
      } else if (t == -1) {
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          dvps[(dn * 2) * KTH] += dv[dn][2 * half];
          dvps[(dn * 2 + 1) * KTH] += dv[dn][2 * half + 1];
        }
      }
    }
    // this warp's part of the panel is written and it no longer reads stage `stage`
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&pfull[stage]);
  }
}

template <int WS>
int launch_mma_bwd_spec(const MmaArgs& a, cudaStream_t st) {
  using Sc = SCfg<WS>;
  const size_t smem = Sc::SMEM;
  BSW_REQUIRE(smem <= 227 * 1024, "attn_bwd(mma): window %dx%d needs %zu bytes of shared memory", WS, WS, smem);
  BSW_CUDA(cudaFuncSetAttribute(attn_mma_bwd_spec_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = sm_count();
  if (grid > a.nitems) grid = a.nitems;
  CUtensorMap tm_qkv, tm_dout;                   // window boxes of qkv [B][H][W][3C] and dout [B][H][W][C]
  memset(&tm_qkv, 0, sizeof(tm_qkv));
  memset(&tm_dout, 0, sizeof(tm_dout));
  const int use_tma = make_tmap_window_bf16(&tm_qkv, a.qkv, a.g.B, a.g.H, a.g.W, 3 * a.C, WS) == B200SWIN_OK &&
                      make_tmap_window_bf16(&tm_dout, a.dout, a.g.B, a.g.H, a.g.W, a.C, WS) == B200SWIN_OK;
  attn_mma_bwd_spec_kernel<WS><<<(unsigned)grid, Sc::THREADS, smem, st>>>(a, tm_qkv, tm_dout, use_tma);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

int fill_mma_args(MmaArgs* a, int B, int H, int W, int C, int nH, int ws, int shift) {
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && nH > 0, "attn(mma): bad dimension");
  BSW_REQUIRE(C == nH * HD, "attn(mma): head_dim must be 32 (C=%d, nH=%d)", C, nH);
  BSW_REQUIRE(shift >= 0 && shift < ws, "attn(mma): bad shift");
  BSW_REQUIRE((int64_t)B * H * W < (1ll << 29), "attn(mma): too many tokens");
  make_geom(&a->g, B, H, W, ws, shift);
  a->C = C; a->nH = nH;
  a->nwin = (int64_t)B * a->g.nWh * a->g.nWw;
  a->nitems = a->nwin * nH;
  BSW_REQUIRE(a->nwin < (1ll << 31), "attn(mma): too many windows");
  return B200SWIN_OK;
}
}  // namespace

bool attn_fwd_mma_supported(int ws) { return ws == 4 || ws == 6 || ws == 7 || ws == 8 || ws == 12; }

int attn_fwd_mma(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                 const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st) {
  MmaArgs a = {};
  int rc = fill_mma_args(&a, B, H, W, C, nH, ws, shift);
  if (rc) return rc;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.out_lo = (__nv_bfloat16*)out_lo; a.lse = lse;
  a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  switch (ws) {
    case 4: return launch_mma_fwd<4>(a, st);
    case 6: return launch_mma_fwd<6>(a, st);
    case 7: return launch_mma_fwd<7>(a, st);
    case 8: return launch_mma_fwd<8>(a, st);
    case 12: return launch_mma_fwd<12>(a, st);
    default: break;
  }
  set_error("attn_fwd(mma): window %dx%d not instantiated", ws, ws);
  return B200SWIN_EINVAL;
}

bool attn_bwd_mma_supported(int ws) { return ws == 4 || ws == 6 || ws == 7 || ws == 8 || ws == 12; }

// D = <dO, O> per (token, head) (attn_bwd_ws.cu)
int attn_bwd_prep(const void* dout, const void* out, const void* out_lo, float* dvec, int64_t n, cudaStream_t st);

size_t attn_bwd_mma_workspace_bytes(int B, int H, int W, int nH) { return (size_t)B * H * W * nH * sizeof(float); }

int attn_bwd_mma(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse, const float* inv_norm,
                 const float* table16, const float* scale, const float* qpad, const float* vpad, void* dqkv,
                 float* dtable16, float* dscale, float* dvpad, float* dcol, void* workspace, int B, int H, int W, int C,
                 int nH, int ws, int shift, bool spec, cudaStream_t st) {
  BSW_REQUIRE(workspace, "attn_bwd(mma): workspace for D = <dO, O> missing");
  BSW_REQUIRE(!dcol || (spec && ws == 12), "attn_bwd(mma): column sums come only from the warp-specialised 12x12 kernel");
  MmaArgs a = {};
  int rc = fill_mma_args(&a, B, H, W, C, nH, ws, shift);
  if (rc) return rc;
  a.qkv = (const __nv_bfloat16*)qkv; a.dout = (const __nv_bfloat16*)dout; a.lse = const_cast<float*>(lse);
  a.dvec = (const float*)workspace; a.inv_norm = inv_norm; a.table16 = table16; a.scale = scale; a.qpad = qpad;
  a.vpad = vpad; a.dqkv = (__nv_bfloat16*)dqkv; a.dtable16 = dtable16; a.dscale = dscale; a.dvpad = dvpad; a.dcol = dcol;
  rc = attn_bwd_prep(dout, out, out_lo, (float*)workspace, (int64_t)B * H * W * nH, st);
  if (rc) return rc;
  switch (ws) {
    case 4: return launch_mma_bwd<4>(a, st);
    case 6: return launch_mma_bwd<6>(a, st);
    case 7: return launch_mma_bwd<7>(a, st);
    case 8: return launch_mma_bwd<8>(a, st);
    case 12: return spec ? launch_mma_bwd_spec<12>(a, st) : launch_mma_bwd<12>(a, st);
    default: break;
  }
  set_error("attn_bwd(mma): window %dx%d not instantiated", ws, ws);
  return B200SWIN_EINVAL;
}

}  // namespace b200swin
