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
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
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
// of the next two items are in flight (cp.async, three stages) while the warps work on the current one, one CTA barrier
// per item.  Two CTAs per SM.
constexpr int kFwdStages = 3;

template <int WS>
__global__ void __launch_bounds__(MCfg<WS>::THREADS, (MCfg<WS>::THREADS <= 320 ? 2 : 1))
attn_mma_fwd_kernel(const __grid_constant__ MmaArgs a) {
  using Cf = MCfg<WS>;
  constexpr int N = Cf::N, NP = Cf::NP, TW = Cf::TW, CHN = Cf::CHN;
  constexpr uint32_t TILE = Cf::TILE, STAGE = 3 * TILE;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 127u) & ~127u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  unsigned char* tiles = sm;                                              // [stages][q | k | v]
  int* tokm = reinterpret_cast<int*>(tiles + kFwdStages * STAGE);         // [stages][NP] source token of a row
  int* ridm = tokm + kFwdStages * NP;                                     // [stages][NP] shift-mask region id
  int* kofk = ridm + kFwdStages * NP;                                     // [NP] byte offset 4 (y TW + x) of a window row
  float* tab = reinterpret_cast<float*>(kofk + NP);                       // [NTAB] bias table of the head, log2 units
  const uint32_t tiles_s = base_u32;

  const WinGeom& g = a.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int64_t per = a.nitems / gridDim.x, rem = a.nitems % gridDim.x;
  const int64_t it0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int nit = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));
  const int C3 = 3 * a.C;

  for (int r = tid; r < NP; r += Cf::THREADS) kofk[r] = r < N ? 4 * ((r / WS) * TW + (r % WS)) : 0;

  // a thread copies the same two (row, 16-byte column) slots of every item: rows (tid >> 2) and (tid >> 2) + THREADS / 4
  static_assert(NP * 4 == 2 * Cf::THREADS, "two slots per thread");
  const int pc = tid & 3, prow0 = tid >> 2, prow1 = prow0 + Cf::THREADS / 4;
  const int py0 = prow0 / WS, px0 = prow0 - py0 * WS, py1 = prow1 / WS, px1 = prow1 - py1 * WS;
  const uint32_t poff0 = sw64(prow0, pc), poff1 = sw64(prow1, pc);
  ItemPos pp = item_pos(a, it0);                 // cursor of the prefetch stream
  auto prefetch = [&](int i) {
    const int stage = i % kFwdStages;
    unsigned char* q0 = tiles + (size_t)stage * STAGE;
    const uint32_t q0_s = tiles_s + (uint32_t)stage * STAGE;
    const float* qp = a.qpad ? a.qpad + pp.h * HD + pc * 8 : nullptr;
    const float* vp = a.vpad ? a.vpad + pp.h * HD + pc * 8 : nullptr;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int row = k ? prow1 : prow0;
      const uint32_t off = k ? poff1 : poff0;
      int region = 0, t = -2;
      if (!Cf::RAGGED || row < N) t = row_token<WS>(g, pp, k ? py1 : py0, k ? px1 : px0, &region);
      if (pc == 0) {
        tokm[stage * NP + row] = t;
        ridm[stage * NP + row] = region;
      }
      const __nv_bfloat16* src = t >= 0 ? a.qkv + (int64_t)t * C3 + pp.h * HD + pc * 8 : nullptr;
      put16(q0, q0_s, off, src, t == -1 ? qp : nullptr);
      put16(q0 + TILE, q0_s + TILE, off, t >= 0 ? src + a.C : nullptr, nullptr);
      put16(q0 + 2 * TILE, q0_s + 2 * TILE, off, t >= 0 ? src + 2 * a.C : nullptr, t == -1 ? vp : nullptr);
    }
    item_next(a, pp);
  };

  if (nit > 0) prefetch(0);
  ptx::cp_async_commit();
  if (nit > 1) prefetch(1);
  ptx::cp_async_commit();

  int cur_h = -1;
  float scale2 = 0.f;
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
    __syncthreads();                              // ... everybody's; and every warp is done with item i - 1
    if (i + 2 < nit) prefetch(i + 2);             // into the stage item i - 1 has just released
    ptx::cp_async_commit();
    if (p.h != cur_h) {                           // CTA-uniform
      for (int t = tid; t < Cf::NTAB; t += Cf::THREADS) tab[t] = a.table16[(int64_t)t * a.nH + p.h] * kLog2e;
      scale2 = a.scale[p.h] * kLog2e;
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
    float o[4][4];
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;

    auto chunk = [&](int c, auto mask_c) {
      constexpr bool MASK = decltype(mask_c)::value;
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
        mxA = fmaxf(mxA, fmaxf(s[n][0], s[n][1]));
        mxB = fmaxf(mxB, fmaxf(s[n][2], s[n][3]));
      }
      // lazy rescale: the reference maximum of a row only moves when a chunk exceeds it by more than 2^8 (P <= 256 is
      // harmless in fp32 / bf16), so after the first chunk the accumulators are almost never touched
      mxA = quad_max(mxA);
      mxB = quad_max(mxB);
      const float mnA = mxA > mA + kLazy ? mxA : mA, mnB = mxB > mB + kLazy ? mxB : mB;
      if (__any_sync(0xffffffffu, mnA != mA || mnB != mB)) {
        const float cA = ex2f(mA - mnA), cB = ex2f(mB - mnB);
        lA *= cA; lB *= cB;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          o[dn][0] *= cA; o[dn][1] *= cA; o[dn][2] *= cB; o[dn][3] *= cB;
        }
        mA = mnA; mB = mnB;
      }
#pragma unroll
      for (int n = 0; n < CHN; ++n) {
        s[n][0] = ex2f(s[n][0] - mnA);
        s[n][1] = ex2f(s[n][1] - mnA);
        s[n][2] = ex2f(s[n][2] - mnB);
        s[n][3] = ex2f(s[n][3] - mnB);
        lA += s[n][0] + s[n][1];
        lB += s[n][2] + s[n][3];
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
      }
    };
    if (need_mask) {
#pragma unroll 1
      for (int c = 0; c < Cf::NCH; ++c) chunk(c, std::true_type{});
    } else {
#pragma unroll 1
      for (int c = 0; c < Cf::NCH; ++c) chunk(c, std::false_type{});
    }

    // ---- epilogue: normalise, store O (+ its bf16 residual) and the row's log-sum-exp
    lA = quad_sum(lA);
    lB = quad_sum(lB);
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
  return 128 + (size_t)kFwdStages * 3 * Cf::TILE + 2 * (size_t)kFwdStages * Cf::NP * 4 + (size_t)Cf::NTAB * 4 + (size_t)Cf::NP * 4;
}

template <int WS>
int launch_mma_fwd(const MmaArgs& a, cudaStream_t st) {
  using Cf = MCfg<WS>;
  const size_t smem = mma_fwd_smem<WS>();
  BSW_CUDA(cudaFuncSetAttribute(attn_mma_fwd_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BSW_CUDA(cudaFuncSetAttribute(attn_mma_fwd_kernel<WS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int occ = (int)((227 * 1024) / (smem + 1024));
  const int occ_threads = 2048 / Cf::THREADS;
  if (occ > occ_threads) occ = occ_threads;
  if (occ > 2) occ = 2;                               // __launch_bounds__(THREADS, 2)
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > a.nitems) grid = a.nitems;
  attn_mma_fwd_kernel<WS><<<(unsigned)grid, Cf::THREADS, smem, st>>>(a);
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

bool attn_fwd_mma_supported(int ws) { return ws == 12; }

int attn_fwd_mma(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                 const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st) {
  MmaArgs a = {};
  int rc = fill_mma_args(&a, B, H, W, C, nH, ws, shift);
  if (rc) return rc;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.out_lo = (__nv_bfloat16*)out_lo; a.lse = lse;
  a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  switch (ws) {
    case 12: return launch_mma_fwd<12>(a, st);
    default: break;
  }
  set_error("attn_fwd(mma): window %dx%d not instantiated", ws, ws);
  return B200SWIN_EINVAL;
}

}  // namespace b200swin
