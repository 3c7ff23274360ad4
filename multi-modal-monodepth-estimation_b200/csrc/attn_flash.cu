// KV-blocked ("flash") attention core on tcgen05 for ANY window size up to 32x32 (bf16 storage): the tensor-core path of
// every window the single-tile kernels of attn_fwd_ws.cu / attn_bwd_ws.cu do not cover -- 16x16, 24x24 and the
// reference's default 30x30 (configs/config.yaml:55; N = 256 / 576 / 900 tokens per window), and the odd sizes.
//
// Same math as the other attention kernels (models/swin_transformer_v2.py:295-328 with the pad / roll / partition /
// reverse / crop of :429-463 and the shift mask of :874-892 as address math; backward per SURVEY.md appendix A).
//
// One kernel template, three modes.  In every mode a CTA of 128 threads owns 128 "stationary" rows of one
// (window, head) -- one row per thread = one TMEM lane -- and streams the other side of the window through a
// double-buffered cp.async ring in blocks of 64 tokens (the co-resident CTAs hide the rest of the latency):
//   FWD  stationary = queries, streamed = keys:    S = Q K^T -> online softmax -> O += P V        (P: TMEM A operand)
//   DQ   stationary = queries, streamed = keys:    S, dP = dO V^T -> dS -> dQ += dS K            (dS: TMEM A operand)
//                                                  + the bias-table and temperature gradients
//   DKV  stationary = keys,    streamed = queries: S^T = K Q^T, dP^T = V dO^T -> P^T, dS^T ->
//                                                  dV += P^T dO, dK += dS^T Q                    (both TMEM A operands)
// The backward is two passes (DQ, DKV) that each recompute S and dP on the tensor cores: with the transposed
// formulation of DKV every contraction has its left operand in TMEM exactly where the thread that produced it wrote
// it -- no shared-memory panels, no transposes, no atomics on dQ -- and every pass has the shape of the forward.
// Per-block products (O, dQ, dK, dV of one 64-token block) come back from TMEM and are accumulated in REGISTERS, which
// makes the online-softmax rescale a register multiply.
//
// A CTA needs 128 TMEM columns and ~45-110 KB of shared memory, so up to FOUR CTAs share an SM: the latency of the
// MMA -> softmax -> MMA chain of one CTA is hidden by the other three (the single-tile kernels hide it by warp
// specialisation inside one CTA instead).  Work units are (head, window, row tile), head-major, a contiguous range per
// CTA, so the bias table is rebuilt and the gradient sums are flushed only when the head changes.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "wingeom.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

namespace {
constexpr int HD = 32;
constexpr int kThreads = 128;
constexpr int NSTAGE = 2;
constexpr int MODE_FWD = 0, MODE_DQ = 1, MODE_DKV = 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMaskLog2 = -100.0f * 1.4426950408889634f;
constexpr uint32_t kSw64 = 4;          // UMMA layout type SWIZZLE_64B
constexpr uint32_t kXTile = 128 * 64;  // stationary operand tile [128][64 B]
// TMEM columns of a CTA (128 allocated): S | dP, 64 fp32 columns each.  The bf16 A operand a thread derives from its
// row of S (dP) overwrites the first 32 columns of that row; the block product lands in the last 32.
constexpr uint32_t S_COL = 0, DP_COL = 64, RES_OFF = 32;

struct FlArgs {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* dout;
  __nv_bfloat16* out;
  __nv_bfloat16* out_lo;    // optional bf16 residual O - bf16(O) (see attn_fwd_ws.cu)
  __nv_bfloat16* dqkv;
  float* lse;
  const float* dvec;        // [B*H*W, nH]  D = <dO, O>
  const float* inv_norm;
  const float* table16;
  const float* scale;
  const float* qpad;
  const float* vpad;
  float* dtable16;
  float* dscale;
  float* dvpad;
  WinGeom g;
  int C, nH, N, ntiles, rpt, kb, nkb, ntab, nmeta;
  int ntabc;                  // DQ: entries of a warp's COMPACT gradient table (see the kernel), <= ntab
  int ntabq;                  // DQ: entries of the CTA's compact copy of the bias table (rows a 128-row tile can touch), <= ntab
  int64_t nwin, nunits;
};

__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// shared-memory accesses by 32-bit shared-space address (generic 64-bit pointer arithmetic per element costs two adds)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 16 bytes of one row of an operand tile: global (a real token), a pad value (fp32 -> bf16) or zeros
__device__ __forceinline__ void put16(unsigned char* tile, uint32_t tile_s, uint32_t off, const __nv_bfloat16* src,
                                      const float* padv) {
  if (src) {
    ptx::cp_async_16(tile_s + off, src);
  } else {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (padv) v = make_uint4(pack_bf16(padv[0], padv[1]), pack_bf16(padv[2], padv[3]), pack_bf16(padv[4], padv[5]),
                             pack_bf16(padv[6], padv[7]));
    *reinterpret_cast<uint4*>(tile + off) = v;
  }
}

// F.normalize backward of one gradient row held in registers: d = (g*sc - x_hat <g*sc, x_hat>) * inv_norm.
// invn == 0 stands for "q / k were not normalised" (attn_type='normal', no inv_norm tensor): d = g*sc.
__device__ __forceinline__ void normalize_bwd_store(const float (&gacc)[HD], const unsigned char* tile, int r, float sc,
                                                    float invn, __nv_bfloat16* dst) {
  float xh[HD];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint4 w = *reinterpret_cast<const uint4*>(tile + sw64_off(r, c));
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
      xh[c * 8 + 2 * e] = f.x;
      xh[c * 8 + 2 * e + 1] = f.y;
    }
  }
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < HD; ++c) dot = fmaf(gacc[c], xh[c], dot);
  dot *= sc;
  if (invn == 0.f) { dot = 0.f; invn = 1.f; }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      pk[e] = pack_bf16((gacc[c * 8 + 2 * e] * sc - xh[c * 8 + 2 * e] * dot) * invn,
                        (gacc[c * 8 + 2 * e + 1] * sc - xh[c * 8 + 2 * e + 1] * dot) * invn);
    reinterpret_cast<uint4*>(dst)[c] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// KB = streamed tokens per block (64; 48 for 12x12 windows, whose 144 tokens are exactly three blocks)
template <int MODE, int KB>
__global__ void __launch_bounds__(kThreads, 4)
attn_flash_kernel(const __grid_constant__ FlArgs a) {
  constexpr uint32_t kYTile = KB * 64;             // streamed operand tile [KB][64 B]
  constexpr uint32_t kStage = 2 * kYTile;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar_mma;
  __shared__ uint32_t tmem_slot;
  __shared__ float red_s[4];

  const WinGeom& g = a.g;
  const int ws = g.ws, N = a.N, TW = 2 * ws - 1;
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  unsigned char* xt = sm;                                          // two stationary tiles
  unsigned char* ring = xt + 2 * kXTile;                           // NSTAGE x (two streamed tiles)
  float* blk_lse = reinterpret_cast<float*>(ring + NSTAGE * kStage);   // [NSTAGE][KB] (DKV: per streamed query)
  float* blk_d = blk_lse + NSTAGE * KB;
  int* tok = reinterpret_cast<int*>(blk_d + NSTAGE * KB);          // [nmeta] flat token index, -1 pad, -2 beyond the window
  int* kof = tok + a.nmeta;                                        // [nmeta] byte offset 4 (y TW + x) into the bias table
  unsigned char* rid = reinterpret_cast<unsigned char*>(kof + a.nmeta);   // [nmeta] shift-mask region id | beyond the window << 7
  // bias table of the head, log2 units: [ntab]; DQ: only the [ntabq] entries of the table rows the CTA's row tile can touch
  float* tab = reinterpret_cast<float*>(rid + a.nmeta);
  const int ntab_s = MODE == MODE_DQ ? a.ntabq : a.ntab;
  // DQ: [4 warps][ntabc] private gradient sums of the bias table.  A warp's 32 query rows span at most 31 / ws + 2 window
  // rows, so it only ever touches ws - 1 + that many of the table's 2 ws - 1 rows (dy = y_query - y_key): the private
  // tables hold just those rows, indexed from the warp's first window row.  That is 54 % of the full table at 24 / 30
  // windows and makes room for one more CTA per SM (30x30: two instead of one).  The price: the units of a CTA run in
  // (head, row tile, window) order -- a warp's rows stay put while the windows stream by -- and the tables are flushed
  // when the (head, tile) pair changes.
  float* dtab = tab + ntab_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t per = a.nunits / gridDim.x, rem = a.nunits % gridDim.x;
  const int64_t u0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int nu = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));
  const int nW = g.nWh * g.nWw;
  const int C3 = 3 * a.C;

  if (tid == 0) {
    ptx::mbar_init(&bar_mma, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  if (MODE == MODE_DQ)
    for (int i = tid; i < 4 * a.ntabc; i += kThreads) dtab[i] = 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
  uint32_t ph = 0;                                                 // phase of bar_mma (every thread waits every phase)

  const uint32_t xt_s = ptx::smem_u32(xt), ring_s = ptx::smem_u32(ring);
  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, KB, 0, 0);        // [128 x 32] . [KB x 32]^T
  constexpr uint32_t idesc_r = ptx::make_idesc_bf16(128, HD, 0, 1);        // [128 x 64](TMEM) . [64 x 32] (MN-major B)
  const uint64_t desc_k = ptx::make_smem_desc(0, 16, 512, kSw64);          // K-major tile of 64 B rows
  const uint64_t desc_mn = ptx::make_smem_desc(0, 512, 512, kSw64);        // the same bytes read MN-major

  int cur_h = -1, cur_tile = -1;
  int64_t cur_win = -1;
  float sc = 0.f, scale2 = 0.f, dsc = 0.f;
  float* dtabw = dtab + warp * a.ntabc;
  bool need_mask = false;
  const bool compact = a.ntabc < a.ntab;
  // first table row (dy index) a warp's compact table holds while the CTA works on row tile `tile`
  auto first_dy = [&](int tile, int w) { return compact ? (tile * a.rpt + w * 32) / ws : 0; };
  const bool compact_q = MODE == MODE_DQ && a.ntabq < a.ntab;
  auto first_dy_cta = [&](int tile) { return compact_q ? (tile * a.rpt) / ws : 0; };

  auto flush_head = [&](int h, int tile) {
    // DQ: gradient sums of head h -> global (one atomic per touched table entry and warp)
    __syncthreads();
    for (int w = 0; w < 4; ++w) {
      const int r0 = first_dy(tile, w) * TW;
      float* tw_ = dtab + w * a.ntabc;
      for (int r = tid; r < a.ntabc; r += kThreads) {
        const float v = tw_[r];
        tw_[r] = 0.f;
        if (v != 0.f && r0 + r < a.ntab) atomicAdd(a.dtable16 + (int64_t)(r0 + r) * a.nH + h, v);
      }
    }
    const float s = warp_sum(dsc);
    dsc = 0.f;
    if (lane == 0) red_s[warp] = s;
    __syncthreads();
    if (tid == 0) atomicAdd(a.dscale + h, (red_s[0] + red_s[1]) + (red_s[2] + red_s[3]));
  };

#pragma unroll 1
  for (int ui = 0; ui < nu; ++ui) {
    const int64_t u = u0 + ui;
    int tile, h;
    int64_t win;
    if (MODE == MODE_DQ) {                           // (head, tile, window): see the compact gradient tables above
      win = u % a.nwin;
      const int64_t it = u / a.nwin;
      tile = (int)(it % a.ntiles);
      h = (int)(it / a.ntiles);
    } else {                                         // (head, window, tile): the window's row tables serve all its tiles
      tile = (int)(u % a.ntiles);
      const int64_t iw = u / a.ntiles;
      win = iw % a.nwin;
      h = (int)(iw / a.nwin);
    }
    const int b = (int)(win / nW);
    const int wrem = (int)(win - (int64_t)b * nW);
    const int wh = wrem / g.nWw, ww = wrem - wh * g.nWw;

    // ---- per-head / per-window tables (every thread has finished the previous unit: its last MMA wait is behind it)
    if (h != cur_h || win != cur_win || (MODE == MODE_DQ && tile != cur_tile)) {
      if (MODE == MODE_DQ && cur_h >= 0 && (h != cur_h || tile != cur_tile)) flush_head(cur_h, cur_tile);
      __syncthreads();
      if (h != cur_h || (compact_q && tile != cur_tile)) {
        const int t0 = first_dy_cta(tile) * TW;
        for (int t = tid; t < ntab_s; t += kThreads)
          tab[t] = t0 + t < a.ntab ? a.table16[(int64_t)(t0 + t) * a.nH + h] * kLog2e : 0.f;
        sc = a.scale[h];
        scale2 = sc * kLog2e;
      }
      if (win != cur_win) {
        need_mask = g.shift > 0 && (wh == g.nWh - 1 || ww == g.nWw - 1);
        for (int r = tid; r < a.nmeta; r += kThreads) {
          int t = -2, ko = 0, rg = 1 << 7;
          if (r < N) {
            const int y = r / ws, x = r - y * ws;
            const int si = wh * ws + y, sj = ww * ws + x;
            int i = si + g.shift; if (i >= g.Hp) i -= g.Hp;
            int j = sj + g.shift; if (j >= g.Wp) j -= g.Wp;
            t = (i < g.H && j < g.W) ? (b * g.H + i) * g.W + j : -1;
            const int region = g.shift > 0 ? 3 * region_1d(si, g.Hp, ws, g.shift) + region_1d(sj, g.Wp, ws, g.shift) : 0;
            ko = 4 * (y * TW + x);
            rg = region;
          }
          tok[r] = t;
          kof[r] = ko;
          rid[r] = (unsigned char)rg;
        }
      }
      cur_h = h;
      cur_win = win;
      cur_tile = tile;
      __syncthreads();
    }

    // ---- this thread's stationary row
    const int r_loc = tid;
    const int r_st = tile * a.rpt + r_loc;
    const bool row_valid = r_loc < a.rpt && r_st < N;
    // a warp without a single row (short last tile) only keeps the barriers company
    const bool warp_live = warp * 32 < a.rpt && tile * a.rpt + warp * 32 < N;
    const int t_st = row_valid ? tok[r_st] : -2;
    const int kof_st = kof[row_valid ? r_st : 0], rid_st = rid[row_valid ? r_st : 0] & 0x7f;
    // bias index = koff(query) + (ws-1)(TW+1) - koff(key).  FWD / DQ: this row is the query, `tabq - kof[key]` is the
    // entry; DKV: this row is the key, `tabq + kof[query]`.
    const int off_st = MODE == MODE_DKV ? 4 * (ws - 1) * (TW + 1) - kof_st : 4 * (ws - 1) * (TW + 1) + kof_st;
    // (DQ: the CTA's copy of the table starts at the first table row its tile can touch; a row beyond the window reads as
    // if it sat at the start of that row)
    const uint32_t tabq = ptx::smem_u32(tab) +
                          (uint32_t)((!compact_q || row_valid) ? off_st - 4 * first_dy_cta(tile) * TW : 4 * (ws - 1) * (TW + 1));
    // DQ: the same entry of this warp's gradient sums, counted from the warp's first table row; a row beyond the window
    // (short last tile) never stores, and reads as if it sat at the start of the warp's first window row
    const int dy0 = first_dy(tile, warp);
    const uint32_t dtabq = ptx::smem_u32(dtabw) +
                           (uint32_t)(row_valid ? off_st - 4 * dy0 * TW : 4 * (ws - 1) * (TW + 1));
    float lse2_st = INFINITY, d_st = 0.f;                           // DQ: per-query constants
    if (MODE == MODE_DQ && row_valid) {
      lse2_st = a.lse[(win * a.nH + h) * N + r_st] * kLog2e;
      if (t_st >= 0) d_st = a.dvec[(int64_t)t_st * a.nH + h];
    }

    // ---- gather: the stationary tiles and the first two streamed blocks
    auto gather_block = [&](int kb) {
      const int stage = kb % NSTAGE;
      unsigned char* y0 = ring + (size_t)stage * kStage;
      const uint32_t y0_s = ring_s + (uint32_t)stage * kStage;
      for (int idx = tid; idx < KB * 4; idx += kThreads) {
        const int jj = idx >> 2, c = idx & 3;
        const int j = kb * KB + jj;
        const int t = tok[j];
        const uint32_t off = sw64_off(jj, c);
        const __nv_bfloat16* base = t >= 0 ? a.qkv + (int64_t)t * C3 + h * HD + c * 8 : nullptr;
        if (MODE == MODE_DKV) {
          // streamed queries: q_hat (pad: normalised q_bias) | dO (pad: 0)
          put16(y0, y0_s, off, base, (t == -1 && a.qpad) ? a.qpad + h * HD + c * 8 : nullptr);
          put16(y0 + kYTile, y0_s + kYTile, off, t >= 0 ? a.dout + (int64_t)t * a.C + h * HD + c * 8 : nullptr, nullptr);
        } else {
          // streamed keys: k_hat (pad: 0) | v (pad: v_bias)
          put16(y0, y0_s, off, t >= 0 ? base + a.C : nullptr, nullptr);
          put16(y0 + kYTile, y0_s + kYTile, off, t >= 0 ? base + 2 * a.C : nullptr,
                (t == -1 && a.vpad) ? a.vpad + h * HD + c * 8 : nullptr);
        }
      }
      if (MODE == MODE_DKV && tid < KB) {
        const int j = kb * KB + tid;
        const int t = tok[j];
        blk_lse[stage * KB + tid] = j < N ? a.lse[(win * a.nH + h) * N + j] * kLog2e : INFINITY;
        blk_d[stage * KB + tid] = t >= 0 ? a.dvec[(int64_t)t * a.nH + h] : 0.f;
      }
    };
    for (int idx = tid; idx < 128 * 4; idx += kThreads) {
      const int rr = idx >> 2, c = idx & 3;
      const int r = tile * a.rpt + rr;
      const int t = (rr < a.rpt && r < N) ? tok[r] : -2;
      const uint32_t off = sw64_off(rr, c);
      const __nv_bfloat16* base = t >= 0 ? a.qkv + (int64_t)t * C3 + h * HD + c * 8 : nullptr;
      if (MODE == MODE_DKV) {
        put16(xt, xt_s, off, t >= 0 ? base + a.C : nullptr, nullptr);                                   // k_hat
        put16(xt + kXTile, xt_s + kXTile, off, t >= 0 ? base + 2 * a.C : nullptr,
              (t == -1 && a.vpad) ? a.vpad + h * HD + c * 8 : nullptr);                                 // v
      } else {
        put16(xt, xt_s, off, base, (t == -1 && a.qpad) ? a.qpad + h * HD + c * 8 : nullptr);           // q_hat
        if (MODE == MODE_DQ)
          put16(xt + kXTile, xt_s + kXTile, off, t >= 0 ? a.dout + (int64_t)t * a.C + h * HD + c * 8 : nullptr, nullptr);
      }
    }
    gather_block(0);
    ptx::cp_async_commit();
    if (NSTAGE > 2) {
      if (a.nkb > 1) gather_block(1);
      ptx::cp_async_commit();
    }

    float acc0[HD];                     // FWD: O, DQ: dQ, DKV: dV
    float acc1[HD];                     // DKV: dK (dead in the other modes)
#pragma unroll
    for (int c = 0; c < HD; ++c) { acc0[c] = 0.f; acc1[c] = 0.f; }
    float m_run = -INFINITY, l_run = 0.f;

    // ------------------------------------------------------------------------------------------------ FWD pipeline
    // TMEM reads are the scarce resource of this kernel (a tcgen05.ld moves ~64 B per clock and SM: re-reading S for a
    // second softmax pass, or fetching the block product every block, costs more than the exponentials).  So:
    //  * S is read ONCE: the 48 logits of a row live in registers between the maximum and the exponentials;
    //  * O accumulates in TMEM across the blocks of a row tile (tcgen05.mma accumulate) and is read once at the end.  The
    //    reference maximum m_ref of a row only moves when a block exceeds it by more than 2^8 ("lazy rescale": P <= 256
    //    is harmless in bf16 / fp32); then the row's O and l are rescaled in place -- rare after the first block;
    //  * S is double-buffered: S = Q K^T of block kb + 1 is issued together with O += P V of block kb, one barrier and
    //    one MMA round trip per block.
    // TMEM columns: S buffer 0 | S buffer 1 | O = 48 + 48 + 32 (the forward always streams 48-token blocks).
    if constexpr (MODE == MODE_FWD) {
      static_assert(MODE != MODE_FWD || KB == 48, "the forward streams 48-token blocks");
      constexpr uint32_t O_COL = 96;
      constexpr float kLazy = 8.0f;
      const uint64_t ax = desc_k + (xt_s >> 4);
      ptx::cp_async_wait<0>();                         // the stationary tile and block 0 have landed
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        ptx::tc_fence_after();
        const uint64_t by = desc_k + (ring_s >> 4);
        ptx::mma_bf16_ss(tmem_base, ax, by, idesc_s, 0u);
        ptx::mma_bf16_ss(tmem_base, ax + 2, by + 2, idesc_s, 1u);
        ptx::mma_commit(&bar_mma);
      }
#pragma unroll 1
      for (int kb = 0; kb < a.nkb; ++kb) {
        const uint32_t sb = (uint32_t)(kb & 1) * KB;
        ptx::mbar_wait(&bar_mma, ph);                  // S of block kb exists, O holds the blocks before kb
        ph ^= 1;
        ptx::tc_fence_after();
        // the other stage was last read by the MMAs of block kb - 1, which have just completed
        if (kb + 1 < a.nkb) gather_block(kb + 1);
        ptx::cp_async_commit();
        if (warp_live) {
          const bool tail_blk = (kb + 1) * KB > N;
          const bool general = need_mask || tail_blk;
          const int4* kof4 = reinterpret_cast<const int4*>(kof + kb * KB);
          const uint32_t* rid4 = reinterpret_cast<const uint32_t*>(rid + kb * KB);     // four keys per word
          uint32_t sv[KB];
#pragma unroll
          for (int c = 0; c < KB / 16; ++c) tmem_ld16(t_row + sb + c * 16, &sv[c * 16]);
          ptx::tmem_ld_wait();
          float mx = -INFINITY;
          auto logits = [&](auto gen_c) {
            constexpr bool GEN = decltype(gen_c)::value;
#pragma unroll
            for (int j4 = 0; j4 < KB / 4; ++j4) {
              // keep the loads of a group with their group: hoisting all twelve 16-byte key records to the top costs 48
              // registers the 48 logits need
              if ((j4 & 1) == 0) asm volatile("" ::: "memory");
              const int4 kk = kof4[j4];
              const int kj[4] = {kk.x, kk.y, kk.z, kk.w};
              int rj[4] = {0, 0, 0, 0};
              if (GEN) {
                const uint32_t rr = rid4[j4];
                rj[0] = rr & 0xff; rj[1] = (rr >> 8) & 0xff; rj[2] = (rr >> 16) & 0xff; rj[3] = rr >> 24;
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float s2 = fmaf(__uint_as_float(sv[j4 * 4 + k]), scale2, lds_f32(tabq - (uint32_t)kj[k]));
                if (GEN) {
                  if (need_mask && (rj[k] & 0x7f) != rid_st) s2 += kMaskLog2;
                  if (rj[k] >> 7) s2 = -INFINITY;
                }
                mx = fmaxf(mx, s2);
                sv[j4 * 4 + k] = __float_as_uint(s2);
              }
            }
          };
          if (general) logits(std::true_type{}); else logits(std::false_type{});
          if (kb == 0) {
            m_run = mx;                                // nothing accumulated yet: the first block sets the reference
          } else {
            const bool grow = mx > m_run + kLazy;
            if (__any_sync(0xffffffffu, grow)) {       // rescale this warp's rows of O in place (factor 1 where not needed)
              const float m_new = grow ? mx : m_run;
              const float f = ex2(m_run - m_new);
#pragma unroll
              for (int hf = 0; hf < HD / 16; ++hf) {   // 16 columns at a time: the 48 logits stay live meanwhile
                uint32_t o[16];
                tmem_ld16(t_row + O_COL + hf * 16, o);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * f);
                tmem_st16(t_row + O_COL + hf * 16, o);
              }
              l_run *= f;
              m_run = m_new;
            }
          }
          float l0 = 0.f, l1 = 0.f;
#pragma unroll
          for (int e = 0; e < KB / 2; ++e) {           // packed in place: sv[e] <- (p[2e], p[2e+1])
            const float p0 = ex2(__uint_as_float(sv[2 * e]) - m_run), p1 = ex2(__uint_as_float(sv[2 * e + 1]) - m_run);
            l0 += p0;
            l1 += p1;
            sv[e] = pack_bf16(p0, p1);
          }
          l_run += l0 + l1;
          tmem_st16(t_row + sb, sv);                   // P over the start of this block's own S buffer
          tmem_st8(t_row + sb + 16, sv + 16);
          ptx::tmem_st_wait();
        }
        ptx::cp_async_wait<0>();                       // block kb + 1 has landed
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncthreads();                               // every row's P (and any rescaled O) is in TMEM
        if (tid == 0) {
          ptx::tc_fence_after();
          const uint32_t y0_s = ring_s + (uint32_t)(kb % NSTAGE) * kStage;
          const uint64_t bv = desc_mn + ((y0_s + kYTile) >> 4);
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks)
            ptx::mma_bf16_ts(tmem_base + O_COL, tmem_base + sb + ks * 8, bv + ks * 64, idesc_r, (kb | ks) != 0 ? 1u : 0u);
          if (kb + 1 < a.nkb) {
            const uint64_t by = desc_k + ((ring_s + (uint32_t)((kb + 1) % NSTAGE) * kStage) >> 4);
            ptx::mma_bf16_ss(tmem_base + (KB - sb), ax, by, idesc_s, 0u);
            ptx::mma_bf16_ss(tmem_base + (KB - sb), ax + 2, by + 2, idesc_s, 1u);
          }
          ptx::mma_commit(&bar_mma);
        }
      }
      ptx::mbar_wait(&bar_mma, ph);
      ph ^= 1;
      ptx::tc_fence_after();
      if (warp_live) {
        uint32_t o[HD];
        tmem_ld16(t_row + O_COL, o);
        tmem_ld16(t_row + O_COL + 16, o + 16);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < HD; ++c) acc0[c] = __uint_as_float(o[c]);
      }
    }

#pragma unroll 1
    for (int kb = 0; MODE != MODE_FWD && kb < a.nkb; ++kb) {
      const int stage = kb % NSTAGE;
      ptx::cp_async_wait<NSTAGE - 2>();              // block kb (and the stationary tiles) have landed
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncthreads();                               // (A) also: everybody has read the products of block kb - 1
      if (tid == 0) {
        ptx::tc_fence_after();
        const uint32_t y0_s = ring_s + (uint32_t)stage * kStage;
        const uint64_t ax = desc_k + (xt_s >> 4), by = desc_k + (y0_s >> 4);
        ptx::mma_bf16_ss(tmem_base + S_COL, ax, by, idesc_s, 0u);
        ptx::mma_bf16_ss(tmem_base + S_COL, ax + 2, by + 2, idesc_s, 1u);
        if (MODE != MODE_FWD) {
          const uint64_t ax1 = desc_k + ((xt_s + kXTile) >> 4), by1 = desc_k + ((y0_s + kYTile) >> 4);
          ptx::mma_bf16_ss(tmem_base + DP_COL, ax1, by1, idesc_s, 0u);
          ptx::mma_bf16_ss(tmem_base + DP_COL, ax1 + 2, by1 + 2, idesc_s, 1u);
        }
        ptx::mma_commit(&bar_mma);
      }
      // the stage of block kb + NSTAGE - 1 was last read by the MMAs of block kb - 1, which every thread has waited for
      if (kb + NSTAGE - 1 < a.nkb) gather_block(kb + NSTAGE - 1);
      ptx::cp_async_commit();

      const bool tail_blk = (kb + 1) * KB > N;       // keys / queries beyond the window in this block
      const bool general = need_mask || tail_blk;    // CTA-uniform: most blocks take the copy without mask / tail tests
      const int4* kof4 = reinterpret_cast<const int4*>(kof + kb * KB);
      const uint2* rid8 = reinterpret_cast<const uint2*>(rid + kb * KB);            // eight keys per pair of words
      ptx::mbar_wait(&bar_mma, ph);
      ph ^= 1;
      ptx::tc_fence_after();

      if (!warp_live) {
        // nothing to compute; the block products of these lanes are never read
      } else {
        // ---- P = exp2(s - lse), dS = P (dP - D), eight streamed tokens per step
        const float4* lse4 = reinterpret_cast<const float4*>(blk_lse + stage * KB);
        const float4* d4 = reinterpret_cast<const float4*>(blk_d + stage * KB);
        auto sweep = [&](auto gen_c) {
          constexpr bool GEN = decltype(gen_c)::value;
#pragma unroll 1
          for (int cc = 0; cc < KB / 8; ++cc) {
            uint32_t sv[8], dv[8];
            tmem_ld8(t_row + S_COL + cc * 8, sv);
            tmem_ld8(t_row + DP_COL + cc * 8, dv);
            int kj[8], rj[8];
            {
              const int4 m0 = kof4[cc * 2], m1 = kof4[cc * 2 + 1];
              kj[0] = m0.x; kj[1] = m0.y; kj[2] = m0.z; kj[3] = m0.w; kj[4] = m1.x; kj[5] = m1.y; kj[6] = m1.z; kj[7] = m1.w;
            }
            if (GEN) {
              const uint2 m = rid8[cc];
              rj[0] = m.x & 0xff; rj[1] = (m.x >> 8) & 0xff; rj[2] = (m.x >> 16) & 0xff; rj[3] = m.x >> 24;
              rj[4] = m.y & 0xff; rj[5] = (m.y >> 8) & 0xff; rj[6] = (m.y >> 16) & 0xff; rj[7] = m.y >> 24;
            }
            float lsev[8], dvv[8];
            if (MODE == MODE_DKV) {
              const float4 l0 = lse4[cc * 2], l1 = lse4[cc * 2 + 1], e0 = d4[cc * 2], e1 = d4[cc * 2 + 1];
              lsev[0] = l0.x; lsev[1] = l0.y; lsev[2] = l0.z; lsev[3] = l0.w; lsev[4] = l1.x; lsev[5] = l1.y; lsev[6] = l1.z; lsev[7] = l1.w;
              dvv[0] = e0.x; dvv[1] = e0.y; dvv[2] = e0.z; dvv[3] = e0.w; dvv[4] = e1.x; dvv[5] = e1.y; dvv[6] = e1.z; dvv[7] = e1.w;
            }
            ptx::tmem_ld_wait();
            uint32_t pkd[4], pkp[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
              float pl[2], dl[2];
#pragma unroll
              for (int e1 = 0; e1 < 2; ++e1) {
                const int e = 2 * e2 + e1;
                // DQ: this thread is the query, the streamed token the key; DKV: the other way round
                const int boff = MODE == MODE_DQ ? -kj[e] : kj[e];
                const float cosv = __uint_as_float(sv[e]);
                float s2 = fmaf(cosv, scale2, lds_f32(tabq + (uint32_t)boff));
                if (GEN && need_mask && (rj[e] & 0x7f) != rid_st) s2 += kMaskLog2;
                float p = ex2(s2 - (MODE == MODE_DQ ? lse2_st : lsev[e]));
                if (GEN && MODE == MODE_DQ && (rj[e] >> 7)) p = 0.f;              // key beyond the window
                const float dsv = p * (__uint_as_float(dv[e]) - (MODE == MODE_DQ ? d_st : dvv[e]));
                pl[e1] = p;
                dl[e1] = dsv;
                if (MODE == MODE_DQ) {
                  dsc = fmaf(dsv, cosv, dsc);
                  // gradient of the bias table: warp-private sums.  The 32 lanes of a step hit 32 distinct entries (one
                  // key, 32 different queries); consecutive steps of different lanes alias, hence the warp barrier.
                  const float cur = lds_f32(dtabq + (uint32_t)boff);
                  if (row_valid) sts_f32(dtabq + (uint32_t)boff, cur + dsv);
                  __syncwarp();
                }
              }
              pkd[e2] = pack_bf16(dl[0], dl[1]);
              pkp[e2] = pack_bf16(pl[0], pl[1]);
            }
            tmem_st4(t_row + DP_COL + cc * 4, pkd);
            if (MODE == MODE_DKV) tmem_st4(t_row + S_COL + cc * 4, pkp);
          }
        };
        if (general) sweep(std::true_type{}); else sweep(std::false_type{});
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      __syncthreads();                               // (B) every row's A operand is in TMEM
      if (tid == 0) {
        ptx::tc_fence_after();
        const uint32_t y0_s = ring_s + (uint32_t)stage * kStage;
        if (MODE == MODE_DQ) {
          const uint64_t bk = desc_mn + (y0_s >> 4);
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks)
            ptx::mma_bf16_ts(tmem_base + DP_COL + RES_OFF, tmem_base + DP_COL + ks * 8, bk + ks * 64, idesc_r, ks);
        } else {
          const uint64_t bq = desc_mn + (y0_s >> 4), bg = desc_mn + ((y0_s + kYTile) >> 4);
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks)
            ptx::mma_bf16_ts(tmem_base + S_COL + RES_OFF, tmem_base + S_COL + ks * 8, bg + ks * 64, idesc_r, ks);
#pragma unroll
          for (int ks = 0; ks < KB / 16; ++ks)
            ptx::mma_bf16_ts(tmem_base + DP_COL + RES_OFF, tmem_base + DP_COL + ks * 8, bq + ks * 64, idesc_r, ks);
        }
        ptx::mma_commit(&bar_mma);
      }
      ptx::mbar_wait(&bar_mma, ph);
      ph ^= 1;
      ptx::tc_fence_after();
      if (warp_live) {
        uint32_t o[HD];
        tmem_ld16(t_row + (MODE == MODE_DQ ? DP_COL : S_COL) + RES_OFF, o);
        tmem_ld16(t_row + (MODE == MODE_DQ ? DP_COL : S_COL) + RES_OFF + 16, o + 16);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < HD; ++c)
          acc0[c] += __uint_as_float(o[c]);
        if (MODE == MODE_DKV) {
          tmem_ld16(t_row + DP_COL + RES_OFF, o);
          tmem_ld16(t_row + DP_COL + RES_OFF + 16, o + 16);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < HD; ++c) acc1[c] += __uint_as_float(o[c]);
        }
      }
    }
    ptx::cp_async_wait<0>();

    // ---- epilogue of the unit (the stationary tiles stay valid until the next unit's gather, behind a barrier)
    if (MODE == MODE_FWD) {
      if (row_valid) {
        a.lse[(win * a.nH + h) * N + r_st] = (m_run + log2f(l_run)) * kLn2;
        if (t_st >= 0) {
          const float inv = 1.0f / l_run;
          uint4* dst = reinterpret_cast<uint4*>(a.out + (int64_t)t_st * a.C + h * HD);
          uint4* dlo = a.out_lo ? reinterpret_cast<uint4*>(a.out_lo + (int64_t)t_st * a.C + h * HD) : nullptr;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o8[e] = acc0[c * 8 + e] * inv;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              hi[e] = pack_bf16(o8[2 * e], o8[2 * e + 1]);
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[e]));
              lo[e] = pack_bf16(o8[2 * e] - f.x, o8[2 * e + 1] - f.y);
            }
            dst[c] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (dlo) dlo[c] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
      }
    } else if (MODE == MODE_DQ) {
      if (t_st >= 0)
        normalize_bwd_store(acc0, xt, r_loc, sc, a.inv_norm ? a.inv_norm[((int64_t)t_st * 2 + 0) * a.nH + h] : 0.f,
                            a.dqkv + (int64_t)t_st * C3 + h * HD);
    } else {
      if (t_st >= 0) {
        uint4* dst = reinterpret_cast<uint4*>(a.dqkv + (int64_t)t_st * C3 + 2 * a.C + h * HD);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_uint4(pack_bf16(acc0[c * 8 + 0], acc0[c * 8 + 1]), pack_bf16(acc0[c * 8 + 2], acc0[c * 8 + 3]),
                              pack_bf16(acc0[c * 8 + 4], acc0[c * 8 + 5]), pack_bf16(acc0[c * 8 + 6], acc0[c * 8 + 7]));
        normalize_bwd_store(acc1, xt, r_loc, sc, a.inv_norm ? a.inv_norm[((int64_t)t_st * 2 + 1) * a.nH + h] : 0.f,
                            a.dqkv + (int64_t)t_st * C3 + a.C + h * HD);
      }
      // pad keys carry v = v_bias: their dV rows are gradient of v_bias (one atomic per column and warp)
      if (a.dvpad && (g.Hp != g.H || g.Wp != g.W)) {
        const bool is_pad = t_st == -1;
        if (__any_sync(0xffffffffu, is_pad)) {
#pragma unroll
          for (int c = 0; c < HD; ++c) {
            const float v = warp_sum(is_pad ? acc0[c] : 0.f);
            if (lane == 0 && v != 0.f) atomicAdd(a.dvpad + h * HD + c, v);
          }
        }
      }
    }
    __syncthreads();                                 // the stationary tiles / tables may be overwritten now
  }
  if (MODE == MODE_DQ && cur_h >= 0) flush_head(cur_h, cur_tile);

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

size_t flash_smem(int mode, int KB, int ntab, int ntabc, int ntabq, int nmeta) {
  const size_t kStage = 2 * (size_t)KB * 64;
  size_t s = 1024 + 2 * (size_t)kXTile + (size_t)NSTAGE * kStage + 2 * (size_t)NSTAGE * KB * 4 + 2 * (size_t)nmeta * 4 +
             (size_t)nmeta + (mode == MODE_DQ ? (size_t)ntabq * 4 + 4 * (size_t)ntabc * 4 : (size_t)ntab * 4) + 16;
  // at most four CTAs per SM (128 TMEM columns each): never let a fifth fit by shared memory
  const size_t floor_bytes = 46 * 1024;
  return s < floor_bytes ? floor_bytes : s;
}

template <int MODE, int KB>
int launch_flash_kb(const FlArgs& a, cudaStream_t st) {
  const size_t smem = flash_smem(MODE, KB, a.ntab, a.ntabc, a.ntabq, a.nmeta);
  BSW_REQUIRE(smem <= 227 * 1024, "attn(flash): window %dx%d needs %zu bytes of shared memory", a.g.ws, a.g.ws, smem);
  BSW_CUDA(cudaFuncSetAttribute(attn_flash_kernel<MODE, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // whole unified L1 as shared memory: several CTAs per SM.  (The occupancy query answers for the carve-out of the
  // moment -- 1 CTA per SM before the first launch -- so the residency is computed here: 128 threads x 128 registers and
  // 128 TMEM columns allow four CTAs, shared memory decides the rest.)
  BSW_CUDA(cudaFuncSetAttribute(attn_flash_kernel<MODE, KB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int occ = (int)((227 * 1024) / (smem + 1024));
  if (occ < 1) occ = 1;
  if (occ > 4) occ = 4;
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > a.nunits) grid = a.nunits;
  attn_flash_kernel<MODE, KB><<<(unsigned)grid, kThreads, smem, st>>>(a);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

template <int MODE>
int launch_flash(const FlArgs& a, cudaStream_t st) {
  if constexpr (MODE == MODE_FWD) return launch_flash_kb<MODE, 48>(a, st);
  else return a.kb == 48 ? launch_flash_kb<MODE, 48>(a, st) : launch_flash_kb<MODE, 64>(a, st);
}

void set_block(FlArgs* a, int kb) {
  a->kb = kb;
  a->nkb = (a->N + kb - 1) / kb;
  a->nmeta = a->nkb * kb;
}

int fill_args(FlArgs* a, int B, int H, int W, int C, int nH, int ws, int shift) {
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && nH > 0 && ws >= 1 && ws <= 32, "attn(flash): bad dimension");
  BSW_REQUIRE(C == nH * HD, "attn(flash): head_dim must be 32 (C=%d, nH=%d)", C, nH);
  BSW_REQUIRE(shift >= 0 && shift < ws, "attn(flash): bad shift");
  BSW_REQUIRE(C % 8 == 0, "attn(flash): C must be a multiple of 8");
  BSW_REQUIRE((int64_t)B * H * W < (1ll << 29), "attn(flash): too many tokens");
  make_geom(&a->g, B, H, W, ws, shift);
  a->C = C; a->nH = nH;
  a->N = ws * ws;
  // row tiles of 128 (the last one short: its empty warps skip the element-wise work, so a 144-row window costs five
  // warp passes, not eight) and streamed blocks of 64 tokens -- 48 where that divides the window exactly
  a->ntiles = (a->N + 127) / 128;
  a->rpt = a->N < 128 ? a->N : 128;
  a->ntab = (2 * ws - 1) * (2 * ws - 1);
  {
    // rows of the bias table a warp (32 consecutive query rows of a tile) / a CTA (the whole tile) can touch: ws - 1 + the
    // window rows it spans, maximised over the tiles and warps that exist
    int span_w = 1, span_c = 1;
    for (int t = 0; t < a->ntiles; ++t) {
      const int r0 = t * a->rpt, rows_t = a->N - r0 < a->rpt ? a->N - r0 : a->rpt;
      const int sc_ = (r0 % ws + rows_t - 1) / ws + 1;
      if (sc_ > span_c) span_c = sc_;
      for (int w = 0; w * 32 < rows_t; ++w) {
        const int rw = r0 + w * 32, n_w = rows_t - w * 32 < 32 ? rows_t - w * 32 : 32;
        const int sw = (rw % ws + n_w - 1) / ws + 1;
        if (sw > span_w) span_w = sw;
      }
    }
    const int TW_ = 2 * ws - 1;
    a->ntabc = ws - 1 + span_w < TW_ ? (ws - 1 + span_w) * TW_ : a->ntab;
    a->ntabq = ws - 1 + span_c < TW_ ? (ws - 1 + span_c) * TW_ : a->ntab;
  }
  set_block(a, 64);
  a->nwin = (int64_t)B * a->g.nWh * a->g.nWw;
  a->nunits = a->nwin * nH * a->ntiles;
  BSW_REQUIRE(a->nwin < (1ll << 31), "attn(flash): too many windows");
  return B200SWIN_OK;
}
}  // namespace

// D = <dO, O> per (token, head) (attn_bwd_ws.cu)
int attn_bwd_prep(const void* dout, const void* out, const void* out_lo, float* dvec, int64_t n, cudaStream_t st);

int attn_fwd_flash(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                   const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st) {
  FlArgs a = {};
  int rc = fill_args(&a, B, H, W, C, nH, ws, shift);
  if (rc) return rc;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.out_lo = (__nv_bfloat16*)out_lo; a.lse = lse;
  a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  set_block(&a, 48);                                  // the forward always streams 48-token blocks (TMEM: 48 + 48 + 32)
  return launch_flash<MODE_FWD>(a, st);
}

size_t attn_bwd_flash_workspace_bytes(int B, int H, int W, int nH) { return (size_t)B * H * W * nH * sizeof(float); }

int attn_bwd_flash(const void* qkv, const void* out, const void* out_lo, const void* dout, const float* lse, const float* inv_norm,
                   const float* table16, const float* scale, const float* qpad, const float* vpad, void* dqkv,
                   float* dtable16, float* dscale, float* dvpad, void* workspace, int B, int H, int W, int C, int nH,
                   int ws, int shift, cudaStream_t st) {
  BSW_REQUIRE(workspace, "attn_bwd(flash): workspace for D = <dO, O> missing");
  FlArgs a = {};
  int rc = fill_args(&a, B, H, W, C, nH, ws, shift);
  if (rc) return rc;
  a.qkv = (const __nv_bfloat16*)qkv; a.dout = (const __nv_bfloat16*)dout; a.lse = const_cast<float*>(lse);
  a.dvec = (const float*)workspace; a.inv_norm = inv_norm; a.table16 = table16; a.scale = scale; a.qpad = qpad;
  a.vpad = vpad; a.dqkv = (__nv_bfloat16*)dqkv; a.dtable16 = dtable16; a.dscale = dscale; a.dvpad = dvpad;
  rc = attn_bwd_prep(dout, out, out_lo, (float*)workspace, (int64_t)B * H * W * nH, st);
  if (rc) return rc;
  // backward passes: blocks of 64 streamed tokens, or of 48 where that pads the window less (12x12: 144 = 3 x 48 exactly;
  // 30x30: 900 -> 912 instead of 960)
  set_block(&a, (a.N + 47) / 48 * 48 < (a.N + 63) / 64 * 64 ? 48 : 64);
  rc = launch_flash<MODE_DQ>(a, st);
  if (rc) return rc;
  return launch_flash<MODE_DKV>(a, st);
}

}  // namespace b200swin
