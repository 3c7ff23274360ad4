// Warp-specialised attention-core forward on tcgen05 (bf16 storage, windows up to 12x12 = 144 tokens).
//
// Same math as attn_tc.cu (models/swin_transformer_v2.py:295-328 with the pad / roll / partition / reverse /
// crop of :429-463 and the shift mask of :874-892 as address math), restructured as a dataflow pipeline so that
// the gather, the tensor pipe and the softmax never wait for each other:
//
//   warps 1-3   gather: cp.async of the window's q_hat / k_hat / v rows from the natural [B,H,W,3C] tensor into a
//               4-stage ring of 64-byte-swizzled UMMA tiles (pad tokens written in place);
//   warp 0      one thread issues tcgen05.mma:  S = Q K^T into one of up to four TMEM slots,  O = P V (P read
//               from TMEM) -- scheduled by polling mbarriers, so S of the next items is in flight while the
//               softmax warps work;
//   warps 4-19  softmax: two unit groups on alternating items; inside a group every query row is owned by TWO
//               threads (warp pair sharing a TMEM lane quarter), one per half of the keys.  Each half keeps its
//               part of the fp32 S row in registers, adds the bias from a per-head [N][N] fp32 matrix in shared
//               memory (conflict-free float4 reads), applies the shift mask from per-thread bit masks (only
//               windows on the last window row / column pay for it), takes its OWN maximum and sum, and writes
//               P = exp2(s - m_half) back to TMEM as packed bf16 over the S columns.  The two halves feed two
//               accumulators O_a, O_b (split-K flash-attention style); the second-half thread combines them as
//               (O_a 2^(m_a-m) + O_b 2^(m_b-m)) / (l_a 2^(m_a-m) + l_b 2^(m_b-m)) and stores the natural
//               [B,H,W,C] layout.  No thread ever waits for the other half's maximum.
//
// A 12x12 window has 144 = 128 + 16 rows.  The 16-row tail is a second M=128 MMA whose A operand starts
// 32*rot rows early, so the tail lands in TMEM lane quarter `rot` -- rot rotates per item and the extra softmax
// pass is spread over the four SM sub-partitions instead of always hitting the first.
//
// Items are ordered head-major and every CTA owns one contiguous range, so the expanded bias matrix is rebuilt
// only when the head changes (at most a few times per CTA).
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "wingeom.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

namespace {
constexpr int HD = 32;
// Warps 0-15: softmax; warps 16-17: gather; warp 18: issuer of S = Q K^T; warp 19: issuer of O = P V.  The warp
// scheduler favours the highest warp id of a sub-partition, so the latency-critical (but nearly idle) issuer and gather
// warps sit ABOVE the softmax warps: as warp 0 the issuer was starved of issue slots and every MMA hand-off took ~500
// cycles.  TWO issuer threads: with one, S of a later unit queued behind the blocking wait for P of an earlier one
// (a convoy that left the softmax warps waiting for S two thirds of the time, tools trace of round 1).
constexpr int kThreads = 640;
constexpr int kIssuerWarp = 19;      // P V issuer (also allocates TMEM)
constexpr int kSIssuerWarp = 18;     // Q K^T issuer
constexpr int kSoftmax = 512;
constexpr int kLoaders = 64;
// register budget: 640 threads x 96 at launch = 61440 = 128 x kRegService + 512 x kRegSoftmax
constexpr int kRegService = 64, kRegSoftmax = 104;
constexpr int NSTAGE = 4, LAG = 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMaskLog2 = -100.0f * 1.4426950408889634f;
constexpr uint32_t kSw64 = 4;       // UMMA layout type SWIZZLE_64B

struct WsArgs {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  __nv_bfloat16* out_lo;      // optional: bf16 residual O - bf16(O), so that the backward's D = <dO, O> sees O to ~2^-17
  float* lse;
  const float* table16;
  const float* scale;
  const float* qpad;
  const float* vpad;
  WinGeom g;
  int C, nH;
  int64_t nwin;       // B * nWh * nWw
  int64_t nitems;     // nH * nwin, item = head * nwin + win
  long long* trace;   // debug builds (-DB200SWIN_TRACE): [0] = event count, then (event, warp, index, clock) records
};

#ifdef B200SWIN_TRACE
// per-warp private event log (no atomics: a returning global atomic costs ~500 cycles and hides what it measures)
#define TR(ev, idx)                                                                           \
  do {                                                                                        \
    if (a.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && trc_ < 3000) {               \
      long long* e_ = a.trace + 1 + 4 * ((threadIdx.x >> 5) * 3000 + trc_);                   \
      e_[0] = (ev); e_[1] = threadIdx.x >> 5; e_[2] = (idx); e_[3] = clock64();               \
      ++trc_;                                                                                 \
    }                                                                                         \
  } while (0)
#else
#define TR(ev, idx) do {} while (0)
#endif

template <int WS>
struct Cfg {
  static constexpr int N = WS * WS;
  static constexpr int NPAD = (N + 15) / 16 * 16;
  static constexpr int MT = (NPAD + 127) / 128;
  static constexpr int TAIL = N - 128 * (MT - 1);                  // valid rows of the last tile
  static constexpr bool ROT = MT > 1 && TAIL <= 32;
  static constexpr int KA = (NPAD + 31) / 32 * 16, KB = NPAD - KA;  // keys of the first / second half
  // Slot layout.  Each half writes its packed P over the START OF ITS OWN S columns (P_a at 0, P_b at KA): a half
  // must never write into columns whose S the other half may not have read yet (the two warps of a row are not
  // synchronised -- the first layout put P_b behind P_a, inside S_a, and a delayed first-half warp read garbage).
  // The accumulators are written by the MMA only after all eight warps have read S: O_a / O_b sit behind P_a / P_b
  // inside the S columns when those are wide enough (12x12 windows), else behind S.
  static constexpr bool INPLACE = (KA >= 80) && (KB >= 64);
  static constexpr int OA = INPLACE ? (KA / 2 + 15) / 16 * 16 : NPAD, OB = INPLACE ? KA + KB / 2 : NPAD + HD;
  static_assert(!INPLACE || (OA + HD <= KA && OB + HD <= NPAD), "accumulators must fit behind the packed P");
  static constexpr int SLOTW = INPLACE ? NPAD : NPAD + 2 * HD;        // TMEM columns per slot
  static constexpr int NSLOT = 512 / SLOTW > 4 ? 4 : 512 / SLOTW;
  // Slot of a unit and how often that slot has been used before.  One tile per item: a plain ring.  Two tiles (12x12):
  // slots by ROLE -- the main tile of unit group g lives in slot g, every tail in slot 2 -- so that S of a group's
  // next item is issued as soon as the epilogue of its current main tile has drained the slot, i.e. while the group
  // is still busy with the tail (with a ring, the next main tile had to wait for the tail's slot).
  static constexpr bool ROLES = MT == 2 && NSLOT >= 3;
  __device__ static __forceinline__ int slot_of(int u) {
    if constexpr (ROLES) return (u & 1) ? 2 : ((u >> 1) & 1);
    else return u % NSLOT;
  }
  __device__ static __forceinline__ uint32_t uses_of(int u) {
    if constexpr (ROLES) return (u & 1) ? (uint32_t)(u >> 1) : (uint32_t)(u >> 2);
    else return (uint32_t)(u / NSLOT);
  }
  static constexpr int NS = N + ((12 - N % 8) % 8);                 // bias row stride, NS % 8 == 4: float4 reads
  static_assert(NS % 8 == 4 && NS >= N, "bias stride");             //   of 8 consecutive rows hit 8 bank groups
  static constexpr uint32_t kRow = NPAD * 64;                       // one [NPAD][64 B] operand tile
  static constexpr uint32_t kStage = 3 * kRow;                      // Q | K | V
  static constexpr int TW = 2 * WS - 1, NTAB = TW * TW;
  static constexpr size_t kSmem =
      1024 + 16 + (size_t)NSTAGE * kStage + (size_t)N * NS * 4 + (size_t)NTAB * 4 + (size_t)NSLOT * 128 * 8;
};

__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// bf16 of (fp32 values - the bf16 values already packed in `hi`): the low half of a hi/lo split
__device__ __forceinline__ uint4 residual_bf16(const uint4& hi, const float* v) {
  const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w};
  uint32_t r[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h[e]));
    r[e] = pack_bf16(v[2 * e] - f.x, v[2 * e + 1] - f.y);
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

// source token of in-window position r of window (b, wh, ww); -1 for a pad token
__device__ __forceinline__ int src_token(const WinGeom& g, int b, int wh, int ww, int y, int x) {
  int i = wh * g.ws + y + g.shift; if (i >= g.Hp) i -= g.Hp;
  int j = ww * g.ws + x + g.shift; if (j >= g.Wp) j -= g.Wp;
  return (i < g.H && j < g.W) ? (b * g.H + i) * g.W + j : -1;
}

// One thread = one query row x the keys [C0, C0 + NC) of the window.  Reads its part of S from TMEM, turns it into
// logits in log2 units (scale * cos + bias, shift mask), takes the maximum m and sum l over ITS keys and writes
// P = exp2(s - m) as packed bf16 over the first NC/2 of ITS OWN S columns [C0, C0 + NC/2).  All lanes must call it
// (tcgen05.ld / st are warp-collective); `valid` lanes own a real row.
template <int WS, int C0, int NC>
__device__ __forceinline__ void softmax_half(uint32_t t_s, const float* brow, float scale2, bool need_mask, uint32_t by,
                                             uint32_t bx, bool valid, float& m_out, float& l_out
#ifdef B200SWIN_TRACE
                                             , const WsArgs& a, int& trc_, int u
#endif
                                             ) {
  constexpr int N = WS * WS;
  static_assert(C0 % 16 == 0 && NC % 16 == 0, "key halves are whole k-steps");
  if constexpr (NC > 0) {
    uint32_t sv[NC];
#pragma unroll
    for (int c = 0; c < NC / 16; ++c) tmem_ld16(t_s + C0 + c * 16, &sv[c * 16]);
    ptx::tmem_ld_wait();
    TR(40, u);
    if (valid) {
      const float4* b4 = reinterpret_cast<const float4*>(brow + C0);
#pragma unroll
      for (int j4 = 0; j4 < NC / 4; ++j4) {
        if (C0 + j4 * 4 < N) {
          const float4 bb = b4[j4];
          const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (C0 + j4 * 4 + k < N) sv[j4 * 4 + k] = __float_as_uint(fmaf(__uint_as_float(sv[j4 * 4 + k]), scale2, bv[k]));
        }
      }
      if (need_mask) {
#pragma unroll
        for (int jj = 0; jj < NC; ++jj) {
          const int j = C0 + jj, yj = j / WS, xj = j % WS;
          if (j < N && (((by >> yj) | (bx >> xj)) & 1u)) sv[jj] = __float_as_uint(__uint_as_float(sv[jj]) + kMaskLog2);
        }
      }
      float m = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < NC; ++jj)
        if (C0 + jj < N) m = fmaxf(m, __uint_as_float(sv[jj]));
      TR(41, u);
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int jj = 0; jj < NC; jj += 2) {
        const float p0 = C0 + jj < N ? ex2(__uint_as_float(sv[jj]) - m) : 0.f;
        const float p1 = C0 + jj + 1 < N ? ex2(__uint_as_float(sv[jj + 1]) - m) : 0.f;
        l0 += p0;
        l1 += p1;
        sv[jj >> 1] = pack_bf16(p0, p1);
      }
      m_out = m;
      l_out = l0 + l1;
    }
    TR(42, u);
#pragma unroll
    for (int c = 0; c < NC / 16; ++c) tmem_st8(t_s + C0 + c * 8, &sv[c * 8]);    // over the start of this half's own S
    ptx::tmem_st_wait();
    TR(43, u);
  }
}

template <int WS>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_ws_kernel(const __grid_constant__ WsArgs a) {
  using CF = Cfg<WS>;
  constexpr int N = CF::N, NPAD = CF::NPAD, MT = CF::MT, NSLOT = CF::NSLOT, NS = CF::NS;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t kv_full[NSTAGE], kv_empty[NSTAGE];
  __shared__ __align__(8) uint64_t s_full[NSLOT], p_full[NSLOT], o_full[NSLOT], slot_free[NSLOT];
  __shared__ uint32_t tmem_slot;
  // Item headers (head, batch, window row / column, window index), written by the gather warps, which derive them anyway.
  // The softmax warps re-read them from shared memory wherever they need them: no divisions per item (three, one of
  // them 64-bit, cost each warp ~1000 cycles per item) and nothing about the item lives in registers across the
  // softmax (the allocator spilled it: seven L2 round trips per epilogue).  Twice as deep as the operand ring: the
  // epilogue of an item may still read its header after the item's stage has been handed back.
  __shared__ int s_hdr[2 * NSTAGE][8];

  const WinGeom& g = a.g;
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  float* bias = reinterpret_cast<float*>(sm + (size_t)NSTAGE * CF::kStage);
  float* tab = bias + N * NS;
  float2* ml = reinterpret_cast<float2*>(tab + CF::NTAB + (CF::NTAB & 1));   // [NSLOT][128] (m, l) of the first half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef B200SWIN_TRACE
  int trc_ = 0;
#endif

  // contiguous, balanced item range of this CTA (items are head-major)
  const int64_t per = a.nitems / gridDim.x, rem = a.nitems % gridDim.x;
  const int64_t g0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int n = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));
  const int nW = g.nWh * g.nWw;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&kv_full[s], kLoaders); ptx::mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_full[s], 8);
      ptx::mbar_init(&o_full[s], 1);
      ptx::mbar_init(&slot_free[s], 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kIssuerWarp) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp >= 16) {
    reg_dec<kRegService>();
    if (warp == kSIssuerWarp) {
      // =================================================================================== issuer of S = Q K^T
      if (lane == 0) {
        // runs ahead of the softmax as far as free slots (and gathered items) allow; blocking hardware-sleep waits
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, NPAD, 0, 0);
        const uint64_t desc_k = ptx::make_smem_desc(0, 16, 512, kSw64);      // K-major Q / K tiles (64 B rows)
        const int U = n * MT;
        for (int u = 0; u < U; ++u) {
          const int slot = CF::slot_of(u);
          const int il = u / MT, tile = u - il * MT, stage = il % NSTAGE;
          TR(10, u);
          ptx::mbar_wait(&slot_free[slot], (CF::uses_of(u) & 1) ^ 1);
          TR(11, u);
          if (tile == 0) ptx::mbar_wait(&kv_full[stage], (il / NSTAGE) & 1);
          TR(12, u);
          ptx::tc_fence_after();
          const uint32_t q_s = base_u32 + (uint32_t)stage * CF::kStage, k_s = q_s + CF::kRow;
          int row0 = tile * 128;
          if (CF::ROT && tile == MT - 1) row0 -= 32 * ((il >> 1) & 3);
          const uint32_t t_s = tmem_base + (uint32_t)slot * CF::SLOTW;
          const uint64_t ad = desc_k + ((q_s + row0 * 64) >> 4), bd = desc_k + (k_s >> 4);
          ptx::mma_bf16_ss(t_s, ad, bd, idesc_qk, 0u);
          ptx::mma_bf16_ss(t_s, ad + 2, bd + 2, idesc_qk, 1u);               // second k-step: +32 B
          ptx::mma_commit(&s_full[slot]);
          TR(13, u);
        }
      }
    } else if (warp == kIssuerWarp) {
      // =================================================================================== issuer of O = P V
      if (lane == 0) {
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, HD, 0, 1);   // A = P (TMEM), B = V MN-major
        const uint64_t desc_v = ptx::make_smem_desc(0, 512, 512, kSw64);     // MN-major V tile
        const int U = n * MT;
        for (int u = 0; u < U; ++u) {
          const int slot = CF::slot_of(u);
          const int il = u / MT, tile = u - il * MT, stage = il % NSTAGE;
          TR(14, u);
          ptx::mbar_wait(&p_full[slot], CF::uses_of(u) & 1);
          TR(15, u);
          ptx::tc_fence_after();
          const uint32_t v_s = base_u32 + (uint32_t)stage * CF::kStage + 2 * CF::kRow;
          const uint32_t t_s = tmem_base + (uint32_t)slot * CF::SLOTW;
          const uint64_t bd = desc_v + (v_s >> 4);
#pragma unroll
          for (int ks = 0; ks < NPAD / 16; ++ks) {
            const bool second = ks >= CF::KA / 16;                           // keys of the second half -> O_b
            ptx::mma_bf16_ts(t_s + (second ? CF::OB : CF::OA),
                             t_s + (second ? CF::KA + (ks - CF::KA / 16) * 8 : ks * 8), bd + ks * 64, idesc_pv,
                             (ks != 0 && ks != CF::KA / 16) ? 1u : 0u);
          }
          ptx::mma_commit(&o_full[slot]);
          // every MMA reading this stage has retired: the S MMAs of the item (other thread) completed before the
          // softmax that produced this P could start
          if (tile == MT - 1) ptx::mma_commit(&kv_empty[stage]);
          TR(16, u);
        }
      }
    } else {
      // =================================================================================== gather warps
      const int lt = threadIdx.x - 512;
      const int C3 = 3 * a.C;
      // `pending` gathers are in flight (items i-pending .. i-1, one cp.async group each).  An item is published
      // (kv_full) as soon as it is LAG groups old -- and everything in flight is published before the warp goes to
      // sleep on a full ring, so the MMA warp can always run ahead on what has already landed.
      int pending = 0;
      // item geometry by running counters (no runtime divisions per item)
      const int nwin_i = (int)a.nwin;
      int c_h = (int)(g0 / a.nwin);
      int c_win = (int)(g0 - (int64_t)c_h * a.nwin);
      int c_b = c_win / nW;
      int c_wh = (c_win - c_b * nW) / g.nWw;
      int c_ww = (c_win - c_b * nW) - c_wh * g.nWw;
      for (int i = 0; i < n; ++i) {
        {
          const int stage = i % NSTAGE;
          const uint32_t par = ((i / NSTAGE) & 1) ^ 1;
          if (!ptx::mbar_test_wait(&kv_empty[stage], par)) {
            if (pending) {
              ptx::cp_async_wait<0>();
              ptx::fence_proxy_async_smem();
              for (int k = i - pending; k < i; ++k) ptx::mbar_arrive(&kv_full[k % NSTAGE]);
              pending = 0;
            }
            TR(20, i);
            ptx::mbar_wait(&kv_empty[stage], par);
          }
          TR(21, i);
          const int h = c_h, b = c_b, wh = c_wh, ww = c_ww;
          const int64_t win = c_win;
          if (++c_win == nwin_i) { c_win = 0; ++c_h; c_b = c_wh = c_ww = 0; }
          else if (++c_ww == g.nWw) { c_ww = 0; if (++c_wh == g.nWh) { c_wh = 0; ++c_b; } }
          unsigned char* st = sm + (size_t)stage * CF::kStage;
          if (lt == 0) {
            int* hd = s_hdr[i % (2 * NSTAGE)];
            hd[0] = h; hd[1] = b; hd[2] = wh; hd[3] = ww; hd[4] = (int)win;
          }
          const uint32_t q_s = ptx::smem_u32(st), k_s = q_s + CF::kRow, v_s = k_s + CF::kRow;
          for (int idx = lt; idx < NPAD * 4; idx += kLoaders) {
            const int r = idx >> 2, c = idx & 3;
            const int t = r < N ? src_token(g, b, wh, ww, r / WS, r % WS) : -2;
            const uint32_t off = sw64_off(r, c);
            if (t >= 0) {
              const __nv_bfloat16* src = a.qkv + (int64_t)t * C3 + h * HD + c * 8;
              ptx::cp_async_16(q_s + off, src);
              ptx::cp_async_16(k_s + off, src + a.C);
              ptx::cp_async_16(v_s + off, src + 2 * a.C);
            } else {
              // pad token: q = normalised q_bias, k = 0, v = v_bias; key padding rows (r >= N): all zero
              uint4 qv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
              if (t == -1) {
                if (a.qpad) {
                  const float* p = a.qpad + h * HD + c * 8;
                  qv = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
                }
                if (a.vpad) {
                  const float* p = a.vpad + h * HD + c * 8;
                  vv = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
                }
              }
              *reinterpret_cast<uint4*>(st + off) = qv;
              *reinterpret_cast<uint4*>(st + CF::kRow + off) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(st + 2 * CF::kRow + off) = vv;
            }
          }
        }
        ptx::cp_async_commit();
        TR(22, i);
        if (++pending > LAG) {
          ptx::cp_async_wait<LAG>();              // the gather of item i - LAG has landed
          ptx::fence_proxy_async_smem();          // generic-proxy writes -> visible to tcgen05.mma
          ptx::mbar_arrive(&kv_full[(i - LAG) % NSTAGE]);
          TR(23, i - LAG);
          --pending;
        }
      }
      if (pending) {
        ptx::cp_async_wait<0>();
        ptx::fence_proxy_async_smem();
        for (int k = n - pending; k < n; ++k) ptx::mbar_arrive(&kv_full[k % NSTAGE]);
      }
    }
  } else {
    // ===================================================================================== softmax warps
    reg_inc<kRegSoftmax>();
    const int sw = warp;
    const int ug = sw >> 3;                         // unit group: items of even / odd local index
    const int half = (sw >> 2) & 1;                 // which half of the keys this warp owns
    const int q = warp & 3;                         // TMEM lane quarter of this warp
    const int st = threadIdx.x;                     // 0..511 over all softmax warps
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int cur_head = -1;
    float scale2 = 0.f;

    auto hdr_ld = [&](int il, int f) {               // volatile: always from shared memory, never kept in a register
      int v;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(ptx::smem_u32(&s_hdr[il % (2 * NSTAGE)][f])));
      return v;
    };
    int c_h = (int)(g0 / a.nwin);                    // head of the current item, by a running counter
    int c_left = (int)((int64_t)(c_h + 1) * a.nwin - g0);   // items left in this head
    for (int il = 0; il < n; ++il) {
      const int h = c_h;
      if (--c_left == 0) { ++c_h; c_left = (int)a.nwin; }
      if (h != cur_head) {
        // every softmax warp has finished all earlier items: rebuild the expanded bias matrix of head h (log2 units)
        named_bar_sync(1, kSoftmax);
        for (int t = st; t < CF::NTAB; t += kSoftmax) tab[t] = a.table16[t * a.nH + h] * kLog2e;
        named_bar_sync(1, kSoftmax);
        for (int e = st; e < N * N; e += kSoftmax) {
          const int i = e / N, j = e - i * N;
          bias[i * NS + j] = tab[(i / WS - j / WS + WS - 1) * CF::TW + (i % WS - j % WS + WS - 1)];
        }
        named_bar_sync(1, kSoftmax);
        cur_head = h;
        scale2 = a.scale[h] * kLog2e;
      }
      if ((il & 1) != ug) continue;

#pragma unroll 1
      for (int tile = 0; tile < MT; ++tile) {
        const int u = il * MT + tile;
        const int slot = CF::slot_of(u);
        const uint32_t par = CF::uses_of(u) & 1;
        const bool rot_tile = CF::ROT && tile == MT - 1;
        int r;                                                            // in-window query row of this thread, or -1
        if (rot_tile) r = (q == ((il >> 1) & 3) && lane < CF::TAIL) ? tile * 128 + lane : -1;
        else { r = tile * 128 + q * 32 + lane; if (r >= N) r = -1; }
        const bool warp_active = __any_sync(0xffffffffu, r >= 0);
        const uint32_t t_s = t_lane + (uint32_t)slot * CF::SLOTW;
        float2* ml_row = ml + slot * 128 + q * 32 + lane;

        TR(30, u);
        ptx::mbar_wait(&s_full[slot], par);
        TR(31, u);
        ptx::tc_fence_after();
        // S of the item exists, so its gather -- and its header -- have landed
        const int wh = hdr_ld(il, 2), ww = hdr_ld(il, 3);
        const bool last_h = wh == g.nWh - 1, last_w = ww == g.nWw - 1;
        const bool need_mask = g.shift > 0 && (last_h || last_w);          // CTA-uniform per item
        float m = -INFINITY, l = 0.f;
        if (warp_active && (half == 0 || CF::KB > 0)) {
          uint32_t by = 0, bx = 0;
          if (need_mask && r >= 0) {
            // regions differ only across the roll seam of the last window row / column
            const int yi = r / WS, xi = r - yi * WS;
            const int cut = WS - g.shift;                                   // in-window coordinate of the seam
            const uint32_t hi = (~0u << cut) & ((1u << WS) - 1u), lo = (1u << cut) - 1u;
            by = last_h ? (yi >= cut ? lo : hi) : 0u;                       // bit y set: key row y is masked
            bx = last_w ? (xi >= cut ? lo : hi) : 0u;
          }
          const float* brow = bias + (r >= 0 ? r : 0) * NS;
#ifdef B200SWIN_TRACE
          if (half == 0) softmax_half<WS, 0, CF::KA>(t_s, brow, scale2, need_mask, by, bx, r >= 0, m, l, a, trc_, u);
          else softmax_half<WS, CF::KA, CF::KB>(t_s, brow, scale2, need_mask, by, bx, r >= 0, m, l, a, trc_, u);
#else
          if (half == 0) softmax_half<WS, 0, CF::KA>(t_s, brow, scale2, need_mask, by, bx, r >= 0, m, l);
          else softmax_half<WS, CF::KA, CF::KB>(t_s, brow, scale2, need_mask, by, bx, r >= 0, m, l);
#endif
        }
        if (half == 0) *ml_row = make_float2(m, l);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[slot]);
        TR(32, u);
        if (half == 0 || !warp_active) {
          // nothing more to do for this unit: the first half never reads O, idle quarters have no rows
          if (lane == 0) ptx::mbar_arrive(&slot_free[slot]);
          continue;
        }

        // ---- second half: O_a, O_b are on their way; combine, normalise and store
        ptx::mbar_wait(&o_full[slot], par);
        TR(33, u);
        ptx::tc_fence_after();
        uint32_t oa[32], ob[32];
        ptx::tmem_ld_32x32b_x32(t_s + CF::OA, oa);
        if (CF::KB > 0) ptx::tmem_ld_32x32b_x32(t_s + CF::OB, ob);
        const float2 mla = *ml_row;
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&slot_free[slot]);
        TR(34, u);
        if (r >= 0) {
          const float mm = fmaxf(mla.x, m);
          const float wa = ex2(mla.x - mm), wb = CF::KB > 0 ? ex2(m - mm) : 0.f;
          const float lt = mla.y * wa + l * wb;
          const int hh = hdr_ld(il, 0);
          a.lse[((int64_t)hdr_ld(il, 4) * a.nH + hh) * N + r] = (mm + log2f(lt)) * kLn2;
          const int t = src_token(g, hdr_ld(il, 1), hdr_ld(il, 2), hdr_ld(il, 3), r / WS, r % WS);
          if (t >= 0) {
            const float ia = wa / lt, ib = wb / lt;
            float o[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              o[c] = __uint_as_float(oa[c]) * ia;
              if (CF::KB > 0) o[c] = fmaf(__uint_as_float(ob[c]), ib, o[c]);
            }
            uint4* dst = reinterpret_cast<uint4*>(a.out + (int64_t)t * a.C + hh * HD);
            uint4* dlo = a.out_lo ? reinterpret_cast<uint4*>(a.out_lo + (int64_t)t * a.C + hh * HD) : nullptr;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint4 hi = make_uint4(pack_bf16(o[c * 8 + 0], o[c * 8 + 1]), pack_bf16(o[c * 8 + 2], o[c * 8 + 3]),
                                          pack_bf16(o[c * 8 + 4], o[c * 8 + 5]), pack_bf16(o[c * 8 + 6], o[c * 8 + 7]));
              dst[c] = hi;
              if (dlo) dlo[c] = residual_bf16(hi, &o[c * 8]);
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int WS>
int launch_ws(const WsArgs& a, cudaStream_t st) {
  using CF = Cfg<WS>;
  static_assert(CF::kSmem <= 227 * 1024, "shared memory budget");
  BSW_CUDA(cudaFuncSetAttribute(attn_fwd_ws_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::kSmem));
  int64_t grid = sm_count();
  if (grid > a.nitems) grid = a.nitems;
#ifdef B200SWIN_TRACE
  WsArgs at = a;
  const char* tpath = getenv("B200SWIN_ATTN_TRACE");
  const size_t tbytes = (1 + 4 * 60000) * sizeof(long long);
  if (tpath) {
    BSW_CUDA(cudaMalloc(&at.trace, tbytes));
    BSW_CUDA(cudaMemsetAsync(at.trace, 0, tbytes, st));
  }
  attn_fwd_ws_kernel<WS><<<(unsigned)grid, kThreads, CF::kSmem, st>>>(at);
  BSW_LAUNCH_CHECK();
  if (tpath) {
    std::vector<long long> hbuf(1 + 4 * 60000);
    BSW_CUDA(cudaStreamSynchronize(st));
    BSW_CUDA(cudaMemcpy(hbuf.data(), at.trace, tbytes, cudaMemcpyDeviceToHost));
    BSW_CUDA(cudaFree(at.trace));
    FILE* f = fopen(tpath, "w");
    if (f) {
      for (long long k = 0; k < 60000; ++k)
        if (hbuf[4 + 4 * k] != 0) fprintf(f, "%lld %lld %lld %lld\n", hbuf[1 + 4 * k], hbuf[2 + 4 * k], hbuf[3 + 4 * k], hbuf[4 + 4 * k]);
      fclose(f);
    }
  }
  return B200SWIN_OK;
#else
  attn_fwd_ws_kernel<WS><<<(unsigned)grid, kThreads, CF::kSmem, st>>>(a);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
#endif
}
}  // namespace

bool attn_fwd_ws_supported(int ws) { return ws == 4 || ws == 6 || ws == 7 || ws == 8 || ws == 12; }

int attn_fwd_ws(const void* qkv, void* out, void* out_lo, float* lse, const float* table16, const float* scale, const float* qpad,
                const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st) {
  WsArgs a;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.out_lo = (__nv_bfloat16*)out_lo; a.lse = lse;
  a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  make_geom(&a.g, B, H, W, ws, shift);
  a.C = C; a.nH = nH;
  a.nwin = (int64_t)B * a.g.nWh * a.g.nWw;
  a.nitems = a.nwin * nH;
  a.trace = nullptr;
  BSW_REQUIRE(a.nwin < (1ll << 31), "attn_fwd(ws): too many windows");
  switch (ws) {
    case 4: return launch_ws<4>(a, st);
    case 6: return launch_ws<6>(a, st);
    case 7: return launch_ws<7>(a, st);
    case 8: return launch_ws<8>(a, st);
    case 12: return launch_ws<12>(a, st);
    default: break;
  }
  set_error("attn_fwd(ws): window %dx%d not instantiated", ws, ws);
  return B200SWIN_EINVAL;
}

}  // namespace b200swin
