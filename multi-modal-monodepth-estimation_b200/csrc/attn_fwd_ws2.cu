// Warp-specialised attention-core forward on tcgen05, SECOND GENERATION, 12x12 windows (144 = 128 + 16 query rows).
//
// Same math and the same gather / operand layout as attn_fwd_ws.cu (models/swin_transformer_v2.py:295-328 with the
// pad / roll / partition / reverse / crop of :429-463 and the shift mask of :874-892 as address math).  What changed is
// the softmax organisation, after pipeline traces of the first generation showed its softmax warps waiting for S two
// thirds of the time: with two unit groups each group needs two TMEM slots to prefetch, and only three slots of 144
// columns exist (the shared tail slot was held ~7000 cycles per item and bounded the kernel at one item per 7300).
//
//   * ONE consumer group: all 12 softmax warps work on the same unit, three warps per TMEM lane quarter, each warp one
//     third of the keys (48 each -- whole MMA k-steps; 12 rather than 16 warps so that each thread may keep 144
//     registers: at 104 the unit loop spilled, and a spill is an L2 round trip here).  Units are consumed strictly in order, so
//     the three slots form a plain ring and S of units u + 1, u + 2 is always in flight.
//   * The three warps of a row agree on the row maximum through shared memory and ONE 96-thread named barrier per
//     unit and quarter; all P share that maximum, so there is a single accumulator O and no split-K combine.
//   * O lives in two 32-column buffers OUTSIDE the slots: a slot is released by the commit of its own PV MMAs, not by
//     the epilogue.  The epilogue of unit u is deferred behind the softmax of unit u + 1 and is spread over all four
//     first two warps of the quarter (16 columns of O per thread).
//   * The 16-row tail is a second M = 128 MMA whose A operand starts 32 * rot rows early, so the tail lands in lane
//     quarter rot (rotating per item); the other quarters pass straight through to the next unit.
//
// STATUS (end of round 1): parity-green (tests/test_attention_gpu.py with B200SWIN_ATTN_FWD_GEN2=1) but not the default:
// 201 us vs 180 us of the first generation on Swin-B stage 2.  The pipeline trace shows what is left: every warp now
// pays the per-unit fixed costs (deferred epilogue ~1000 cycles of dependent latency, barrier hand-offs) that the
// two-group design spread over two groups; the softmax proper is 2000-2900 cycles per main unit for 48 keys per thread.
#include <stdlib.h>
#include <vector>
#include "common.cuh"
#include "wingeom.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

namespace {
constexpr int HD = 32;
// Warps 0-11: softmax; 12-13: gather; 14: issuer of S = Q K^T; 15: issuer of O = P V (service warps have the highest
// warp ids: the scheduler favours them and they are nearly always asleep).
constexpr int kThreads = 512;
constexpr int kSoftmaxWarps = 12, kKT = 3;                // key thirds = warps per lane quarter
constexpr int kIssuerWarp = 15;
constexpr int kSIssuerWarp = 14;
constexpr int kSoftmax = kSoftmaxWarps * 32;
constexpr int kLoaders = 64;
// register budget: 512 threads x 128 at launch; 128 x 64 + 384 x 144 = 63488 <= 65536
constexpr int kRegService = 64, kRegSoftmax = 144;
constexpr int NSTAGE = 4, LAG = 2;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMaskLog2 = -100.0f * 1.4426950408889634f;
constexpr uint32_t kSw64 = 4;       // UMMA layout type SWIZZLE_64B

struct WsArgs {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
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
  static_assert(NPAD == 144 && MT == 2, "second-generation forward: 12x12 windows");
  // key thirds (one per warp of a lane quarter): multiples of 16 so that no MMA k-step straddles two of them
  static constexpr int K0 = 0, K1 = 48, K2 = 96, K3 = NPAD;
  static constexpr int SLOTW = NPAD, NSLOT = 3;
  static constexpr uint32_t O_COL = NSLOT * SLOTW;                  // two accumulators of HD columns behind the slots
  static_assert(O_COL + 2 * HD <= 512, "TMEM budget");
  static constexpr int NS = N + ((12 - N % 8) % 8);                 // bias row stride, NS % 8 == 4: float4 reads
  static_assert(NS % 8 == 4 && NS >= N, "bias stride");             //   of 8 consecutive rows hit 8 bank groups
  static constexpr uint32_t kRow = NPAD * 64;                       // one [NPAD][64 B] operand tile
  static constexpr uint32_t kStage = 3 * kRow;                      // Q | K | V
  static constexpr int TW = 2 * WS - 1, NTAB = TW * TW;
  // exchange buffers: row maxima [2][kKT][128] and row sums [2][kKT][128] (double-buffered by unit parity)
  static constexpr size_t kSmem = 1024 + 16 + (size_t)NSTAGE * kStage + (size_t)N * NS * 4 + (size_t)NTAB * 4 + 8 +
                                  (size_t)2 * 2 * kKT * 128 * 4;
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
// logits in log2 units (scale * cos + bias, shift mask), publishes its maximum, agrees on the row maximum with the
// two other warps of the lane quarter (named barrier `barid`, 96 threads), and writes P = exp2(s - m) as packed
// bf16 over the first NC/2 of ITS OWN S columns.  All lanes must call it (tcgen05.ld / st and the barrier are
// collective); `valid` lanes own a real row.  mx points at this row's entry of key quarter 0 (stride 128 floats).
template <int WS, int C0, int NC>
__device__ __forceinline__ void softmax_part(uint32_t t_s, const float* brow, float scale2, bool need_mask, uint32_t by,
                                             uint32_t bx, bool valid, float* mx, int kk, int barid, float& m_out, float& l_out) {
  constexpr int N = WS * WS;
  static_assert(C0 % 16 == 0 && NC % 16 == 0, "key quarters are whole k-steps");
  uint32_t sv[NC];
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_ld16(t_s + C0 + c * 16, &sv[c * 16]);
  ptx::tmem_ld_wait();
  float m = -INFINITY;
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
#pragma unroll
    for (int jj = 0; jj < NC; ++jj)
      if (C0 + jj < N) m = fmaxf(m, __uint_as_float(sv[jj]));
  }
  mx[kk * 128] = m;
  named_bar_sync(barid, kKT * 32);
  m = fmaxf(fmaxf(mx[0], mx[128]), mx[256]);
  float l0 = 0.f, l1 = 0.f;
  if (valid) {
#pragma unroll
    for (int jj = 0; jj < NC; jj += 2) {
      const float p0 = C0 + jj < N ? ex2(__uint_as_float(sv[jj]) - m) : 0.f;
      const float p1 = C0 + jj + 1 < N ? ex2(__uint_as_float(sv[jj + 1]) - m) : 0.f;
      l0 += p0;
      l1 += p1;
      sv[jj >> 1] = pack_bf16(p0, p1);
    }
  }
#pragma unroll
  for (int c = 0; c < NC / 16; ++c) tmem_st8(t_s + C0 + c * 8, &sv[c * 8]);    // over the start of this quarter's own S
  ptx::tmem_st_wait();
  m_out = m;
  l_out = l0 + l1;
}

template <int WS>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_ws2_kernel(const __grid_constant__ WsArgs a) {
  using CF = Cfg<WS>;
  constexpr int N = CF::N, NPAD = CF::NPAD, MT = CF::MT, NSLOT = CF::NSLOT, NS = CF::NS;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t kv_full[NSTAGE], kv_empty[NSTAGE];
  __shared__ __align__(8) uint64_t s_full[NSLOT], p_full[NSLOT], slot_free[NSLOT], o_full[2], o_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int s_hdr[NSTAGE][8];   // per ring stage: head, batch, window row / column, window index of the item (written by
                                     // the gather warps, which derive them anyway: the softmax warps do no divisions)

  const WinGeom& g = a.g;
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  float* bias = reinterpret_cast<float*>(sm + (size_t)NSTAGE * CF::kStage);
  float* tab = bias + N * NS;
  float* mxb = tab + CF::NTAB + (CF::NTAB & 1);        // [2][kKT][128] row maxima of the key thirds
  float* lsb = mxb + 2 * kKT * 128;                    // [2][kKT][128] row sums of the key thirds

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
      ptx::mbar_init(&p_full[s], kSoftmax / 32);
      ptx::mbar_init(&slot_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&o_full[s], 1); ptx::mbar_init(&o_free[s], kSoftmax / 32); }
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

  if (warp >= kSoftmaxWarps) {
    reg_dec<kRegService>();
    if (warp == kSIssuerWarp) {
      // =================================================================================== issuer of S = Q K^T
      if (lane == 0) {
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, NPAD, 0, 0);
        const uint64_t desc_k = ptx::make_smem_desc(0, 16, 512, kSw64);      // K-major Q / K tiles (64 B rows)
        const int U = n * MT;
        for (int u = 0; u < U; ++u) {
          const int slot = u % NSLOT;
          const int il = u / MT, tile = u - il * MT, stage = il % NSTAGE;
          TR(10, u);
          ptx::mbar_wait(&slot_free[slot], ((u / NSLOT) & 1) ^ 1);           // the PV MMAs that read this slot retired
          TR(11, u);
          if (tile == 0) ptx::mbar_wait(&kv_full[stage], (il / NSTAGE) & 1);
          TR(12, u);
          ptx::tc_fence_after();
          const uint32_t q_s = base_u32 + (uint32_t)stage * CF::kStage, k_s = q_s + CF::kRow;
          int row0 = tile * 128;
          if (CF::ROT && tile == MT - 1) row0 -= 32 * (il & 3);
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
          const int slot = u % NSLOT, ob = u & 1;
          const int il = u / MT, tile = u - il * MT, stage = il % NSTAGE;
          TR(14, u);
          ptx::mbar_wait(&p_full[slot], (u / NSLOT) & 1);
          TR(15, u);
          ptx::mbar_wait(&o_free[ob], ((u >> 1) & 1) ^ 1);                   // the epilogue of unit u - 2 has drained O
          TR(17, u);
          ptx::tc_fence_after();
          const uint32_t v_s = base_u32 + (uint32_t)stage * CF::kStage + 2 * CF::kRow;
          const uint32_t t_s = tmem_base + (uint32_t)slot * CF::SLOTW;
          const uint32_t t_o = tmem_base + CF::O_COL + (uint32_t)ob * HD;
          const uint64_t bd = desc_v + (v_s >> 4);
#pragma unroll
          for (int ks = 0; ks < NPAD / 16; ++ks) {
            // packed P of keys [16 ks, 16 ks + 16): at the start of the S columns of the key quarter that owns them
            const int key = ks * 16;
            const int c0 = key < CF::K1 ? CF::K0 : key < CF::K2 ? CF::K1 : CF::K2;
            ptx::mma_bf16_ts(t_o, t_s + (uint32_t)(c0 + (key - c0) / 2), bd + ks * 64, idesc_pv, ks != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&o_full[ob]);
          ptx::mma_commit(&slot_free[slot]);                                 // S / P of the slot are dead
          // every MMA reading this stage has retired: the S MMAs of the item (other thread) completed before the
          // softmax that produced this P could start
          if (tile == MT - 1) ptx::mma_commit(&kv_empty[stage]);
          TR(16, u);
        }
      }
    } else {
      // =================================================================================== gather warps
      const int lt = threadIdx.x - kSoftmax;
      const int C3 = 3 * a.C;
      // `pending` gathers are in flight (items i-pending .. i-1, one cp.async group each).  An item is published
      // (kv_full) as soon as it is LAG groups old -- and everything in flight is published before the warp goes to
      // sleep on a full ring, so the MMA warp can always run ahead on what has already landed.
      int pending = 0;
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
          const int64_t gi = g0 + i;
          const int h = (int)(gi / a.nwin);
          const int64_t win = gi - (int64_t)h * a.nwin;
          const int b = (int)(win / nW);
          const int w = (int)(win - (int64_t)b * nW);
          const int wh = w / g.nWw, ww = w - wh * g.nWw;
          unsigned char* st = sm + (size_t)stage * CF::kStage;
          if (lt == 0) {
            s_hdr[stage][0] = h; s_hdr[stage][1] = b; s_hdr[stage][2] = wh; s_hdr[stage][3] = ww; s_hdr[stage][4] = (int)win;
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
    const int q = warp & 3;                         // TMEM lane quarter of this warp (= SM sub-partition)
    const int kk = warp >> 2;                       // key third of this warp
    const int st = threadIdx.x;                     // 0..383 over all softmax warps
    const int rowl = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int barid = 2 + q;                        // named barrier of the three warps that share this lane quarter
    int cur_head = -1;
    float scale2 = 0.f;

    // Deferred epilogue of unit u: O / l -> out (natural [B,H,W,C] layout), 16 of the 32 columns per thread of the
    // first two key thirds.  Everything
    // it needs about the unit was computed when the unit's softmax ran (`t`: token of this thread's row or -1, `lse_i`:
    // index of the row's log-sum-exp or -1, `h`: head): no index arithmetic is repeated here.
    auto epilogue = [&](int u, float m, int t, int lse_i, int h) {
      const int ob = u & 1;
      TR(33, u);
      ptx::mbar_wait(&o_full[ob], (u >> 1) & 1);
      TR(35, u);
      ptx::tc_fence_after();
      const float* lsr = lsb + ob * kKT * 128 + rowl;
      if (kk < 2) {
        // key thirds 0 and 1: 16 columns of O each
        uint32_t o[16];
        ptx::tmem_ld_32x32b_x16(t_lane + CF::O_COL + (uint32_t)ob * HD + (uint32_t)kk * 16, o);
        const float lt = (lsr[0] + lsr[128]) + lsr[256];
        ptx::tmem_ld_wait();
        if (t >= 0) {
          const float inv = __fdividef(1.0f, lt);
          uint4* dst = reinterpret_cast<uint4*>(a.out + (int64_t)t * a.C + h * HD + kk * 16);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            dst[c] = make_uint4(pack_bf16(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv),
                                pack_bf16(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv),
                                pack_bf16(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv),
                                pack_bf16(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv));
        }
      } else if (lse_i >= 0) {
        // key third 2: the row's log-sum-exp.  m: the row maximum of this thread's own softmax of unit u (the exchange
        // buffer may already hold unit u + 2)
        const float lt = (lsr[0] + lsr[128]) + lsr[256];
        a.lse[lse_i] = (m + log2f(lt)) * kLn2;
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&o_free[ob]);
      TR(34, u);
    };

    int pend_u = -1, pend_t = -1, pend_lse = -1, pend_h = 0;
    float pend_m = 0.f;
    int h = 0, win = 0, b = 0, wh = 0, ww = 0;
    bool last_h = false, last_w = false, need_mask = false;
    for (int il = 0; il < n; ++il) {
#pragma unroll 1
      for (int tile = 0; tile < MT; ++tile) {
        const int u = il * MT + tile;
        const int slot = u % NSLOT;
        const uint32_t par = (u / NSLOT) & 1;
        const bool rot_tile = CF::ROT && tile == MT - 1;
        const bool quarter_active = !rot_tile || q == (il & 3);           // uniform over the four warps of the quarter
        int r;                                                            // in-window query row of this thread, or -1
        if (rot_tile) r = (quarter_active && lane < CF::TAIL) ? tile * 128 + lane : -1;
        else { r = tile * 128 + rowl; if (r >= N) r = -1; }
        const uint32_t t_s = t_lane + (uint32_t)slot * CF::SLOTW;
        float m = 0.f;

        TR(30, u);
        ptx::mbar_wait(&s_full[slot], par);
        TR(31, u);
        ptx::tc_fence_after();
        if (tile == 0) {
          // new item: its geometry from the header of its ring stage (S of the item exists, so the gather has landed)
          const int* hdr = s_hdr[il % NSTAGE];
          h = hdr[0]; b = hdr[1]; wh = hdr[2]; ww = hdr[3]; win = hdr[4];
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
          last_h = wh == g.nWh - 1;
          last_w = ww == g.nWw - 1;
          need_mask = g.shift > 0 && (last_h || last_w);          // CTA-uniform per item
        }
        if (quarter_active) {
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
          float* mx = mxb + (u & 1) * kKT * 128 + rowl;
          float l = 0.f;
          if (kk == 0) softmax_part<WS, CF::K0, CF::K1 - CF::K0>(t_s, brow, scale2, need_mask, by, bx, r >= 0, mx, kk, barid, m, l);
          else if (kk == 1) softmax_part<WS, CF::K1, CF::K2 - CF::K1>(t_s, brow, scale2, need_mask, by, bx, r >= 0, mx, kk, barid, m, l);
          else softmax_part<WS, CF::K2, CF::K3 - CF::K2>(t_s, brow, scale2, need_mask, by, bx, r >= 0, mx, kk, barid, m, l);
          lsb[(u & 1) * kKT * 128 + kk * 128 + rowl] = l;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[slot]);
        TR(32, u);
        // epilogue of the previous unit: its PV MMAs ran while this unit's softmax was computed
        if (pend_u >= 0) epilogue(pend_u, pend_m, pend_t, pend_lse, pend_h);
        pend_u = u;
        pend_m = m;
        pend_h = h;
        pend_lse = r >= 0 ? (win * a.nH + h) * N + r : -1;
        pend_t = r >= 0 ? src_token(g, b, wh, ww, r / WS, r % WS) : -1;
      }
    }
    if (pend_u >= 0) epilogue(pend_u, pend_m, pend_t, pend_lse, pend_h);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kIssuerWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

template <int WS>
int launch_ws2(const WsArgs& a, cudaStream_t st) {
  using CF = Cfg<WS>;
  static_assert(CF::kSmem <= 227 * 1024, "shared memory budget");
  BSW_CUDA(cudaFuncSetAttribute(attn_fwd_ws2_kernel<WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CF::kSmem));
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
  attn_fwd_ws2_kernel<WS><<<(unsigned)grid, kThreads, CF::kSmem, st>>>(at);
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
  attn_fwd_ws2_kernel<WS><<<(unsigned)grid, kThreads, CF::kSmem, st>>>(a);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
#endif
}
}  // namespace

bool attn_fwd_ws2_supported(int ws) { return ws == 12; }

int attn_fwd_ws2(const void* qkv, void* out, float* lse, const float* table16, const float* scale, const float* qpad,
                 const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st) {
  WsArgs a;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.lse = lse;
  a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  make_geom(&a.g, B, H, W, ws, shift);
  a.C = C; a.nH = nH;
  a.nwin = (int64_t)B * a.g.nWh * a.g.nWw;
  a.nitems = a.nwin * nH;
  a.trace = nullptr;
  BSW_REQUIRE(a.nitems * (int64_t)(ws * ws) < (1ll << 31), "attn_fwd(ws2): too many (window, head, row) triples for 32-bit indices");
  if (ws == 12) return launch_ws2<12>(a, st);
  set_error("attn_fwd(ws2): window %dx%d not instantiated", ws, ws);
  return B200SWIN_EINVAL;
}

}  // namespace b200swin
