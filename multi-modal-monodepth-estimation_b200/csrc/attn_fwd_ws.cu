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
//   warps 4-11  two softmax warpgroups on alternating items, one thread per query row: the whole fp32 row
//               of S lives in registers (setmaxnreg), bias comes from a per-head [N][N] fp32 matrix in shared
//               memory (conflict-free float4 reads), the shift mask from per-thread bit masks (only windows on
//               the last window row / column pay for it), exp2 on the SFU, P written back to TMEM as packed bf16
//               over the S columns, O scaled by 1/rowsum and stored to the natural [B,H,W,C] layout.
//
// A 12x12 window has 144 = 128 + 16 rows.  The 16-row tail is a second M=128 MMA whose A operand starts
// 32*rot rows early, so the tail lands in TMEM lane quarter `rot` -- rot rotates per item and the extra softmax
// pass is spread over the four SM sub-partitions instead of always hitting the first.
//
// Items are ordered head-major and every CTA owns one contiguous range, so the expanded bias matrix is rebuilt
// only when the head changes (at most a few times per CTA).
#include "common.cuh"
#include "wingeom.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

namespace {
constexpr int HD = 32;
constexpr int kThreads = 384;
constexpr int kLoaders = 96;
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
};

template <int WS>
struct Cfg {
  static constexpr int N = WS * WS;
  static constexpr int NPAD = (N + 15) / 16 * 16;
  static constexpr int MT = (NPAD + 127) / 128;
  static constexpr int TAIL = N - 128 * (MT - 1);                  // valid rows of the last tile
  static constexpr bool ROT = MT > 1 && TAIL <= 32;
  static constexpr int SLOTW = NPAD < 64 ? 64 : NPAD;               // TMEM columns per slot: S, then P + O over it
  static constexpr int NSLOT = 512 / SLOTW > 4 ? 4 : 512 / SLOTW;
  static constexpr int OCOL = (NPAD / 2 + 31) / 32 * 32;            // O accumulator behind the packed P
  static_assert(OCOL + HD <= SLOTW, "O must fit into the slot");
  static constexpr int NS = N + ((12 - N % 8) % 8);                 // bias row stride, NS % 8 == 4: float4 reads
  static_assert(NS % 8 == 4 && NS >= N, "bias stride");             //   of 8 consecutive rows hit 8 bank groups
  static constexpr uint32_t kRow = NPAD * 64;                       // one [NPAD][64 B] operand tile
  static constexpr uint32_t kStage = 3 * kRow;                      // Q | K | V
  static constexpr int TW = 2 * WS - 1, NTAB = TW * TW;
  static constexpr size_t kSmem = 1024 + (size_t)NSTAGE * kStage + (size_t)N * NS * 4 + (size_t)NTAB * 4;
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

template <int WS>
__global__ void __launch_bounds__(kThreads, 1)
attn_fwd_ws_kernel(const __grid_constant__ WsArgs a) {
  using CF = Cfg<WS>;
  constexpr int N = CF::N, NPAD = CF::NPAD, MT = CF::MT, NSLOT = CF::NSLOT, NS = CF::NS;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t kv_full[NSTAGE], kv_empty[NSTAGE];
  __shared__ __align__(8) uint64_t s_full[NSLOT], p_full[NSLOT], o_full[NSLOT], slot_free[NSLOT];
  __shared__ uint32_t tmem_slot;

  const WinGeom& g = a.g;
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  float* bias = reinterpret_cast<float*>(sm + (size_t)NSTAGE * CF::kStage);
  float* tab = bias + N * NS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // contiguous, balanced item range of this CTA (items are head-major)
  const int64_t per = a.nitems / gridDim.x, rem = a.nitems % gridDim.x;
  const int64_t g0 = (int64_t)blockIdx.x * per + min((int64_t)blockIdx.x, rem);
  const int n = (int)(per + ((int64_t)blockIdx.x < rem ? 1 : 0));
  const int nW = g.nWh * g.nWw;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) { ptx::mbar_init(&kv_full[s], kLoaders); ptx::mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < NSLOT; ++s) {
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&p_full[s], 4);
      ptx::mbar_init(&o_full[s], 1);
      ptx::mbar_init(&slot_free[s], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      // =================================================================================== MMA issuer
      if (lane == 0) {
        constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, NPAD, 0, 0);
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, HD, 0, 1);   // A = P (TMEM), B = V MN-major
        const int U = n * MT;
        int su = 0, pu = 0;
        long long t0 = clock64();
        while (pu < U) {
          bool progressed = false;
          if (pu < su) {
            const int slot = pu % NSLOT;
            if (ptx::mbar_test_wait(&p_full[slot], (pu / NSLOT) & 1)) {
              const int il = pu / MT, tile = pu - il * MT, stage = il % NSTAGE;
              ptx::tc_fence_after();
              const uint32_t v_s = base_u32 + (uint32_t)stage * CF::kStage + 2 * CF::kRow;
              const uint32_t t_s = tmem_base + (uint32_t)slot * CF::SLOTW;
#pragma unroll
              for (int ks = 0; ks < NPAD / 16; ++ks)
                ptx::mma_bf16_ts(t_s + CF::OCOL, t_s + ks * 8, ptx::make_smem_desc(v_s + ks * 1024, 512, 512, kSw64),
                                 idesc_pv, ks);
              ptx::mma_commit(&o_full[slot]);
              if (tile == MT - 1) ptx::mma_commit(&kv_empty[stage]);   // every MMA that reads this stage has retired
              ++pu;
              progressed = true;
            }
          }
          if (su < U && su - pu < NSLOT) {
            const int slot = su % NSLOT;
            const int il = su / MT, tile = su - il * MT, stage = il % NSTAGE;
            if (ptx::mbar_test_wait(&slot_free[slot], ((su / NSLOT) & 1) ^ 1) &&
                ptx::mbar_test_wait(&kv_full[stage], (il / NSTAGE) & 1)) {
              ptx::tc_fence_after();
              const uint32_t q_s = base_u32 + (uint32_t)stage * CF::kStage, k_s = q_s + CF::kRow;
              int row0 = tile * 128;
              if (CF::ROT && tile == MT - 1) row0 -= 32 * ((il >> 1) & 3);
              const uint32_t t_s = tmem_base + (uint32_t)slot * CF::SLOTW;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                ptx::mma_bf16_ss(t_s, ptx::make_smem_desc(q_s + row0 * 64 + ks * 32, 16, 512, kSw64),
                                 ptx::make_smem_desc(k_s + ks * 32, 16, 512, kSw64), idesc_qk, ks);
              ptx::mma_commit(&s_full[slot]);
              ++su;
              progressed = true;
            }
          }
          if (progressed) t0 = clock64();
          else if (clock64() - t0 > 4000000000ll) __trap();
        }
      }
    } else {
      // =================================================================================== gather warps
      const int lt = threadIdx.x - 32;
      const int C3 = 3 * a.C;
      for (int i = 0; i < n + LAG; ++i) {
        if (i < n) {
          const int stage = i % NSTAGE;
          ptx::mbar_wait(&kv_empty[stage], ((i / NSTAGE) & 1) ^ 1);
          const int64_t gi = g0 + i;
          const int h = (int)(gi / a.nwin);
          const int64_t win = gi - (int64_t)h * a.nwin;
          const int b = (int)(win / nW);
          const int w = (int)(win - (int64_t)b * nW);
          const int wh = w / g.nWw, ww = w - wh * g.nWw;
          unsigned char* st = sm + (size_t)stage * CF::kStage;
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
        if (i >= LAG) {
          ptx::cp_async_wait<LAG>();              // the gather of item i - LAG has landed
          ptx::fence_proxy_async_smem();          // generic-proxy writes -> visible to tcgen05.mma
          ptx::mbar_arrive(&kv_full[(i - LAG) % NSTAGE]);
        }
      }
    }
  } else {
    // ===================================================================================== softmax warpgroups
    reg_inc<224>();
    const int wg = (warp - 4) >> 2;                 // 0 / 1: items of even / odd local index
    const int q = warp & 3;                         // TMEM lane quarter of this warp
    const int st = threadIdx.x - 128;               // 0..255 over both softmax warpgroups
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int cur_head = -1;
    float scale2 = 0.f;

    for (int il = 0; il < n; ++il) {
      const int64_t gi = g0 + il;
      const int h = (int)(gi / a.nwin);
      if (h != cur_head) {
        // both warpgroups have finished every earlier item: rebuild the expanded bias matrix of head h (log2 units)
        named_bar_sync(1, 256);
        for (int t = st; t < CF::NTAB; t += 256) tab[t] = a.table16[t * a.nH + h] * kLog2e;
        named_bar_sync(1, 256);
        for (int e = st; e < N * N; e += 256) {
          const int i = e / N, j = e - i * N;
          bias[i * NS + j] = tab[(i / WS - j / WS + WS - 1) * CF::TW + (i % WS - j % WS + WS - 1)];
        }
        named_bar_sync(1, 256);
        cur_head = h;
        scale2 = a.scale[h] * kLog2e;
      }
      if ((il & 1) != wg) continue;
      const int64_t win = gi - (int64_t)h * a.nwin;
      const int b = (int)(win / nW);
      const int w = (int)(win - (int64_t)b * nW);
      const int wh = w / g.nWw, ww = w - wh * g.nWw;
      const bool last_h = wh == g.nWh - 1, last_w = ww == g.nWw - 1;
      const bool need_mask = g.shift > 0 && (last_h || last_w);          // CTA-uniform per item

#pragma unroll 1
      for (int tile = 0; tile < MT; ++tile) {
        const int u = il * MT + tile;
        const int slot = u % NSLOT;
        const uint32_t par = (u / NSLOT) & 1;
        const bool rot_tile = CF::ROT && tile == MT - 1;
        int r;                                                            // in-window query row of this thread, or -1
        if (rot_tile) r = (q == ((il >> 1) & 3) && lane < CF::TAIL) ? tile * 128 + lane : -1;
        else { r = tile * 128 + q * 32 + lane; if (r >= N) r = -1; }
        const bool warp_active = __any_sync(0xffffffffu, r >= 0);
        const uint32_t t_s = t_lane + (uint32_t)slot * CF::SLOTW;

        ptx::mbar_wait(&s_full[slot], par);
        ptx::tc_fence_after();
        float m = -INFINITY, l0 = 0.f, l1 = 0.f;
        if (warp_active) {
          uint32_t sv[NPAD];
#pragma unroll
          for (int c = 0; c < NPAD / 16; ++c) tmem_ld16(t_s + c * 16, &sv[c * 16]);
          ptx::tmem_ld_wait();
          if (r >= 0) {
            const int yi = r / WS, xi = r - yi * WS;
            const float4* brow = reinterpret_cast<const float4*>(bias + r * NS);
            // ---- pass A: logits in log2 units and the row maximum
#pragma unroll
            for (int j4 = 0; j4 < (N + 3) / 4; ++j4) {
              const float4 bb = brow[j4];
              const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int j = j4 * 4 + k;
                if (j < N) sv[j] = __float_as_uint(fmaf(__uint_as_float(sv[j]), scale2, bv[k]));
              }
            }
            if (need_mask) {
              // regions differ only across the roll seam of the last window row / column
              const int cut = WS - g.shift;                                 // in-window coordinate of the seam
              const uint32_t hi = (~0u << cut) & ((1u << WS) - 1u), lo = (1u << cut) - 1u;
              const uint32_t by = last_h ? (yi >= cut ? lo : hi) : 0u;      // bit y set: key row y is masked
              const uint32_t bx = last_w ? (xi >= cut ? lo : hi) : 0u;
#pragma unroll
              for (int yj = 0; yj < WS; ++yj) {
                const uint32_t rowm = ((by >> yj) & 1u) ? 0xffffffffu : bx;
#pragma unroll
                for (int xj = 0; xj < WS; ++xj) {
                  const int j = yj * WS + xj;
                  if ((rowm >> xj) & 1u) sv[j] = __float_as_uint(__uint_as_float(sv[j]) + kMaskLog2);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < N; ++j) m = fmaxf(m, __uint_as_float(sv[j]));
            // ---- pass B: P = exp2(s - m), row sum, packed bf16 over the S columns
#pragma unroll
            for (int j = 0; j < NPAD; j += 2) {
              const float p0 = j < N ? ex2(__uint_as_float(sv[j]) - m) : 0.f;
              const float p1 = j + 1 < N ? ex2(__uint_as_float(sv[j + 1]) - m) : 0.f;
              l0 += p0;
              l1 += p1;
              sv[j >> 1] = pack_bf16(p0, p1);
            }
          }
#pragma unroll
          for (int c = 0; c < NPAD / 16; ++c) tmem_st8(t_s + c * 8, &sv[c * 8]);
          ptx::tmem_st_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[slot]);

        // ---- O = P V is on its way; read it back, release the slot, normalise and store
        ptx::mbar_wait(&o_full[slot], par);
        ptx::tc_fence_after();
        uint32_t o[32];
        if (warp_active) {
          ptx::tmem_ld_32x32b_x32(t_s + CF::OCOL, o);
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&slot_free[slot]);
        if (r >= 0) {
          const float l = l0 + l1;
          a.lse[((int64_t)win * a.nH + h) * N + r] = (m + log2f(l)) * kLn2;
          const int t = src_token(g, b, wh, ww, r / WS, r % WS);
          if (t >= 0) {
            const float inv = 1.0f / l;
            uint4* dst = reinterpret_cast<uint4*>(a.out + (int64_t)t * a.C + h * HD);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 wv;
              wv.x = pack_bf16(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv);
              wv.y = pack_bf16(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv);
              wv.z = pack_bf16(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv);
              wv.w = pack_bf16(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv);
              dst[c] = wv;
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
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
  attn_fwd_ws_kernel<WS><<<(unsigned)grid, kThreads, CF::kSmem, st>>>(a);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
}  // namespace

bool attn_fwd_ws_supported(int ws) { return ws == 4 || ws == 6 || ws == 7 || ws == 8 || ws == 12; }

int attn_fwd_ws(const void* qkv, void* out, float* lse, const float* table16, const float* scale, const float* qpad,
                const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st) {
  WsArgs a;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.lse = lse;
  a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  make_geom(&a.g, B, H, W, ws, shift);
  a.C = C; a.nH = nH;
  a.nwin = (int64_t)B * a.g.nWh * a.g.nWw;
  a.nitems = a.nwin * nH;
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
