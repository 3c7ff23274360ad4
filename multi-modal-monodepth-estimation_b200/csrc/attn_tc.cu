// Attention core on tcgen05 tensor cores (impl = 1, bf16 storage, windows of up to 256 tokens).
//
// Replaces models/swin_transformer_v2.py:295-328 (cosine logits * clamped scale + CPB bias + shift mask ->
// softmax -> P @ V) with the block's pad/roll/partition/reverse/crop (:429-463) and BasicLayer's shift mask
// (:874-892) folded into the load/store addressing.  One work item = one (window, head):
//
//   gather  q_hat,k_hat,v rows of the window from the NATURAL [B,H,W,3C] tensor with cp.async (16 B chunks)
//           straight into the 64-byte-swizzled UMMA operand layout; next item prefetched (double buffer);
//   S   =   Q[128 x 32] . K^T[32 x N]      tcgen05.mma (2 k-steps)  -> fp32 in TMEM, never in HBM;
//   softmax one thread per row (TMEM lane): bias from the 16*sigmoid table in smem, shift mask from token
//           coordinates, exp2 in fp32; P written back to TMEM as packed bf16 over the S columns;
//   O   =   P[128 x N] (TMEM) . V[N x 32] (smem, MN-major)   tcgen05.mma (N/16 k-steps) -> TMEM;
//   store   O / rowsum as bf16 to the natural [B,H,W,C] layout (window_reverse + un-roll + crop = address
//           math), log-sum-exp per row for the backward.
//
// Windows with more than 128 tokens (ws=12 -> 144) run a second 128-row tile for the remaining rows.
#include <stdlib.h>
#include "common.cuh"
#include "wingeom.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

namespace {
constexpr int HD = 32;
constexpr int kThreads = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kMaskLog2 = -100.0f * 1.4426950408889634f;
constexpr uint32_t kSw64 = 4;       // UMMA layout type SWIZZLE_64B

struct TcArgs {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  float* lse;
  const float* table16;
  const float* scale;
  const float* qpad;
  const float* vpad;
  WinGeom g;
  int C, nH;
  int64_t nitems;     // nwin * nH, item = win * nH + head
};

// byte offset of 16-byte chunk `c` of row `r` in a [rows][64 B] tile with the 64 B swizzle (Swizzle<2,4,3>)
__device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int NPAD>
struct FwdLayout {
  static constexpr int MT = (NPAD + 127) / 128;             // 128-row query tiles per window
  static constexpr int QROWS = MT * 128;
  static constexpr uint32_t kQBytes = QROWS * 64, kKBytes = NPAD * 64;
  static constexpr uint32_t kBufBytes = kQBytes + 2 * kKBytes;           // Q | K | V of one item
  static constexpr uint32_t kTmemCols = (NPAD + HD <= 64) ? 64 : (NPAD + HD <= 128) ? 128 : (NPAD + HD <= 256) ? 256 : 512;
};

// Issue the gather of one item's q_hat / k_hat / v rows into buffer `buf` (cp.async; pads by st.shared).
template <int NPAD>
__device__ __forceinline__ void load_item(const TcArgs& a, int64_t item, unsigned char* buf, int* tok, int N) {
  using LY = FwdLayout<NPAD>;
  const int64_t win = item / a.nH;
  const int h = (int)(item - win * a.nH);
  const uint32_t q_s = ptx::smem_u32(buf), k_s = q_s + LY::kQBytes, v_s = k_s + LY::kKBytes;
  const int C3 = 3 * a.C;
  for (int r = threadIdx.x; r < NPAD; r += kThreads) {
    int t = -2;                                  // -2: key padding row (beyond the window), -1: pad token
    if (r < N) {
      int b, i, j, si, sj;
      bool real = win_token(a.g, win, r, b, i, j, si, sj);
      t = real ? ((b * a.g.H + i) * a.g.W + j) : -1;
    }
    tok[r] = t;
    if (t >= 0) {
      const __nv_bfloat16* src = a.qkv + (int64_t)t * C3 + h * HD;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ptx::cp_async_16(q_s + sw64_off(r, c), src + c * 8);
        ptx::cp_async_16(k_s + sw64_off(r, c), src + a.C + c * 8);
        ptx::cp_async_16(v_s + sw64_off(r, c), src + 2 * a.C + c * 8);
      }
    } else {
      // pad token: q = normalised q_bias, k = 0, v = v_bias; key padding rows: all zero
#pragma unroll
      for (int c = 0; c < 4; ++c) {
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
        *reinterpret_cast<uint4*>(buf + sw64_off(r, c)) = qv;
        *reinterpret_cast<uint4*>(buf + LY::kQBytes + sw64_off(r, c)) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(buf + LY::kQBytes + LY::kKBytes + sw64_off(r, c)) = vv;
      }
    }
  }
}

template <int NPAD>
__global__ void __launch_bounds__(kThreads)
attn_fwd_tc_kernel(const __grid_constant__ TcArgs a) {
  using LY = FwdLayout<NPAD>;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar_s, bar_o;
  __shared__ uint32_t tmem_slot;

  const WinGeom& g = a.g;
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1, ntab = tw * tw;
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  unsigned char* bufs[2] = {sm, sm + LY::kBufBytes};
  float* tab = reinterpret_cast<float*>(sm + 2 * LY::kBufBytes);
  int* toks[2] = {reinterpret_cast<int*>(tab + ntab), reinterpret_cast<int*>(tab + ntab) + NPAD};
  int* meta = toks[1] + NPAD;        // [NPAD] koff | region << 16 of the CURRENT item's window

  const int warp = threadIdx.x >> 5;
  // zero the Q rows beyond the window once (tile 1 reads 128 rows; they must at least be finite)
  for (int i = threadIdx.x; i < 2 * (int)LY::kBufBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_s, 1);
    ptx::mbar_init(&bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, LY::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);     // this warp's lane quarter
  constexpr uint32_t S_COL = 0, O_COL = NPAD;                            // P (packed bf16) aliases the S columns

  constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, NPAD, 0, 0);
  constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, HD, 0, 1);     // A = P (TMEM, K-major), B = V MN-major

  uint32_t ph_s = 0, ph_o = 0;
  int cur_head = -1;
  int64_t item = blockIdx.x;
  int it = 0;
  if (item < a.nitems) load_item<NPAD>(a, item, bufs[0], toks[0], N);
  ptx::cp_async_commit();

  for (; item < a.nitems; item += gridDim.x, ++it) {
    const int b = it & 1;
    const int64_t win = item / a.nH;
    const int h = (int)(item - win * a.nH);
    const int64_t nxt = item + gridDim.x;
    if (nxt < a.nitems) load_item<NPAD>(a, nxt, bufs[b ^ 1], toks[b ^ 1], N);
    ptx::cp_async_commit();
    // per-window metadata and (when the head changes) the bias table, in log2 units
    for (int r = threadIdx.x; r < NPAD; r += kThreads) {
      int region = 0;
      if (g.shift > 0 && r < N) {
        int bb, i, j, si, sj;
        win_token(g, win, r, bb, i, j, si, sj);
        region = 3 * region_1d(si, g.Hp, ws, g.shift) + region_1d(sj, g.Wp, ws, g.shift);
      }
      meta[r] = (r < N ? ((r / ws) * tw + (r % ws)) : 0) | (region << 16);
    }
    if (h != cur_head) {
      for (int r = threadIdx.x; r < ntab; r += kThreads) tab[r] = a.table16[r * a.nH + h] * kLog2e;
      cur_head = h;
    }
    ptx::cp_async_wait<1>();                 // this item's gather has landed (the prefetch may still fly)
    ptx::fence_proxy_async_smem();           // st.shared / cp.async data -> visible to tcgen05.mma
    __syncthreads();

    const float scale2 = a.scale[h] * kLog2e;
    const uint32_t q_s = ptx::smem_u32(bufs[b]), k_s = q_s + LY::kQBytes, v_s = k_s + LY::kKBytes;
    const int* tok = toks[b];

#pragma unroll 1
    for (int tile = 0; tile < LY::MT; ++tile) {
      if (tile * 128 >= N) break;
      // ---- S = Q_tile . K^T
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t ad = ptx::make_smem_desc(q_s + tile * 128 * 64 + ks * 32, 16, 512, kSw64);
          const uint64_t bd = ptx::make_smem_desc(k_s + ks * 32, 16, 512, kSw64);
          ptx::mma_bf16_ss(tmem_base + S_COL, ad, bd, idesc_qk, ks);
        }
        ptx::mma_commit(&bar_s);
      }
      ptx::mbar_wait(&bar_s, ph_s);
      ph_s ^= 1;
      ptx::tc_fence_after();

      // ---- softmax over the row owned by this thread
      const int r = tile * 128 + threadIdx.x;
      const int rr = r < N ? r : 0;                         // rows beyond the window compute garbage, never stored
      const int base_i = (meta[rr] & 0xffff) + (ws - 1) * (tw + 1);
      const int reg_i = meta[rr] >> 16;
      float m = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < NPAD / 16; ++c) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(t_row + S_COL + c * 16, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int j = c * 16 + u;
          const int mj = meta[j];
          float s = fmaf(__uint_as_float(v[u]), scale2, tab[base_i - (mj & 0xffff)]);
          if ((mj >> 16) != reg_i) s += kMaskLog2;
          if (NPAD != N && j >= N) s = -INFINITY;
          m = fmaxf(m, s);
          v[u] = __float_as_uint(s);
        }
        ptx::tmem_st_32x32b_x16(t_row + S_COL + c * 16, v);
      }
      ptx::tmem_st_wait();
      float l = 0.f;
#pragma unroll 1
      for (int c = 0; c < NPAD / 16; ++c) {
        uint32_t v[16], pk[8];
        ptx::tmem_ld_32x32b_x16(t_row + S_COL + c * 16, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
          const float p0 = exp2f(__uint_as_float(v[u]) - m);
          const float p1 = exp2f(__uint_as_float(v[u + 1]) - m);
          l += p0 + p1;
          pk[u >> 1] = pack_bf16(p0, p1);
        }
        ptx::tmem_st_32x32b_x8(t_row + S_COL + c * 8, pk);   // P chunk c lands on columns already consumed
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncthreads();

      // ---- O = P . V
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ks = 0; ks < NPAD / 16; ++ks) {
          const uint64_t bd = ptx::make_smem_desc(v_s + ks * 1024, 512, 512, kSw64);
          ptx::mma_bf16_ts(tmem_base + O_COL, tmem_base + S_COL + ks * 8, bd, idesc_pv, ks);
        }
        ptx::mma_commit(&bar_o);
      }
      ptx::mbar_wait(&bar_o, ph_o);
      ph_o ^= 1;
      ptx::tc_fence_after();
      {
        uint32_t o[32];
        ptx::tmem_ld_32x32b_x32(t_row + O_COL, o);
        ptx::tmem_ld_wait();
        if (r < N) {
          a.lse[(win * a.nH + h) * N + r] = (m + log2f(l)) * kLn2;
          const int t = tok[r];
          if (t >= 0) {
            const float inv = 1.0f / l;
            uint4* dst = reinterpret_cast<uint4*>(a.out + (int64_t)t * a.C + h * HD);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(o[c * 8 + 0]) * inv, __uint_as_float(o[c * 8 + 1]) * inv);
              w.y = pack_bf16(__uint_as_float(o[c * 8 + 2]) * inv, __uint_as_float(o[c * 8 + 3]) * inv);
              w.z = pack_bf16(__uint_as_float(o[c * 8 + 4]) * inv, __uint_as_float(o[c * 8 + 5]) * inv);
              w.w = pack_bf16(__uint_as_float(o[c * 8 + 6]) * inv, __uint_as_float(o[c * 8 + 7]) * inv);
              dst[c] = w;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncthreads();                       // all lanes done with S/P/O before the next MMA overwrites them
    }
  }
  ptx::cp_async_wait<0>();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, LY::kTmemCols);
  }
}

template <int NPAD>
static int launch_fwd(const TcArgs& a, cudaStream_t st) {
  using LY = FwdLayout<NPAD>;
  const int ws = a.g.ws, ntab = (2 * ws - 1) * (2 * ws - 1);
  size_t smem = 1024 + 2 * (size_t)LY::kBufBytes + (size_t)ntab * 4 + 3 * (size_t)NPAD * 4;
  BSW_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)(512 / LY::kTmemCols);
  int by_smem = (int)((227 * 1024) / (smem + 1024));
  if (by_smem < per_sm) per_sm = by_smem;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > a.nitems) grid = a.nitems;
  attn_fwd_tc_kernel<NPAD><<<(unsigned)grid, kThreads, smem, st>>>(a);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
}  // namespace

// warp-specialised forward (attn_fwd_ws.cu)
bool attn_fwd_ws_supported(int ws);
int attn_fwd_ws(const void* qkv, void* out, float* lse, const float* table16, const float* scale, const float* qpad,
                const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st);
// second-generation warp-specialised forward for 12x12 windows (attn_fwd_ws2.cu)
bool attn_fwd_ws2_supported(int ws);
int attn_fwd_ws2(const void* qkv, void* out, float* lse, const float* table16, const float* scale, const float* qpad,
                 const float* vpad, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st);
// Opt-in (B200SWIN_ATTN_FWD_GEN2=1): parity-green, but on B200 still 10 % slower than the first generation (201 us vs
// 180 us for Swin-B stage 2) -- see the header of attn_fwd_ws2.cu and DESIGN.md section 8.
static bool use_gen2_fwd() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200SWIN_ATTN_FWD_GEN2"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
// warp-specialised backward (attn_bwd_ws.cu)
bool attn_bwd_ws_supported(int ws);
size_t attn_bwd_ws_workspace_bytes(int B, int H, int W, int nH);
int attn_bwd_ws(const void* qkv, const void* out, const void* dout, const float* lse, const float* inv_norm,
                const float* table16, const float* scale, const float* qpad, const float* vpad, void* dqkv,
                float* dtable16, float* dscale, float* dvpad, void* workspace, int B, int H, int W, int C, int nH, int ws,
                int shift, cudaStream_t st);
static bool use_legacy_bwd() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200SWIN_ATTN_BWD_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static bool use_legacy_fwd() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200SWIN_ATTN_FWD_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

bool attn_tc_supported(int ws, int C, int nH, const void* mask) {
  const int N = ws * ws;
  return mask == nullptr && C == nH * HD && N <= 256 && N >= 4;
}

int attn_fwd_tc(const void* qkv, void* out, float* lse, const float* table16, const float* scale, const float* qpad,
                const float* vpad, const float* mask, int nWm, int B, int H, int W, int C, int nH, int ws, int shift,
                cudaStream_t st) {
  (void)nWm;
  BSW_REQUIRE(attn_tc_supported(ws, C, nH, mask),
              "attn_fwd(tc): needs head_dim 32, window <= 16x16 and the on-the-fly mask (no explicit mask tensor)");
  BSW_REQUIRE(shift >= 0 && shift < ws, "attn_fwd(tc): bad shift");
  BSW_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && C % 8 == 0,
              "attn_fwd(tc): qkv/out must be 16-byte aligned");
  BSW_REQUIRE((int64_t)B * H * W < (1ll << 29), "attn_fwd(tc): too many tokens");
  if (attn_fwd_ws2_supported(ws) && !use_legacy_fwd() && use_gen2_fwd())
    return attn_fwd_ws2(qkv, out, lse, table16, scale, qpad, vpad, B, H, W, C, nH, ws, shift, st);
  if (attn_fwd_ws_supported(ws) && !use_legacy_fwd())
    return attn_fwd_ws(qkv, out, lse, table16, scale, qpad, vpad, B, H, W, C, nH, ws, shift, st);
  TcArgs a;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (__nv_bfloat16*)out; a.lse = lse;
  a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  make_geom(&a.g, B, H, W, ws, shift);
  a.C = C; a.nH = nH;
  a.nitems = (int64_t)B * a.g.nWh * a.g.nWw * nH;
  const int N = ws * ws;
  const int npad = (N + 15) / 16 * 16;
  switch (npad) {
    case 16: return launch_fwd<16>(a, st);
    case 32: return launch_fwd<32>(a, st);
    case 48: return launch_fwd<48>(a, st);
    case 64: return launch_fwd<64>(a, st);
    case 96: return launch_fwd<96>(a, st);      // never hit by square windows; kept for completeness of the switch
    case 112: return launch_fwd<112>(a, st);    // ws = 10 -> 100 -> 112
    case 128: return launch_fwd<128>(a, st);    // ws = 11 -> 121 -> 128
    case 144: return launch_fwd<144>(a, st);
    case 176: return launch_fwd<176>(a, st);    // ws = 13
    case 208: return launch_fwd<208>(a, st);    // ws = 14
    case 240: return launch_fwd<240>(a, st);    // ws = 15 -> 225 -> 240
    case 256: return launch_fwd<256>(a, st);
    default: break;
  }
  set_error("attn_fwd(tc): window %dx%d not instantiated", ws, ws);
  return B200SWIN_EINVAL;
}

// ================================================================================================ backward
// Per (window, head), with S and dP recomputed on the tensor cores (never in HBM):
//   S  = Q K^T,  dP = dO V^T                 (tcgen05, fp32 in TMEM)
//   P  = exp2(S' - lse),  dS = P * (dP - D)  (one thread per row and column half; D = <dO, O>)
//   P, dS -> bf16 panels [128 query rows][64 keys] in smem (128 B swizzle).  The SAME bytes are a K-major
//   A operand (dQ = dS K) and an MN-major A operand (dV = P^T dO, dK = dS^T Q): no transpose is ever made.
//   dQ, dK (through the F.normalize backward) and dV go to the natural-layout dqkv tensor.
// The gradient of the 16*sigmoid bias table is a sum of dS over ALL windows: each CTA works on ONE head
// (persistent, head-major) and keeps its dS sums in REGISTERS across windows, flushing once at the end
// (tail rows of >128-token windows use shared-memory atomics).
}  // namespace b200swin
namespace b200swin {
namespace {
constexpr int kBwdThreads = 256;

struct TcBwdArgs {
  const __nv_bfloat16* qkv;
  const __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  const float* lse;
  const float* inv_norm;
  const float* table16;
  const float* scale;
  const float* qpad;
  const float* vpad;
  __nv_bfloat16* dqkv;
  float* dtable16;
  float* dscale;
  float* dvpad;
  WinGeom g;
  int C, nH;
  int64_t nwin;
};

template <int NPAD>
struct BwdLayout {
  static constexpr int MT = (NPAD + 127) / 128;
  // 64-key panels of P / dS.  The M=128 transposed MMAs always address two panels (keys 0..127), so at least two
  // are allocated (unused ones stay zero); windows with more than 128 tokens add a third for the M=64 tail.
  static constexpr int NP = NPAD > 128 ? 3 : 2;
  static constexpr uint32_t kRow = NPAD * 64;                        // one [NPAD][64 B] operand tile
  static constexpr uint32_t kBufBytes = 5 * kRow;                    // Q | dO | K | V | O
  static constexpr uint32_t kPanel = 128 * 128;                      // [128 rows][128 B]
  static constexpr uint32_t kPBytes = NP * kPanel;
  static constexpr uint32_t kTmemCols = 512;
  static constexpr uint32_t S_COL = 0, DP_COL = NPAD, DQ_COL = 2 * NPAD, DV_COL = DQ_COL + 32, DK_COL = DQ_COL + 64,
                            DVT_COL = DQ_COL + 96, DKT_COL = DQ_COL + 128;
  static_assert(2 * NPAD + 160 <= 512, "TMEM budget: S + dP + dQ + dV + dK (+ tails)");
  static constexpr int NH = NPAD / 2;                                // columns per thread (two warps share a lane quarter)
  static_assert(NH % 8 == 0, "column halves are processed in chunks of 8");
};

template <int NPAD>
__device__ __forceinline__ void load_item_bwd(const TcBwdArgs& a, int64_t win, int h, unsigned char* buf, int* tok,
                                              int N) {
  using LY = BwdLayout<NPAD>;
  const uint32_t q_s = ptx::smem_u32(buf), g_s = q_s + LY::kRow, k_s = g_s + LY::kRow, v_s = k_s + LY::kRow,
                 o_s = v_s + LY::kRow;
  const int C3 = 3 * a.C;
  for (int r = threadIdx.x; r < NPAD; r += kBwdThreads) {
    int t = -2;
    if (r < N) {
      int b, i, j, si, sj;
      bool real = win_token(a.g, win, r, b, i, j, si, sj);
      t = real ? ((b * a.g.H + i) * a.g.W + j) : -1;
    }
    tok[r] = t;
    if (t >= 0) {
      const __nv_bfloat16* src = a.qkv + (int64_t)t * C3 + h * HD;
      const __nv_bfloat16* gsrc = a.dout + (int64_t)t * a.C + h * HD;
      const __nv_bfloat16* osrc = a.out + (int64_t)t * a.C + h * HD;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t off = sw64_off(r, c);
        ptx::cp_async_16(q_s + off, src + c * 8);
        ptx::cp_async_16(k_s + off, src + a.C + c * 8);
        ptx::cp_async_16(v_s + off, src + 2 * a.C + c * 8);
        ptx::cp_async_16(g_s + off, gsrc + c * 8);
        ptx::cp_async_16(o_s + off, osrc + c * 8);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
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
        const uint32_t off = sw64_off(r, c);
        *reinterpret_cast<uint4*>(buf + off) = qv;
        *reinterpret_cast<uint4*>(buf + LY::kRow + off) = make_uint4(0, 0, 0, 0);          // dO of a cropped row
        *reinterpret_cast<uint4*>(buf + 2 * LY::kRow + off) = make_uint4(0, 0, 0, 0);      // k of a pad token
        *reinterpret_cast<uint4*>(buf + 3 * LY::kRow + off) = vv;
        *reinterpret_cast<uint4*>(buf + 4 * LY::kRow + off) = make_uint4(0, 0, 0, 0);
      }
    }
  }
}

// 32 bf16 of row r of a [rows][64 B] swizzled tile -> fp32
__device__ __forceinline__ void read_row_sw64(const unsigned char* tile, int r, float (&v)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint4 w = *reinterpret_cast<const uint4*>(tile + sw64_off(r, c));
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&ww[e]);
      float2 f = __bfloat1622float2(t);
      v[c * 8 + 2 * e] = f.x;
      v[c * 8 + 2 * e + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void store_row_bf16(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int c = 0; c < 4; ++c)
    d[c] = make_uint4(pack_bf16(v[c * 8], v[c * 8 + 1]), pack_bf16(v[c * 8 + 2], v[c * 8 + 3]),
                      pack_bf16(v[c * 8 + 4], v[c * 8 + 5]), pack_bf16(v[c * 8 + 6], v[c * 8 + 7]));
}

template <int NPAD>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ TcBwdArgs a) {
  using LY = BwdLayout<NPAD>;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar_a, bar_b;
  __shared__ uint32_t tmem_slot;
  __shared__ float dvpad_s[HD];
  __shared__ float red_s[kBwdThreads / 32];

  const WinGeom& g = a.g;
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1, ntab = tw * tw;
  const uint32_t base_u32 = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base_u32 - ptx::smem_u32(smem_dyn));
  unsigned char* bufs[2] = {sm, sm + LY::kBufBytes};
  unsigned char* Pp = sm + 2 * LY::kBufBytes;
  unsigned char* dSp = Pp + LY::kPBytes;
  float* tab = reinterpret_cast<float*>(dSp + LY::kPBytes);
  float* dtab = tab + ntab;
  float* D_s = dtab + ntab;                    // [NPAD] <dO_r, O_r>
  float* lse_s = D_s + NPAD;                   // [NPAD] lse in log2 units
  int* toks[2] = {reinterpret_cast<int*>(lse_s + NPAD), reinterpret_cast<int*>(lse_s + NPAD) + NPAD};
  int* meta = toks[1] + NPAD;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, half = warp >> 2;
  const int row_local = q * 32 + lane;
  const int h = blockIdx.x % a.nH;
  const int64_t wstep = gridDim.x / a.nH;

  for (int i = threadIdx.x; i < (int)(2 * LY::kBufBytes + 2 * LY::kPBytes) / 16; i += kBwdThreads)
    reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  for (int r = threadIdx.x; r < ntab; r += kBwdThreads) { tab[r] = a.table16[r * a.nH + h] * kLog2e; dtab[r] = 0.f; }
  if (threadIdx.x < HD) dvpad_s[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_a, 1);
    ptx::mbar_init(&bar_b, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_slot, LY::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);

  constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, NPAD, 0, 0);
  constexpr uint32_t idesc_dq = ptx::make_idesc_bf16(128, HD, 0, 1);     // A = dS (K-major), B = K MN-major
  constexpr uint32_t idesc_t = ptx::make_idesc_bf16(128, HD, 1, 1);      // A = P^T / dS^T (MN-major), B MN-major
  constexpr uint32_t idesc_t64 = ptx::make_idesc_bf16(64, HD, 1, 1);     // tail keys 128.. (M = 64)

  const float sc = a.scale[h], scale2 = sc * kLog2e;
  float acc[LY::NH];                       // register-resident sum over windows of dS[row_local, my columns]
#pragma unroll
  for (int c = 0; c < LY::NH; ++c) acc[c] = 0.f;
  float dsc = 0.f;
  uint32_t ph_a = 0, ph_b = 0;

  int64_t win = blockIdx.x / a.nH;
  int it = 0;
  if (win < a.nwin) load_item_bwd<NPAD>(a, win, h, bufs[0], toks[0], N);
  ptx::cp_async_commit();

  for (; win < a.nwin; win += wstep, ++it) {
    const int b = it & 1;
    if (win + wstep < a.nwin) load_item_bwd<NPAD>(a, win + wstep, h, bufs[b ^ 1], toks[b ^ 1], N);
    ptx::cp_async_commit();
    for (int r = threadIdx.x; r < NPAD; r += kBwdThreads) {
      int region = 0;
      if (g.shift > 0 && r < N) {
        int bb, i, j, si, sj;
        win_token(g, win, r, bb, i, j, si, sj);
        region = 3 * region_1d(si, g.Hp, ws, g.shift) + region_1d(sj, g.Wp, ws, g.shift);
      }
      meta[r] = (r < N ? ((r / ws) * tw + (r % ws)) : 0) | (region << 16);
    }
    ptx::cp_async_wait<1>();
    __syncthreads();                                   // everybody's gather of this item is visible
    unsigned char* buf = bufs[b];
    const int* tok = toks[b];
    // D_r = <dO_r, O_r>, lse in log2 units (pad / padding rows: dO = 0 -> D = 0; lse = +inf -> P = 0)
    for (int r = threadIdx.x; r < NPAD; r += kBwdThreads) {
      float gv[32], ov[32];
      read_row_sw64(buf + LY::kRow, r, gv);
      read_row_sw64(buf + 4 * LY::kRow, r, ov);
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) d = fmaf(gv[c], ov[c], d);
      D_s[r] = d;
      lse_s[r] = (r < N) ? a.lse[(win * a.nH + h) * N + r] * kLog2e : INFINITY;
    }
    ptx::fence_proxy_async_smem();
    __syncthreads();

    const uint32_t q_s = ptx::smem_u32(buf), g_s = q_s + LY::kRow, k_s = g_s + LY::kRow, v_s = k_s + LY::kRow;
    const uint32_t p_s = ptx::smem_u32(Pp), ds_s = ptx::smem_u32(dSp);

#pragma unroll 1
    for (int tile = 0; tile < LY::MT; ++tile) {
      const int rows_valid = min(128, N - tile * 128);
      if (rows_valid <= 0) break;
      // ---- S = Q_t K^T, dP = dO_t V^T
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t bk = ptx::make_smem_desc(k_s + ks * 32, 16, 512, kSw64);
          const uint64_t bv = ptx::make_smem_desc(v_s + ks * 32, 16, 512, kSw64);
          ptx::mma_bf16_ss(tmem_base + LY::S_COL, ptx::make_smem_desc(q_s + tile * 8192 + ks * 32, 16, 512, kSw64), bk,
                           idesc_s, ks);
          ptx::mma_bf16_ss(tmem_base + LY::DP_COL, ptx::make_smem_desc(g_s + tile * 8192 + ks * 32, 16, 512, kSw64), bv,
                           idesc_s, ks);
        }
        ptx::mma_commit(&bar_a);
      }
      ptx::mbar_wait(&bar_a, ph_a);
      ph_a ^= 1;
      ptx::tc_fence_after();

      // ---- P, dS for (row, my half of the columns)
      {
        const int r = tile * 128 + row_local;
        const bool rvalid = r < N;
        const int rr = rvalid ? r : 0;
        const int base_i = (meta[rr] & 0xffff) + (ws - 1) * (tw + 1);
        const int reg_i = meta[rr] >> 16;
        const float lse2 = rvalid ? lse_s[rr] : INFINITY;
        const float Dr = D_s[rr];
#pragma unroll
        for (int cc = 0; cc < LY::NH / 8; ++cc) {
          const int j0 = half * LY::NH + cc * 8;
          uint32_t sv[8], dv[8];
          ptx::tmem_ld_32x32b_x8(t_row + LY::S_COL + j0, sv);
          ptx::tmem_ld_32x32b_x8(t_row + LY::DP_COL + j0, dv);
          ptx::tmem_ld_wait();
          float p[8], ds[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int j = j0 + u;
            const int mj = meta[j];
            const int rel = base_i - (mj & 0xffff);
            const float cosv = __uint_as_float(sv[u]);
            float s2 = fmaf(cosv, scale2, tab[rel]);
            if ((mj >> 16) != reg_i) s2 += kMaskLog2;
            float pv = exp2f(s2 - lse2);
            float dsv = pv * (__uint_as_float(dv[u]) - Dr);
            if (!rvalid || (NPAD != N && j >= N)) { pv = 0.f; dsv = 0.f; }
            p[u] = pv;
            ds[u] = dsv;
            dsc = fmaf(dsv, cosv, dsc);
            if (tile == 0) acc[cc * 8 + u] += dsv;
            else if (dsv != 0.f) atomicAdd(dtab + rel, dsv);
          }
          const uint32_t off = (uint32_t)((j0 >> 6) * LY::kPanel + row_local * 128 + ((((j0 & 63) >> 3) ^ (row_local & 7)) << 4));
          *reinterpret_cast<uint4*>(Pp + off) =
              make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
          *reinterpret_cast<uint4*>(dSp + off) =
              make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7]));
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncthreads();

      // ---- dQ_t = dS K ;  dV += P^T dO_t ;  dK += dS^T Q_t
      if (threadIdx.x == 0) {
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ks = 0; ks < NPAD / 16; ++ks) {
          const uint64_t ad = ptx::make_smem_desc(ds_s + (ks >> 2) * LY::kPanel + (ks & 3) * 32, 16, 1024, 2);
          const uint64_t bd = ptx::make_smem_desc(k_s + ks * 1024, 512, 512, kSw64);
          ptx::mma_bf16_ss(tmem_base + LY::DQ_COL, ad, bd, idesc_dq, ks);
        }
        const int ksteps = (rows_valid + 15) / 16;
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t accf = (tile | ks) != 0 ? 1u : 0u;
          const uint64_t bg = ptx::make_smem_desc(g_s + (tile * 128 + ks * 16) * 64, 512, 512, kSw64);
          const uint64_t bq = ptx::make_smem_desc(q_s + (tile * 128 + ks * 16) * 64, 512, 512, kSw64);
          ptx::mma_bf16_ss(tmem_base + LY::DV_COL, ptx::make_smem_desc(p_s + ks * 2048, LY::kPanel, 1024, 2), bg, idesc_t, accf);
          ptx::mma_bf16_ss(tmem_base + LY::DK_COL, ptx::make_smem_desc(ds_s + ks * 2048, LY::kPanel, 1024, 2), bq, idesc_t, accf);
          if (NPAD > 128) {
            ptx::mma_bf16_ss(tmem_base + LY::DVT_COL, ptx::make_smem_desc(p_s + 2 * LY::kPanel + ks * 2048, LY::kPanel, 1024, 2),
                             bg, idesc_t64, accf);
            ptx::mma_bf16_ss(tmem_base + LY::DKT_COL, ptx::make_smem_desc(ds_s + 2 * LY::kPanel + ks * 2048, LY::kPanel, 1024, 2),
                             bq, idesc_t64, accf);
          }
        }
        ptx::mma_commit(&bar_b);
      }
      ptx::mbar_wait(&bar_b, ph_b);
      ph_b ^= 1;
      ptx::tc_fence_after();
      if (half == 0) {
        // dq = (scale * dS K  -  q_hat <.,q_hat>) / max(|q|, eps)
        const int r = tile * 128 + row_local;
        uint32_t o[32];
        ptx::tmem_ld_32x32b_x32(t_row + LY::DQ_COL, o);
        ptx::tmem_ld_wait();
        if (r < N && tok[r] >= 0) {
          const int t = tok[r];
          float qh[32], dq[32];
          read_row_sw64(buf, r, qh);
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) { dq[c] = __uint_as_float(o[c]) * sc; dot = fmaf(dq[c], qh[c], dot); }
          const float invn = a.inv_norm[((int64_t)t * 2 + 0) * a.nH + h];
#pragma unroll
          for (int c = 0; c < 32; ++c) dq[c] = (dq[c] - qh[c] * dot) * invn;
          store_row_bf16(a.dqkv + (int64_t)t * 3 * a.C + h * HD, dq);
        }
      }
      ptx::tc_fence_before();
      __syncthreads();          // S/dP/dQ columns and the P/dS panels are free for the next tile
    }

    // ---- dK, dV rows: keys 0..127 from the M=128 accumulators (warps 0-3), keys 128.. from the M=64 ones (warp 4)
    {
      const bool main_rows = half == 0;
      const bool tail_rows = NPAD > 128 && warp == 4;
      if (main_rows || tail_rows) {
        const int j = main_rows ? row_local : 128 + lane;
        uint32_t ov[32], ok[32];
        ptx::tmem_ld_32x32b_x32(t_row + (main_rows ? LY::DV_COL : LY::DVT_COL), ov);
        ptx::tmem_ld_32x32b_x32(t_row + (main_rows ? LY::DK_COL : LY::DKT_COL), ok);
        ptx::tmem_ld_wait();
        const bool jvalid = j < N && (main_rows || lane < 16);
        const int t = jvalid ? tok[j] : -2;
        if (t >= 0) {
          float kh[32], dk[32], dvv[32];
          read_row_sw64(buf + 2 * LY::kRow, j, kh);
          float dot = 0.f;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            dk[c] = __uint_as_float(ok[c]) * sc;
            dot = fmaf(dk[c], kh[c], dot);
            dvv[c] = __uint_as_float(ov[c]);
          }
          const float invn = a.inv_norm[((int64_t)t * 2 + 1) * a.nH + h];
#pragma unroll
          for (int c = 0; c < 32; ++c) dk[c] = (dk[c] - kh[c] * dot) * invn;
          __nv_bfloat16* dst = a.dqkv + (int64_t)t * 3 * a.C + h * HD;
          store_row_bf16(dst + a.C, dk);
          store_row_bf16(dst + 2 * a.C, dvv);
        }
        // pad tokens carry v = v_bias: their dV rows belong to v_bias (reduced over the warp, then smem)
        const bool is_pad = t == -1;
        if (__any_sync(0xffffffffu, is_pad)) {
#pragma unroll 1
          for (int c = 0; c < 32; ++c) {
            float v = is_pad ? __uint_as_float(ov[c]) : 0.f;
            v = warp_sum(v);
            if (lane == 0) atomicAdd(&dvpad_s[c], v);
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
  }
  ptx::cp_async_wait<0>();

  // ---- flush the per-CTA accumulators: register dS sums -> smem table -> global (one atomic per entry)
  {
    const int r = row_local;
    if (r < N) {
      const int base_i = ((r / ws) * tw + (r % ws)) + (ws - 1) * (tw + 1);
#pragma unroll
      for (int c = 0; c < LY::NH; ++c) {
        const int j = half * LY::NH + c;
        if (j < N) atomicAdd(dtab + base_i - ((j / ws) * tw + (j % ws)), acc[c]);
      }
    }
  }
  dsc = warp_sum(dsc);
  if (lane == 0) red_s[warp] = dsc;
  __syncthreads();
  for (int r = threadIdx.x; r < ntab; r += kBwdThreads) {
    const float v = dtab[r];
    if (v != 0.f) atomicAdd(a.dtable16 + r * a.nH + h, v);
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kBwdThreads / 32; ++w) s += red_s[w];
    atomicAdd(a.dscale + h, s);
  }
  if (threadIdx.x < HD && a.dvpad) {
    const float v = dvpad_s[threadIdx.x];
    if (v != 0.f) atomicAdd(a.dvpad + h * HD + threadIdx.x, v);
  }
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, LY::kTmemCols);
  }
}

template <int NPAD>
static int launch_bwd(const TcBwdArgs& a, cudaStream_t st) {
  using LY = BwdLayout<NPAD>;
  const int ws = a.g.ws, ntab = (2 * ws - 1) * (2 * ws - 1);
  size_t smem = 1024 + 2 * (size_t)LY::kBufBytes + 2 * (size_t)LY::kPBytes + 2 * (size_t)ntab * 4 + 5 * (size_t)NPAD * 4;
  BSW_REQUIRE(smem <= 227 * 1024, "attn_bwd(tc): shared memory budget exceeded (%zu B)", smem);
  BSW_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t groups = sm_count() / a.nH;                 // one CTA per SM, every CTA pinned to one head
  if (groups < 1) groups = 1;
  if (groups > a.nwin) groups = a.nwin;
  attn_bwd_tc_kernel<NPAD><<<(unsigned)(groups * a.nH), kBwdThreads, smem, st>>>(a);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
}  // namespace

bool attn_tc_bwd_supported(int ws, int C, int nH, const void* mask) {
  const int npad = (ws * ws + 15) / 16 * 16;
  return mask == nullptr && C == nH * HD && npad <= 176 && ws >= 2;
}

size_t attn_bwd_tc_workspace_bytes(int B, int H, int W, int nH, int ws) {
  return (attn_bwd_ws_supported(ws) && !use_legacy_bwd()) ? attn_bwd_ws_workspace_bytes(B, H, W, nH) : 0;
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, const float* inv_norm,
                const float* table16, const float* scale, const float* qpad, const float* vpad, const float* mask,
                int nWm, void* dqkv, float* dtable16, float* dscale, float* dvpad, void* workspace,
                size_t workspace_bytes, int B, int H, int W, int C, int nH, int ws, int shift, cudaStream_t st) {
  (void)nWm;
  BSW_REQUIRE(attn_tc_bwd_supported(ws, C, nH, mask),
              "attn_bwd(tc): needs head_dim 32, window <= 13x13 and the on-the-fly mask");
  BSW_REQUIRE(shift >= 0 && shift < ws, "attn_bwd(tc): bad shift");
  BSW_REQUIRE(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(dout) |
                reinterpret_cast<uintptr_t>(dqkv)) & 15) == 0 && C % 8 == 0,
              "attn_bwd(tc): tensors must be 16-byte aligned");
  BSW_REQUIRE((int64_t)B * H * W < (1ll << 29), "attn_bwd(tc): too many tokens");
  if (attn_bwd_ws_supported(ws) && !use_legacy_bwd()) {
    BSW_REQUIRE(workspace && workspace_bytes >= attn_bwd_ws_workspace_bytes(B, H, W, nH),
                "attn_bwd(tc): workspace too small (b200swin_attn_bwd_workspace_bytes)");
    return attn_bwd_ws(qkv, out, dout, lse, inv_norm, table16, scale, qpad, vpad, dqkv, dtable16, dscale, dvpad,
                       workspace, B, H, W, C, nH, ws, shift, st);
  }
  TcBwdArgs a;
  a.qkv = (const __nv_bfloat16*)qkv; a.out = (const __nv_bfloat16*)out; a.dout = (const __nv_bfloat16*)dout;
  a.lse = lse; a.inv_norm = inv_norm; a.table16 = table16; a.scale = scale; a.qpad = qpad; a.vpad = vpad;
  a.dqkv = (__nv_bfloat16*)dqkv; a.dtable16 = dtable16; a.dscale = dscale; a.dvpad = dvpad;
  make_geom(&a.g, B, H, W, ws, shift);
  a.C = C; a.nH = nH;
  a.nwin = (int64_t)B * a.g.nWh * a.g.nWw;
  const int npad = (ws * ws + 15) / 16 * 16;
  switch (npad) {
    case 16: return launch_bwd<16>(a, st);
    case 32: return launch_bwd<32>(a, st);
    case 48: return launch_bwd<48>(a, st);
    case 64: return launch_bwd<64>(a, st);
    case 112: return launch_bwd<112>(a, st);
    case 128: return launch_bwd<128>(a, st);
    case 144: return launch_bwd<144>(a, st);
    case 176: return launch_bwd<176>(a, st);
    default: break;
  }
  set_error("attn_bwd(tc): window %dx%d not instantiated", ws, ws);
  return B200SWIN_EINVAL;
}

}  // namespace b200swin
