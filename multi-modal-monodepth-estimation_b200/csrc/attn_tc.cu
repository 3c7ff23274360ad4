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

int attn_bwd_tc(const void*, const void*, const void*, const float*, const float*, const float*, const float*,
                const float*, const float*, const float*, int, void*, float*, float*, float*, int, int, int, int, int,
                int, int, cudaStream_t) {
  set_error("attn_bwd: tensor-core backward not available in this build");
  return B200SWIN_EINVAL;
}

}  // namespace b200swin
