// tcgen05 GEMM for the dense contractions of the Swin-V2 block (qkv, proj, fc1, fc2; forward, dgrad, wgrad):
//     D[M,N] = epilogue( A[M,K] . B[N,K]^T )      bf16 operands, fp32 accumulation in TMEM.
// Replaces F.linear / nn.Linear at models/swin_transformer_v2.py:286 (qkv), :334 (proj), :77/:87 (Mlp) and
// their autograd.
//
// * PERSISTENT: one CTA per SM walks a static tile schedule (consecutive CTAs share the A tile -> L2 reuse);
// * operands are staged by TMA (cp.async.bulk.tensor, 128 B swizzle) into a 4/6-stage mbarrier ring that
//   runs ahead across tile boundaries (no pipeline drain between tiles);
// * one elected thread issues tcgen05.mma (M=128, N=128|256, K=16 per instruction) into one of TWO TMEM
//   accumulators; tcgen05.commit releases smem stages and signals the epilogue, which drains accumulator i
//   while the tensor pipe fills accumulator i^1;
// * 16 epilogue warps read the accumulator with tcgen05.ld (one row per thread), apply the fused epilogue in
//   registers, stage 32-row x 64-byte units in shared memory (64 B swizzle, conflict-free) and hand them to
//   TMA stores (cp.async.bulk.tensor ... bulk_group), so HBM sees full-sector row writes and the M/N edges are
//   clipped by the tensor map:
//     NONE  (+bias) | GELU (bias, exact-erf GELU, side output gelu'(pre-activation) for the backward) | QKV
//     (q_bias/0/v_bias, per-head L2-normalisation of q and k in fp32 + 1/|q|,1/|k| side output) | DGELU (multiply
//     by the saved gelu', read straight from global memory while the accumulator load is in flight);
// * either operand may be "MN-major" (stored [K][M] / [K][N]), which is how dgrad (B = W as stored) and wgrad
//   (A = dY, B = X as stored) run WITHOUT any transposed copy in HBM;
// * fp32-accurate mode: operands split as hi+lo bf16 pairs, 3 MMAs per k-step (hi.hi + hi.lo + lo.hi);
// * split-K (blockIdx.z) with fp32 partials + a fixed-order reduce, used by wgrad where K = #tokens.
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

namespace {
constexpr int BM = 128, BK = 64;
// warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..: epilogue.  SIXTEEN epilogue warps (four per TMEM lane quarter,
// i.e. four per SM sub-partition, each owning every fourth 32-column chunk of the tile): with one warp per
// sub-partition the epilogue ran at ~0.3 instructions per cycle (MUFU, tcgen05.ld and staging latencies fully
// exposed) and took 5x the MMA time of a K = 512 tile; four warps hide each other's latencies.
constexpr int kEpiWarps = 16, kEpiSub = kEpiWarps / 4;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kABytes = BM * BK * 2;
// epilogue staging: one unit of 32 rows x 64 B (SWIZZLE_64B) per epilogue warp
constexpr uint32_t kEpiUnitBytes = 32 * 64;
constexpr uint32_t kEpiSmemBytes = kEpiWarps * kEpiUnitBytes;       // 32 KB
// CG = CTAs per tile (cta_group): 1, or 2 = a CTA pair computing a 256 x BN tile with ONE tcgen05.mma.cta_group::2 per
// k-step -- each CTA stages its own 128 A rows and HALF of the B rows (BN / 2), so a k-block costs 32 KB per CTA instead
// of 48 KB: a third less L2 -> SM traffic per flop (the 128 x 256 mainloop was TMA-latency-bound) and six stages in flight.
template <int BN, int CG>
struct Tile {
  static constexpr uint32_t kBBytes = (BN / CG) * BK * 2;            // this CTA's share of the B tile
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kStageBytes > 32768 ? 4 : 6;        // 4 x 48 KB or 6 x 32 KB = 192 KB
  static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kEpiSmemBytes + 1024;
  static constexpr uint32_t kTmemCols = 2 * BN;                      // double-buffered accumulator
};
constexpr float kInvSqrt2 = 0.70710678118654752f;
constexpr float kInvSqrt2Pi = 0.39894228040143268f;
}  // namespace

struct GemmParams {
  CUtensorMap tmA[2], tmB[2];
  CUtensorMap tmOut, tmAux;                 // store maps: box {64 B of columns, 32 rows}, SWIZZLE_64B
  int nseg;
  int64_t M, N, K;
  int num_kb, kb_per_split, splits;
  int m_tiles, n_tiles;
  int64_t num_tiles;
  int epilogue, out_dtype;
  void* out;
  int64_t ldo;
  const float* bias;
  const float* bias2;
  const void* aux_in;
  void* aux_out;
  float* inv_norm;
  int nH, Cq;
  int partial;                              // split-K: tmOut covers the fp32 workspace [splits * m_pad, N]
  int64_t m_pad;
};

// Exact-erf GELU and its derivative from ONE exponential and ONE reciprocal (the libdevice erff + expf pair cost
// more issue slots than the MMAs of a 128x256 tile at K = 512):
//   erf(x) = sign(x) (1 - (a1 t + ... + a5 t^5) exp(-x^2)),  t = 1 / (1 + p |x|)      (Abramowitz-Stegun 7.1.26,
//   |error| <= 1.5e-7 -- fp32 rounding level), and exp(-x^2) with x = z / sqrt(2) is also the Gaussian of gelu'.
//   h = z Phi(z),   g = gelu'(z) = Phi(z) + z phi(z),   Phi = (1 + erf(z / sqrt 2)) / 2,  phi = exp(-z^2/2) / sqrt(2 pi)
__device__ __forceinline__ void gelu_pair(float z, float& h, float& g) {
  const float ax = fabsf(z) * kInvSqrt2;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -0.72134752044448170f));     // exp(-z^2 / 2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float half_erfc = 0.5f * poly * e;                      // (1 - erf|x|) / 2
  const float Phi = z >= 0.f ? 1.0f - half_erfc : half_erfc;
  h = z * Phi;
  g = fmaf(z * kInvSqrt2Pi, e, Phi);
}

// bf16 epilogue: two elements at a time on the packed fp32x2 instructions of sm_100 (FFMA2 / FMUL2: one issue slot per
// two results) and with ONE MUFU operation per element instead of two -- at K = 128 / 256 the two MUFU operations per
// output (16 per clock and SM) cost as much time as the HBM traffic of the whole GEMM.  The reciprocal of the
// Abramowitz-Stegun form is replaced by a polynomial for the scaled complementary error function:
//   (1 - erf(a / sqrt 2)) / 2 = F(a) exp(-a^2 / 2),   F(a) = erfcx(a / sqrt 2) / 2 ~ P8(2 a / 4.5 - 1) on [0, 4.5]
// (weighted minimax fit, |error of Phi| <= 1.7e-6, of gelu <= 4.7e-6, of gelu' <= 1.8e-6 over all z -- 1/1000 of a
// bf16 ulp at 1; beyond a = 4.5 the Gaussian factor is below 4e-5 and a clamped F is exact to 1e-6).
// Phi = 1/2 + sign(z) (1/2 - F e),  gelu = z Phi,  gelu' = Phi + z e / sqrt(2 pi).
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ void gelu_pair2(float2 z, float2& h, float2& g) {
  const float2 ac = make_float2(fminf(fabsf(z.x), 4.5f), fminf(fabsf(z.y), 4.5f));
  const float2 t = __ffma2_rn(ac, splat2(2.0f / 4.5f), splat2(-1.0f));
  const float2 w = __fmul2_rn(__fmul2_rn(z, z), splat2(-0.72134752044448170f));
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(w.x));                              // exp(-z^2 / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(w.y));
  float2 F = __ffma2_rn(splat2(1.882177644e-02f), t, splat2(2.564300909e-03f));
  F = __ffma2_rn(F, t, splat2(-2.922779805e-03f));
  F = __ffma2_rn(F, t, splat2(-3.627864966e-02f));
  F = __ffma2_rn(F, t, splat2(3.699574419e-02f));
  F = __ffma2_rn(F, t, splat2(-5.326356786e-02f));
  F = __ffma2_rn(F, t, splat2(8.664873954e-02f));
  F = __ffma2_rn(F, t, splat2(-1.198449042e-01f));
  F = __ffma2_rn(F, t, splat2(1.536320908e-01f));
  const float2 d = __ffma2_rn(__fmul2_rn(F, e), splat2(-1.0f), splat2(0.5f));
  const float2 sg = make_float2(copysignf(1.0f, z.x), copysignf(1.0f, z.y));
  const float2 Phi = __ffma2_rn(sg, d, splat2(0.5f));
  h = __fmul2_rn(z, Phi);
  g = __ffma2_rn(__fmul2_rn(z, splat2(kInvSqrt2Pi)), e, Phi);
}

// Per-warp staging unit for the TMA-store epilogue: 32 rows x 64 B with the 64-byte swizzle (16-byte piece j of
// row r sits at r*64 + ((j ^ ((r >> 1) & 3)) << 4)): eight consecutive rows hit eight distinct 16-byte bank groups,
// so the st.shared.v4 of a warp is conflict-free.  While the TMA engine reads the unit of one warp the other three
// warps of the sub-partition compute.
struct Stager {
  uint32_t base;       // shared address of this warp's unit
  int lane;

  __device__ __forceinline__ uint32_t acquire() {
    if (lane == 0) ptx::bulk_wait_read<0>();               // the store that last used the unit has read it
    __syncwarp();
    return base;
  }
  __device__ __forceinline__ void piece(uint32_t unit, int j, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t addr = unit + (uint32_t)lane * 64u + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
  }
  __device__ __forceinline__ void release(const CUtensorMap* tm, uint32_t unit, int c0, int c1) {
    ptx::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_2d(tm, unit, c0, c1);
      ptx::bulk_commit();
    }
  }
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// 32 consecutive columns of one row per thread -> global through the staging ring (col0 / row0: tile coordinates of
// the warp's 32 x 32 chunk; the tensor map clips everything beyond M and N)
template <typename OutT>
__device__ __forceinline__ void store_chunk(Stager& st, const CUtensorMap* tm, const float (&v)[32], int col0, int row0,
                                            int64_t N) {
  if constexpr (sizeof(OutT) == 2) {
    const uint32_t u = st.acquire();
#pragma unroll
    for (int j = 0; j < 4; ++j)
      st.piece(u, j, pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
               pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    st.release(tm, u, col0, row0);
  } else {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (col0 + 16 * h >= N) break;                     // warp-uniform
      const uint32_t u = st.acquire();
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st.piece(u, j, __float_as_uint(v[16 * h + 4 * j]), __float_as_uint(v[16 * h + 4 * j + 1]),
                 __float_as_uint(v[16 * h + 4 * j + 2]), __float_as_uint(v[16 * h + 4 * j + 3]));
      st.release(tm, u, col0 + 16 * h, row0);
    }
  }
}

template <typename T>
__device__ __forceinline__ void load_chunk32(const T* p, float (&v)[32], int nvalid) {
  if (nvalid >= 32) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) {
      float t[4];
      ld4(p + c, t);
      v[c] = t[0]; v[c + 1] = t[1]; v[c + 2] = t[2]; v[c + 3] = t[3];
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = (c < nvalid) ? Io<T>::ld(p + c) : 0.f;
  }
}

// v[c] += bias[c0 + c] for the columns that exist (N is a multiple of 4, so float4 granularity is exact)
__device__ __forceinline__ void add_bias32(float (&v)[32], const float* __restrict__ bias, int nvalid) {
#pragma unroll
  for (int c = 0; c < 32; c += 4) {
    if (c < nvalid) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
      v[c] += b.x; v[c + 1] += b.y; v[c + 2] += b.z; v[c + 3] += b.w;
    }
  }
}

// `valid_row`: this thread's row exists (row < M); rows beyond M still take part in the staging (the store clips).
template <int EPI, typename OutT>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, Stager& st, float (&v)[32], int64_t row, int row0,
                                               int col0, int nvalid, bool valid_row, const uint4 (&aux)[4]) {
  if constexpr (EPI == B200SWIN_EPI_NONE) {
    if (p.bias) add_bias32(v, p.bias + col0, nvalid);
    store_chunk<OutT>(st, &p.tmOut, v, col0, row0, p.N);
  } else if constexpr (EPI == B200SWIN_EPI_GELU) {
    if (p.bias) add_bias32(v, p.bias + col0, nvalid);
    float g[32];
    if constexpr (sizeof(OutT) == 2) {
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        float2 h2, g2;
        gelu_pair2(make_float2(v[c], v[c + 1]), h2, g2);
        v[c] = h2.x; v[c + 1] = h2.y; g[c] = g2.x; g[c + 1] = g2.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) gelu_pair(v[c], v[c], g[c]);
    }
    if (p.aux_out) store_chunk<OutT>(st, &p.tmAux, g, col0, row0, p.N);
    store_chunk<OutT>(st, &p.tmOut, v, col0, row0, p.N);
  } else if constexpr (EPI == B200SWIN_EPI_RELU) {
    if (p.bias) add_bias32(v, p.bias + col0, nvalid);
    if (p.aux_out) {
      float g[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) g[c] = v[c] > 0.f ? 1.f : 0.f;
      store_chunk<OutT>(st, &p.tmAux, g, col0, row0, p.N);
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = fmaxf(v[c], 0.f);
    store_chunk<OutT>(st, &p.tmOut, v, col0, row0, p.N);
  } else if constexpr (EPI == B200SWIN_EPI_DGELU || EPI == B200SWIN_EPI_ADD) {
    constexpr bool MUL = EPI == B200SWIN_EPI_DGELU;
    if constexpr (sizeof(OutT) == 2) {
      // this row's 64 bytes of gelu' were requested (4 x ld.global.v4) before the accumulator was read
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t ww[4] = {aux[j].x, aux[j].y, aux[j].z, aux[j].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
          v[8 * j + 2 * e] = MUL ? v[8 * j + 2 * e] * f.x : v[8 * j + 2 * e] + f.x;
          v[8 * j + 2 * e + 1] = MUL ? v[8 * j + 2 * e + 1] * f.y : v[8 * j + 2 * e + 1] + f.y;
        }
      }
    } else {
      float z[32];
      if (valid_row) {
        load_chunk32<OutT>(reinterpret_cast<const OutT*>(p.aux_in) + row * p.ldo + col0, z, nvalid);
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) z[c] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = MUL ? v[c] * z[c] : v[c] + z[c];
    }
    store_chunk<OutT>(st, &p.tmOut, v, col0, row0, p.N);
  } else {  // B200SWIN_EPI_QKV: one 32-column chunk == one head of q, k or v  (swin_transformer_v2.py:283-293)
    const int part = col0 / p.Cq;
    const int cin = col0 - part * p.Cq;
    const float* b = part == 0 ? p.bias : (part == 2 ? p.bias2 : nullptr);
    if (b) add_bias32(v, b + cin, 32);
    if (part < 2) {
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) ss = fmaf(v[c], v[c], ss);
      const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);          // F.normalize(eps=1e-12)
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] *= inv;
      if (p.inv_norm && valid_row) p.inv_norm[(row * 2 + part) * p.nH + (cin >> 5)] = inv;
    }
    store_chunk<OutT>(st, &p.tmOut, v, col0, row0, p.N);
  }
}

// EPI / OUT_BF16 are template parameters so that every launch carries only its own epilogue (registers, code size);
// split-K launches (p.partial) are EPI_NONE with fp32 stores.
template <bool A_MN, bool B_MN, int BN, int EPI, bool OUT_BF16, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  using TL = Tile<BN, CG>;
  constexpr int STAGES = TL::kStages;
  // CTA pair: rank 0 (the leader) issues the MMAs; both CTAs load, both drain their own 128 accumulator rows
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const uint32_t unit = CG == 2 ? blockIdx.x >> 1 : blockIdx.x;       // persistent work unit: a CTA or a CTA pair
  const uint32_t nunits = CG == 2 ? gridDim.x >> 1 : gridDim.x;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t acc_full[2];
  __shared__ __align__(8) uint64_t acc_empty[2];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;      // SWIZZLE_128B needs 1024 B alignment
  unsigned char* smem_al = smem_dyn + (smem_base - ptx::smem_u32(smem_dyn));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    // pair: the leader's full barrier counts one arrival per CTA (+ the bytes of both), its accumulator-free barrier
    // the epilogue warps of both CTAs
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], CG); ptx::mbar_init(&empty_bar[s], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], kEpiWarps * CG); }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.tmA[0]);
    ptx::prefetch_tmap(&p.tmB[0]);
    ptx::prefetch_tmap(&p.tmOut);
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      ptx::tmem_alloc_2sm(&tmem_slot, TL::kTmemCols);
      ptx::tmem_relinquish_2sm();
    } else {
      ptx::tmem_alloc(&tmem_slot, TL::kTmemCols);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all();          // the peer's barriers exist before anything signals them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // static persistent schedule: tile t -> (split z, m block, n block); n fastest so that CTAs running at the
  // same time share the streamed A tile in L2 (the weight-side B operand is small and stays resident)
  // 32-bit arithmetic (the host checks num_tiles < 2^31): every role decodes every tile, the epilogue warps of the
  // side-tensor epilogues three times per tile, and a 64-bit division is ~100 instructions
  // (pair: p.m_tiles counts 256-row super-tiles; this CTA takes the 128-row block `rank` of it)
  auto decode = [&](int64_t t, int& mb, int& nb, int& z) {
    const uint32_t tt = (uint32_t)t, nt = (uint32_t)p.n_tiles, mt = (uint32_t)p.m_tiles;
    const uint32_t r = tt / nt;
    nb = (int)(tt - r * nt);
    z = (int)(r / mt);
    mb = (int)((r - (uint32_t)z * mt) * CG + rank);
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      uint32_t it = 0;
      constexpr int BNH = BN / CG;                                    // B rows staged by this CTA
      for (int64_t t = unit; t < p.num_tiles; t += nunits) {
        int mb, nb, z;
        decode(t, mb, nb, z);
        const int m0 = mb * BM, n0 = nb * BN + (int)rank * BNH;
        const int kb0 = z * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int seg = 0; seg < p.nseg; ++seg) {
          const CUtensorMap* ta = &p.tmA[seg == 2 ? 1 : 0];           // segments: hi.hi, hi.lo, lo.hi
          const CUtensorMap* tb = &p.tmB[seg == 1 ? 1 : 0];
          for (int kb = kb0; kb < kb1; ++kb, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            ptx::mbar_wait(&empty_bar[s], ph ^ 1);
            unsigned char* sa = smem_al + (size_t)s * TL::kStageBytes;
            unsigned char* sb = sa + kABytes;
            const int k0 = kb * BK;
            if constexpr (CG == 2) {
              // every byte of the pair is counted on the LEADER's barrier (the only one the MMA thread waits on)
              const uint32_t lead_bar = ptx::mapa_shared(ptx::smem_u32(&full_bar[s]), 0);
              if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[s], TL::kStageBytes * 2);
              else ptx::mbar_arrive_cluster(lead_bar);
              if (A_MN) {
                ptx::tma_load_2d_2sm(sa, ta, lead_bar, m0, k0);
                ptx::tma_load_2d_2sm(sa + kABytes / 2, ta, lead_bar, m0 + 64, k0);
              } else {
                ptx::tma_load_2d_2sm(sa, ta, lead_bar, k0, m0);
              }
              if (B_MN) {
#pragma unroll
                for (int j = 0; j < BNH / 64; ++j) ptx::tma_load_2d_2sm(sb + j * 8192, tb, lead_bar, n0 + 64 * j, k0);
              } else {
                ptx::tma_load_2d_2sm(sb, tb, lead_bar, k0, n0);
              }
            } else {
              ptx::mbar_arrive_expect_tx(&full_bar[s], TL::kStageBytes);
              if (A_MN) {
                ptx::tma_load_2d(sa, ta, &full_bar[s], m0, k0);
                ptx::tma_load_2d(sa + kABytes / 2, ta, &full_bar[s], m0 + 64, k0);
              } else {
                ptx::tma_load_2d(sa, ta, &full_bar[s], k0, m0);
              }
              if (B_MN) {
#pragma unroll
                for (int j = 0; j < BN / 64; ++j) ptx::tma_load_2d(sb + j * 8192, tb, &full_bar[s], n0 + 64 * j, k0);
              } else {
                ptx::tma_load_2d(sb, tb, &full_bar[s], k0, n0);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------------------------------------------ MMA issuer (pair: the leader only)
      constexpr uint32_t idesc = ptx::make_idesc_bf16(BM * CG, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // descriptors differ between k-steps only in the start-address field (low 14 bits, 16 B units)
      const uint64_t adesc0 = A_MN ? ptx::make_smem_desc(0, kABytes / 2, 1024) : ptx::make_smem_desc(0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? ptx::make_smem_desc(0, 8192, 1024) : ptx::make_smem_desc(0, 16, 1024);
      constexpr uint32_t a_step = (A_MN ? 2048 : 32) >> 4, b_step = (B_MN ? 2048 : 32) >> 4;
      uint32_t it = 0, tcount = 0;
      for (int64_t t = unit; t < p.num_tiles; t += nunits, ++tcount) {
        int mb, nb, z;
        decode(t, mb, nb, z);
        const int kb0 = z * p.kb_per_split, kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        const int iters = (kb1 - kb0) * p.nseg;
        const uint32_t acc = tcount & 1;
        ptx::mbar_wait(&acc_empty[acc], ((tcount >> 1) & 1) ^ 1);     // epilogue(s) have drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int i = 0; i < iters; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)s * TL::kStageBytes;
          const uint64_t ad = adesc0 + (sa >> 4), bd = bdesc0 + ((sa + kABytes) >> 4);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if constexpr (CG == 2) ptx::mma_bf16_ss_2sm(d_tmem, ad + k * a_step, bd + k * b_step, idesc, (i | k) != 0 ? 1u : 0u);
            else ptx::mma_bf16_ss(d_tmem, ad + k * a_step, bd + k * b_step, idesc, (i | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if constexpr (CG == 2) ptx::mma_commit_2sm(&empty_bar[s], 3); else ptx::mma_commit(&empty_bar[s]);
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CG == 2) ptx::mma_commit_2sm(&acc_full[acc], 3); else ptx::mma_commit(&acc_full[acc]);
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps (TMEM -> smem -> TMA store)
    const int q = warp & 3;                      // TMEM lane quarter this warp may access
    const int sub = (warp - 2) >> 2;             // which of the quarter's four warps: chunks sub, sub + 4, ...
    Stager st;
    st.base = smem_base + (uint32_t)STAGES * TL::kStageBytes + (uint32_t)(warp - 2) * kEpiUnitBytes;
    st.lane = lane;
    constexpr bool aux_bf16 = (EPI == B200SWIN_EPI_DGELU || EPI == B200SWIN_EPI_ADD) && OUT_BF16;
    // bf16 MUL epilogue (dgrad of fc2 times gelu'): every thread reads its row's 64 bytes of gelu' straight from global
    // memory, ONE UNIT AHEAD (the tile schedule is static), so the HBM latency hides behind the previous unit
    uint4 aux_next[4];
    auto aux_fetch = [&](int64_t tt, int cc) {
#pragma unroll
      for (int j = 0; j < 4; ++j) aux_next[j] = make_uint4(0u, 0u, 0u, 0u);
      if (tt >= p.num_tiles) return;
      int mb, nb, z;
      decode(tt, mb, nb, z);
      const int64_t rr = (int64_t)mb * BM + q * 32 + lane;
      const int64_t cc0 = (int64_t)nb * BN + cc * 32;
      if (rr < p.M && cc0 < p.N) {
        const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux_in) + rr * p.ldo + cc0);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (cc0 + 8 * j < p.N) aux_next[j] = __ldg(ap + j);
      }
    };
    if constexpr (aux_bf16) aux_fetch(unit, sub);
    uint32_t tcount = 0;
    const uint32_t lead_empty0 = CG == 2 ? ptx::mapa_shared(ptx::smem_u32(&acc_empty[0]), 0) : 0u;
    const uint32_t lead_empty1 = CG == 2 ? ptx::mapa_shared(ptx::smem_u32(&acc_empty[1]), 0) : 0u;
    for (int64_t t = unit; t < p.num_tiles; t += nunits, ++tcount) {
      int mb, nb, z;
      decode(t, mb, nb, z);
      const int m0 = mb * BM + q * 32, n0 = nb * BN;
      const int64_t row = (int64_t)m0 + lane;
      const uint32_t acc = tcount & 1;
      ptx::mbar_wait(&acc_full[acc], (tcount >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = sub; c < BN / 32; c += kEpiSub) {
        const int col0 = n0 + c * 32;
        uint4 aux[4];
        if constexpr (aux_bf16) {
#pragma unroll
          for (int j = 0; j < 4; ++j) aux[j] = aux_next[j];
          if (c + kEpiSub < BN / 32) aux_fetch(t, c + kEpiSub); else aux_fetch(t + nunits, sub);
        }
        if (m0 < p.M && col0 < p.N) {              // warp-uniform: this 32 x 32 chunk exists
          const int nvalid = (int)min((int64_t)32, p.N - col0);
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + (uint32_t)(c * 32), r);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if constexpr (OUT_BF16) {
            epilogue_chunk<EPI, __nv_bfloat16>(p, st, v, row, m0, col0, nvalid, row < p.M, aux);
          } else {
            if (EPI == B200SWIN_EPI_NONE && p.partial)
              store_chunk<float>(st, &p.tmOut, v, col0, (int)((int64_t)z * p.m_pad) + m0, p.N);
            else
              epilogue_chunk<EPI, float>(p, st, v, row, m0, col0, nvalid, row < p.M, aux);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      // one arrival per epilogue warp frees the accumulator -- on the leader's barrier: it owns the MMA thread
      if (lane == 0) {
        if constexpr (CG == 2) ptx::mbar_arrive_cluster(acc ? lead_empty1 : lead_empty0);
        else ptx::mbar_arrive(&acc_empty[acc]);
      }
    }
    if (lane == 0) ptx::bulk_wait<0>();            // all stores of this warp have landed before the CTA retires
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all();  // neither CTA leaves while its peer may still signal it / read its B half
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CG == 2) ptx::tmem_dealloc_2sm(tmem_base, TL::kTmemCols);
    else ptx::tmem_dealloc(tmem_base, TL::kTmemCols);
  }
}

// fixed-order split-K reduction: out[m,n] = sum_z partial[z,m,n] (+ bias[n])
template <typename OutT>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t MN, int64_t N, int64_t split_stride,
                     const float* __restrict__ bias, OutT* __restrict__ out) {
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < MN;
       i += (int64_t)gridDim.x * blockDim.x * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int z = 0; z < splits; ++z) {
      float t[4];
      ld4(partial + (int64_t)z * split_stride + i, t);
      acc[0] += t[0]; acc[1] += t[1]; acc[2] += t[2]; acc[3] += t[3];
    }
    if (bias) {
      int64_t n = i % N;
      acc[0] += bias[n]; acc[1] += bias[n + 1]; acc[2] += bias[n + 2]; acc[3] += bias[n + 3];
    }
    st4(out + i, acc);
  }
}

// --------------------------------------------------------------------------------------- host side
EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_tiled();
  BSW_REQUIRE(enc, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  BSW_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA: base pointer must be 16-byte aligned");
  BSW_REQUIRE(outer_stride_bytes % 16 == 0, "TMA: row stride (%llu B) must be a multiple of 16",
              (unsigned long long)outer_stride_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BSW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return B200SWIN_OK;
}

// ws x ws x 32-column boxes (one head of one window) of a bf16 [B][H][W][ld] tensor, written to shared memory as rows of
// 64 bytes in window order with the 64-byte swizzle; elements beyond H / W are zero-filled
int make_tmap_window_bf16(CUtensorMap* m, const void* base, int B, int H, int W, int ld, int ws) {
  EncodeTiledFn enc = get_encode_tiled();
  BSW_REQUIRE(enc, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  BSW_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && ld % 8 == 0, "TMA: window tensor must be 16-byte aligned");
  cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)ws, (cuuint32_t)ws, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BSW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (window boxes) failed with CUresult %d", (int)r);
  return B200SWIN_OK;
}

int make_tmap_2d(CUtensorMap* m, const void* base, int dtype, uint64_t inner, uint64_t outer,
                 uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode_tiled();
  BSW_REQUIRE(enc, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  BSW_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA: base pointer must be 16-byte aligned");
  BSW_REQUIRE(outer_stride_bytes % 16 == 0, "TMA: row stride (%llu B) must be a multiple of 16",
              (unsigned long long)outer_stride_bytes);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                              : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                     : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(m, dtype == B200SWIN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BSW_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return B200SWIN_OK;
}

// store map of a row-major [rows][N] output: box = 64 bytes of columns x 32 rows, 64 B swizzle
static int store_map(CUtensorMap* m, const void* ptr, int dtype, int64_t rows, int64_t N) {
  const uint32_t es = dtype == B200SWIN_BF16 ? 2 : 4;
  return make_tmap_2d(m, ptr, dtype, (uint64_t)N, (uint64_t)rows, (uint64_t)N * es, 64 / es, 32, 64);
}

static int operand_map(CUtensorMap* m, const void* ptr, int mn_major, int64_t mn, int64_t k, int rows) {
  // K-major: stored [mn][k] -> box {BK, rows};  MN-major: stored [k][mn] -> box {64 mn, BK k-rows}
  if (mn_major) return make_tmap_2d_bf16(m, ptr, (uint64_t)mn, (uint64_t)k, (uint64_t)mn * 2, 64, BK);
  return make_tmap_2d_bf16(m, ptr, (uint64_t)k, (uint64_t)mn, (uint64_t)k * 2, BK, (uint32_t)rows);
}

}  // namespace b200swin

using namespace b200swin;

static int pick_bn(int64_t N) { return (N % 256 == 0 || N >= 1024) ? 256 : 128; }

extern "C" int b200swin_gemm_splits(int64_t M, int64_t N, int64_t K) {
  // Split-K factor for the persistent kernel: minimise  waves x (k-blocks per split + per-tile overhead)  where a
  // wave is one tile per SM -- i.e. prefer tile counts just BELOW a multiple of the SM count over ones just above
  // (160 tiles on 148 SMs cost two full waves).  Smaller factors win ties (fewer fp32 partials to reduce).
  // work units: 128-row tiles on single CTAs, or 256-row tiles on CTA pairs (see b200swin_gemm_bf16)
  const int64_t cg = (pick_bn(N) == 256 && M > BM) ? 2 : 1;
  const int64_t tiles = ((M + BM * cg - 1) / (BM * cg)) * ((N + pick_bn(N) - 1) / pick_bn(N));
  const int64_t num_kb = (K + BK - 1) / BK;
  const int64_t sms = sm_count() / cg;
  int64_t max_splits = num_kb / 4 > 0 ? num_kb / 4 : 1;      // keep >= 4 k-blocks per split
  if (max_splits > 512) max_splits = 512;
  int64_t best = 1, best_cost = -1;
  for (int64_t s = 1; s <= max_splits; ++s) {
    const int64_t kb = (num_kb + s - 1) / s;
    const int64_t eff = (num_kb + kb - 1) / kb;               // splits that actually get work
    const int64_t waves = (tiles * eff + sms - 1) / sms;
    const int64_t cost = waves * (kb + 8) + eff / 8;          // + a small price per partial
    if (best_cost < 0 || cost < best_cost) { best = s; best_cost = cost; }
  }
  return (int)best;
}

extern "C" size_t b200swin_gemm_workspace_bytes(int64_t M, int64_t N, int splits) {
  // fp32 partials [splits][M rounded up to the 128-row tile][N]
  const int64_t m_pad = (M + BM - 1) / BM * BM;
  return splits > 1 ? (size_t)splits * (size_t)m_pad * (size_t)N * sizeof(float) : 0;
}

extern "C" int b200swin_gemm_bf16(const void* a_hi, const void* a_lo, int a_mn_major, const void* b_hi,
                                  const void* b_lo, int b_mn_major, int64_t M, int64_t N, int64_t K, int epilogue,
                                  const float* bias, const float* bias2, const void* aux_in, void* aux_out,
                                  float* inv_norm, int nH, void* out, int out_dtype, int splits, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  BSW_REQUIRE(a_hi && b_hi && out, "gemm: null pointer");
  BSW_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: non-positive dimension");
  BSW_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: dimension exceeds 2^31");
  BSW_REQUIRE((a_lo == nullptr) == (b_lo == nullptr), "gemm: a_lo and b_lo must be given together");
  BSW_REQUIRE(out_dtype == B200SWIN_F32 || out_dtype == B200SWIN_BF16, "gemm: bad out dtype %d", out_dtype);
  BSW_REQUIRE(epilogue >= B200SWIN_EPI_NONE && epilogue <= B200SWIN_EPI_RELU, "gemm: bad epilogue %d", epilogue);
  BSW_REQUIRE(N % 8 == 0, "gemm: N must be a multiple of 8 (16-byte rows for the TMA stores)");
  BSW_REQUIRE((a_mn_major ? M : K) % 8 == 0 && (b_mn_major ? N : K) % 8 == 0,
              "gemm: contiguous operand dimension must be a multiple of 8 (16-byte TMA rows)");
  if (splits < 1) splits = 1;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  if (epilogue == B200SWIN_EPI_QKV) {
    BSW_REQUIRE(N % 96 == 0 && nH > 0 && (N / 3) == (int64_t)nH * 32, "gemm: QKV epilogue needs N = 3*nH*32");
    BSW_REQUIRE(splits == 1, "gemm: QKV epilogue cannot be split");
    p.Cq = (int)(N / 3);
  }
  if (epilogue == B200SWIN_EPI_DGELU || epilogue == B200SWIN_EPI_ADD)
    BSW_REQUIRE(aux_in, "gemm: DGELU / ADD epilogues need aux_in");
  if (splits > 1) {
    BSW_REQUIRE(epilogue == B200SWIN_EPI_NONE, "gemm: split-K supports only the plain epilogue");
    BSW_REQUIRE(workspace && workspace_bytes >= b200swin_gemm_workspace_bytes(M, N, splits),
                "gemm: split-K workspace too small");
  }
  // 128x256 tiles halve the A re-reads and the smem bandwidth per MMA; 128x128 when N does not fill them
  const int bn = pick_bn(N);
  // CTA pairs (256 x 256 tiles, cta_group::2) whenever the tile is 256 wide and there are at least two row blocks
  const int cg = (bn == 256 && M > BM) ? 2 : 1;
  int rc;
  if ((rc = operand_map(&p.tmA[0], a_hi, a_mn_major, M, K, BM))) return rc;
  if ((rc = operand_map(&p.tmB[0], b_hi, b_mn_major, N, K, bn / cg))) return rc;
  p.nseg = 1;
  if (a_lo) {
    if ((rc = operand_map(&p.tmA[1], a_lo, a_mn_major, M, K, BM))) return rc;
    if ((rc = operand_map(&p.tmB[1], b_lo, b_mn_major, N, K, bn / cg))) return rc;
    p.nseg = 3;
  }
  p.M = M; p.N = N; p.K = K;
  p.num_kb = (int)((K + BK - 1) / BK);
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.epilogue = epilogue; p.out_dtype = out_dtype;
  p.out = out; p.ldo = N;
  p.bias = bias; p.bias2 = bias2; p.aux_in = aux_in; p.aux_out = aux_out; p.inv_norm = inv_norm; p.nH = nH;
  p.partial = splits > 1 ? 1 : 0;
  p.m_pad = (M + BM - 1) / BM * BM;
  if (splits > 1) p.bias = nullptr;
  p.splits = splits;
  if (splits > 1) {
    if ((rc = store_map(&p.tmOut, workspace, B200SWIN_F32, (int64_t)splits * p.m_pad, N))) return rc;
  } else {
    if ((rc = store_map(&p.tmOut, out, out_dtype, M, N))) return rc;
    if (aux_out && (rc = store_map(&p.tmAux, aux_out, out_dtype, M, N))) return rc;
  }

  cudaStream_t st = (cudaStream_t)stream;
  p.m_tiles = (int)((M + BM * cg - 1) / (BM * cg));                    // row blocks per work unit: 128, or 256 for a pair
  p.n_tiles = (int)((N + bn - 1) / bn);
  p.num_tiles = (int64_t)p.m_tiles * p.n_tiles * splits;
  BSW_REQUIRE(p.num_tiles < (1ll << 31), "gemm: too many tiles");
  const int64_t max_units = sm_count() / cg;
  const unsigned grid = (unsigned)((p.num_tiles < max_units ? p.num_tiles : max_units) * cg);
  const bool out_bf16 = out_dtype == B200SWIN_BF16 && splits == 1;     // split-K partials are fp32
#define LAUNCH(AM, BMN, BNV, EPI, OB, CGV)                                                                         \
  do {                                                                                                             \
    BSW_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<AM, BMN, BNV, EPI, OB, CGV>,                                      \
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<BNV, CGV>::kSmemBytes));  \
    cudaLaunchConfig_t cfg = {};                                                                                   \
    cfg.gridDim = dim3(grid);                                                                                      \
    cfg.blockDim = dim3(kGemmThreads);                                                                             \
    cfg.dynamicSmemBytes = Tile<BNV, CGV>::kSmemBytes;                                                             \
    cfg.stream = st;                                                                                               \
    cudaLaunchAttribute attr[1];                                                                                   \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                              \
    attr[0].val.clusterDim.x = CGV;                                                                                \
    attr[0].val.clusterDim.y = 1;                                                                                  \
    attr[0].val.clusterDim.z = 1;                                                                                  \
    cfg.attrs = attr;                                                                                              \
    cfg.numAttrs = CGV == 2 ? 1 : 0;                                                                               \
    BSW_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<AM, BMN, BNV, EPI, OB, CGV>, p));                             \
  } while (0)
#define LAUNCH_OB(AM, BMN, BNV, EPI, CGV) do { if (out_bf16) LAUNCH(AM, BMN, BNV, EPI, true, CGV); else LAUNCH(AM, BMN, BNV, EPI, false, CGV); } while (0)
#define LAUNCH_BN(AM, BMN, EPI)                                                  \
  do {                                                                           \
    if (bn == 256 && cg == 2) LAUNCH_OB(AM, BMN, 256, EPI, 2);                   \
    else if (bn == 256) LAUNCH_OB(AM, BMN, 256, EPI, 1);                         \
    else LAUNCH_OB(AM, BMN, 128, EPI, 1);                                        \
  } while (0)
  // instantiated combinations: the plain epilogue for every operand layout; GELU and QKV for the forward layout
  // (both K-major); DGELU for the dgrad layout (B read MN-major)
  if (epilogue == B200SWIN_EPI_NONE) {
    if (a_mn_major && b_mn_major) LAUNCH_BN(true, true, B200SWIN_EPI_NONE);
    else if (a_mn_major) LAUNCH_BN(true, false, B200SWIN_EPI_NONE);
    else if (b_mn_major) LAUNCH_BN(false, true, B200SWIN_EPI_NONE);
    else LAUNCH_BN(false, false, B200SWIN_EPI_NONE);
  } else if (epilogue == B200SWIN_EPI_GELU) {
    BSW_REQUIRE(!a_mn_major && !b_mn_major, "gemm: the GELU epilogue is built for K-major operands only");
    LAUNCH_BN(false, false, B200SWIN_EPI_GELU);
  } else if (epilogue == B200SWIN_EPI_RELU) {
    BSW_REQUIRE(!a_mn_major && !b_mn_major, "gemm: the RELU epilogue is built for K-major operands only");
    LAUNCH_BN(false, false, B200SWIN_EPI_RELU);
  } else if (epilogue == B200SWIN_EPI_QKV) {
    BSW_REQUIRE(!a_mn_major && !b_mn_major, "gemm: the QKV epilogue is built for K-major operands only");
    LAUNCH_BN(false, false, B200SWIN_EPI_QKV);
  } else if (epilogue == B200SWIN_EPI_DGELU) {
    BSW_REQUIRE(!a_mn_major && b_mn_major, "gemm: the DGELU epilogue is built for the dgrad layout (B MN-major) only");
    LAUNCH_BN(false, true, B200SWIN_EPI_DGELU);
  } else {
    BSW_REQUIRE(!a_mn_major && b_mn_major, "gemm: the ADD epilogue is built for the dgrad layout (B MN-major) only");
    LAUNCH_BN(false, true, B200SWIN_EPI_ADD);
  }
#undef LAUNCH_BN
#undef LAUNCH_OB
#undef LAUNCH
  BSW_LAUNCH_CHECK();
  if (splits > 1) {
    int64_t MN = M * N;
    int64_t blocks = (MN / 4 + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    int g = (int)(blocks < cap ? blocks : cap);
    if (out_dtype == B200SWIN_F32)
      splitk_reduce_kernel<float><<<g, 256, 0, st>>>((const float*)workspace, splits, MN, N, p.m_pad * N, bias,
                                                     (float*)out);
    else
      splitk_reduce_kernel<__nv_bfloat16><<<g, 256, 0, st>>>((const float*)workspace, splits, MN, N, p.m_pad * N, bias,
                                                            (__nv_bfloat16*)out);
    BSW_LAUNCH_CHECK();
  }
  return B200SWIN_OK;
}
