// Global (un-windowed) multi-head attention core: softmax(scale * Q K^T) V over whole token sequences, forward and backward.
// Replaces the scaled-dot-product inside nn.MultiheadAttention as the reference uses it in Transformer_Encoder.forward
// (models/cnn_transformer.py:192-216: 1200 tokens of a 30x40 feature map, 8 heads of 64 channels, no bias, no mask) and,
// with head_dim 32, the windowless form of WindowAttention's attn_type='normal' branch (models/swin_transformer_v2.py:296-298).
//
// Layout: q / k / v / out and their gradients are rows of a [B, N, ld] matrix (ld = row stride in elements, so the three
// operands may live in one packed projection buffer); head h occupies columns h*HD .. h*HD+HD-1.  lse is [B, nH, Nq] fp32.
//
// Two kernel families:
//  * bf16 storage: register-resident warp-level MMA (mma.sync.m16n8k16, the same scheme as attn_mma.cu): a CTA owns 64
//    query (or key) rows, one warp per 16 rows; the other side streams through a cp.async double-buffered ring of
//    64-row swizzled tiles read with ldmatrix.  The accumulator layout of S is the A-operand layout of P.V, so P never
//    leaves the registers.  Backward = D prep + a keys-owner pass (dK, dV) + a queries-owner pass (dQ); both recompute
//    S and dP, which keeps every gradient a fixed-order sum (deterministic, no atomics).
//  * fp32 storage (reference precision): CUDA-core kernels, one thread per owner row.
#include <math.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/b200swin.h"

namespace b200swin {
namespace {

typedef __nv_bfloat16 bf16;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct GArgs {
  const void *q, *k, *v, *out, *dout;
  void *o, *dq, *dk, *dv;
  float* lse;              // [B, nH, Nq]   natural log-sum-exp of the scaled logits
  float* dvec;             // [B, nH, Nq]   D = <dO, O>
  int64_t ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int B, Nq, Nk, nH;
  float scale;
};

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

// byte offset of 16-byte chunk c of row r in a tile of HD bf16 per row; the XOR spreads the 8 rows of an ldmatrix
// 8x8 block over the 8 bank groups
template <int HD>
__device__ __forceinline__ uint32_t swz(int r, int c) {
  if (HD == 64) return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
  return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4));
}

// rows r0 .. r0+63 of a [n, ld] matrix (columns of one head) -> swizzled tile; rows beyond n are zero
template <int HD>
__device__ __forceinline__ void load_tile(unsigned char* tile, const bf16* base, int64_t ld, int r0, int n) {
  constexpr int CPR = HD / 8;
  const uint32_t tile_s = ptx::smem_u32(tile);
  for (int idx = threadIdx.x; idx < 64 * CPR; idx += 128) {
    const int r = idx / CPR, c = idx % CPR;
    const uint32_t off = swz<HD>(r, c);
    if (r0 + r < n) ptx::cp_async_16(tile_s + off, base + (int64_t)(r0 + r) * ld + c * 8);
    else *reinterpret_cast<uint4*>(tile + off) = make_uint4(0, 0, 0, 0);
  }
}

// A-operand fragments (16 rows x HD) of rows row0 .. row0+15 of a tile
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[HD / 16][4], uint32_t tile_s, int row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks)
    ldsm4(f[ks], tile_s + swz<HD>(row0 + (lane & 7) + ((lane >> 3) & 1) * 8, ks * 2 + (lane >> 4)));
}

// acc[jj] (jj = 0, 1: columns n0+8jj .. +7) += A(16 x HD) . T[n0 .. n0+15][HD]^T     (T rows are the n index, K-major)
template <int HD>
__device__ __forceinline__ void mma_rows_nt(float (&acc)[2][4], const uint32_t (&a)[HD / 16][4], uint32_t tile_s, int n0,
                                            int lane) {
#pragma unroll
  for (int ks = 0; ks < HD / 16; ++ks) {
    uint32_t b[4];
    ldsm4(b, tile_s + swz<HD>(n0 + (lane & 7) + (lane >> 4) * 8, ks * 2 + ((lane >> 3) & 1)));
    mma16816(acc[0], a[ks], b[0], b[1]);
    mma16816(acc[1], a[ks], b[2], b[3]);
  }
}

// acc (16 x HD) += A(16 x 16, k = tile rows k0 .. k0+15) . T[k0 .. k0+15][HD]        (T rows are the k index)
template <int HD>
__device__ __forceinline__ void mma_rows_nn(float (&acc)[HD / 8][4], const uint32_t (&a)[4], uint32_t tile_s, int k0,
                                            int lane) {
#pragma unroll
  for (int dp = 0; dp < HD / 16; ++dp) {
    uint32_t b[4];
    ldsm4t(b, tile_s + swz<HD>(k0 + (lane & 7) + ((lane >> 3) & 1) * 8, dp * 2 + (lane >> 4)));
    mma16816(acc[2 * dp], a, b[0], b[1]);
    mma16816(acc[2 * dp + 1], a, b[2], b[3]);
  }
}

// store the 16 x HD accumulator of a warp (rows row0+g, row0+g+8) scaled by f0 / f1, as bf16
template <int HD>
__device__ __forceinline__ void store_rows(bf16* base, int64_t ld, int row0, int n, const float (&acc)[HD / 8][4], float f0,
                                           float f1, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int ra = row0 + g, rb = row0 + g + 8;
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    if (ra < n) *reinterpret_cast<uint32_t*>(base + (int64_t)ra * ld + 8 * j + 2 * t) = pack2(acc[j][0] * f0, acc[j][1] * f0);
    if (rb < n) *reinterpret_cast<uint32_t*>(base + (int64_t)rb * ld + 8 * j + 2 * t) = pack2(acc[j][2] * f1, acc[j][3] * f1);
  }
}

// ------------------------------------------------------------------------------------------------ forward (bf16, warp MMA)
// Measured on B200 at the config-3 shape (16 frames x 8 heads x 1200 tokens, head_dim 64): 206 us = 229 TF/s.  A variant
// with two 16-row tiles per warp (every K / V fragment feeding two MMAs, half the ldmatrix traffic) was slower (222 us:
// 218 registers, two CTAs per SM), so shared-memory wavefronts are not what limits this kernel; occupancy is.
template <int HD>
__global__ void __launch_bounds__(128) gattn_fwd_mma_kernel(GArgs a) {
  constexpr int TILE = 64 * HD * 2;
  __shared__ __align__(128) unsigned char sm[5 * TILE];
  unsigned char* Qs = sm;
  unsigned char* Ks = sm + TILE;
  unsigned char* Vs = sm + 3 * TILE;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bf16* qb = (const bf16*)a.q + (int64_t)b * a.Nq * a.ldq + h * HD;
  const bf16* kb = (const bf16*)a.k + (int64_t)b * a.Nk * a.ldk + h * HD;
  const bf16* vb = (const bf16*)a.v + (int64_t)b * a.Nk * a.ldv + h * HD;
  const int nkb = (a.Nk + 63) / 64;
  const float sl2 = a.scale * kLog2e;

  load_tile<HD>(Qs, qb, a.ldq, q0, a.Nq);
  load_tile<HD>(Ks, kb, a.ldk, 0, a.Nk);
  load_tile<HD>(Vs, vb, a.ldv, 0, a.Nk);
  ptx::cp_async_commit();

  uint32_t qf[HD / 16][4];
  float o[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int ib = 0; ib < nkb; ++ib) {
    const int stg = ib & 1;
    if (ib + 1 < nkb) {
      load_tile<HD>(Ks + (stg ^ 1) * TILE, kb, a.ldk, (ib + 1) * 64, a.Nk);
      load_tile<HD>(Vs + (stg ^ 1) * TILE, vb, a.ldv, (ib + 1) * 64, a.Nk);
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait<1>();
    __syncthreads();
    if (ib == 0) load_a_frags<HD>(qf, ptx::smem_u32(Qs), warp * 16, lane);
    const uint32_t ks_s = ptx::smem_u32(Ks + stg * TILE), vs_s = ptx::smem_u32(Vs + stg * TILE);

    float s[8][4];
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_rows_nt<HD>(acc, qf, ks_s, np * 16, lane);
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[2 * np][e] = acc[0][e]; s[2 * np + 1][e] = acc[1][e]; }    // raw dot products
    }
    if ((ib + 1) * 64 > a.Nk) {            // last block: keys beyond the sequence
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int key = ib * 64 + 8 * j + 2 * t;
        if (key >= a.Nk) s[j][0] = s[j][2] = -INFINITY;
        if (key + 1 >= a.Nk) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    // running maxima in the log2 domain (the scale is positive, so the maximum of the raw products is the maximum);
    // finite: key ib*64 always exists
    const float mn0 = fmaxf(m0, quad_max(mx0) * sl2), mn1 = fmaxf(m1, quad_max(mx1) * sl2);
    const float al0 = ex2f(m0 - mn0), al1 = ex2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    l0 *= al0; l1 *= al1;
    if (__any_sync(0xffffffffu, al0 != 1.f || al1 != 1.f)) {     // after the first blocks the maxima rarely move
#pragma unroll
      for (int j = 0; j < HD / 8; ++j) { o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1; }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2f(fmaf(s[j][0], sl2, -mn0)); s[j][1] = ex2f(fmaf(s[j][1], sl2, -mn0));
      s[j][2] = ex2f(fmaf(s[j][2], sl2, -mn1)); s[j][3] = ex2f(fmaf(s[j][3], sl2, -mn1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t pf[4] = {pack2(s[2 * kk][0], s[2 * kk][1]), pack2(s[2 * kk][2], s[2 * kk][3]),
                              pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack2(s[2 * kk + 1][2], s[2 * kk + 1][3])};
      mma_rows_nn<HD>(o, pf, vs_s, kk * 16, lane);
    }
    __syncthreads();                       // this stage is refilled by the next iteration's prefetch
  }
  l0 = quad_sum(l0);
  l1 = quad_sum(l1);
  bf16* ob = (bf16*)a.o + (int64_t)b * a.Nq * a.ldo + h * HD;
  store_rows<HD>(ob, a.ldo, q0 + warp * 16, a.Nq, o, 1.f / l0, 1.f / l1, lane);
  if (t == 0) {
    float* lr = a.lse + ((int64_t)b * a.nH + h) * a.Nq;
    const int ra = q0 + warp * 16 + g, rb = ra + 8;
    if (ra < a.Nq) lr[ra] = (m0 + log2f(l0)) * kLn2;
    if (rb < a.Nq) lr[rb] = (m1 + log2f(l1)) * kLn2;
  }
}

// per-row scalars of rows r0 .. r0+63 -> smem (lse in the log2 domain; +inf / 0 beyond the sequence, which makes P = 0)
__device__ __forceinline__ void load_row_scalars(float* lse2s, float* ds, const float* lse, const float* dvec, int r0, int n) {
  const int r = threadIdx.x & 63;
  if (threadIdx.x < 64) lse2s[r] = (r0 + r < n) ? lse[r0 + r] * kLog2e : INFINITY;
  else ds[r] = (r0 + r < n) ? dvec[r0 + r] : 0.f;
}

// ------------------------------------------------------------------------------------- backward, keys-owner pass (dK, dV)
// warp = 16 keys (M rows).  Per block of 64 queries:  S^T = K Q^T, dP^T = V dO^T  ->  P^T, dS^T in registers, which are the
// A operands of  dV += P^T dO  and  dK += dS^T Q.
template <int HD>
__global__ void __launch_bounds__(128) gattn_dkv_mma_kernel(GArgs a) {
  constexpr int TILE = 64 * HD * 2;
  __shared__ __align__(128) unsigned char sm[4 * TILE];
  __shared__ float lse2s[2][64], dss[2][64];
  unsigned char* Qs = sm;                  // [2] stages
  unsigned char* Gs = sm + 2 * TILE;       // dO, [2] stages
  const int k0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, t = lane & 3;
  const bf16* qb = (const bf16*)a.q + (int64_t)b * a.Nq * a.ldq + h * HD;
  const bf16* kb = (const bf16*)a.k + (int64_t)b * a.Nk * a.ldk + h * HD;
  const bf16* vb = (const bf16*)a.v + (int64_t)b * a.Nk * a.ldv + h * HD;
  const bf16* gb = (const bf16*)a.dout + (int64_t)b * a.Nq * a.lddo + h * HD;
  const float* lse = a.lse + ((int64_t)b * a.nH + h) * a.Nq;
  const float* dvec = a.dvec + ((int64_t)b * a.nH + h) * a.Nq;
  const int nqb = (a.Nq + 63) / 64;
  const float sl2 = a.scale * kLog2e;

  // the CTA's own K / V tiles pass through the stage-1 buffers once; their fragments stay in registers
  load_tile<HD>(Qs + TILE, kb, a.ldk, k0, a.Nk);
  load_tile<HD>(Gs + TILE, vb, a.ldv, k0, a.Nk);
  load_tile<HD>(Qs, qb, a.ldq, 0, a.Nq);
  load_tile<HD>(Gs, gb, a.lddo, 0, a.Nq);
  load_row_scalars(lse2s[0], dss[0], lse, dvec, 0, a.Nq);
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();
  uint32_t kf[HD / 16][4], vf[HD / 16][4];
  load_a_frags<HD>(kf, ptx::smem_u32(Qs + TILE), warp * 16, lane);
  load_a_frags<HD>(vf, ptx::smem_u32(Gs + TILE), warp * 16, lane);
  __syncthreads();

  float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  for (int ib = 0; ib < nqb; ++ib) {
    const int stg = ib & 1;
    if (ib + 1 < nqb) {
      load_tile<HD>(Qs + (stg ^ 1) * TILE, qb, a.ldq, (ib + 1) * 64, a.Nq);
      load_tile<HD>(Gs + (stg ^ 1) * TILE, gb, a.lddo, (ib + 1) * 64, a.Nq);
      load_row_scalars(lse2s[stg ^ 1], dss[stg ^ 1], lse, dvec, (ib + 1) * 64, a.Nq);
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait<1>();
    __syncthreads();
    const uint32_t qs_s = ptx::smem_u32(Qs + stg * TILE), gs_s = ptx::smem_u32(Gs + stg * TILE);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_rows_nt<HD>(st, kf, qs_s, np * 16, lane);
      mma_rows_nt<HD>(dp, vf, gs_s, np * 16, lane);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int col = np * 16 + jj * 8 + 2 * t;                       // query column of c0 / c2; c1 / c3 are col + 1
        const float2 l2 = *reinterpret_cast<const float2*>(&lse2s[stg][col]);
        const float2 dd = *reinterpret_cast<const float2*>(&dss[stg][col]);
        const float p0 = ex2f(st[jj][0] * sl2 - l2.x), p1 = ex2f(st[jj][1] * sl2 - l2.y);
        const float p2 = ex2f(st[jj][2] * sl2 - l2.x), p3 = ex2f(st[jj][3] * sl2 - l2.y);
        st[jj][0] = p0; st[jj][1] = p1; st[jj][2] = p2; st[jj][3] = p3;
        dp[jj][0] = p0 * (dp[jj][0] - dd.x); dp[jj][1] = p1 * (dp[jj][1] - dd.y);
        dp[jj][2] = p2 * (dp[jj][2] - dd.x); dp[jj][3] = p3 * (dp[jj][3] - dd.y);
      }
      const uint32_t pa[4] = {pack2(st[0][0], st[0][1]), pack2(st[0][2], st[0][3]), pack2(st[1][0], st[1][1]),
                              pack2(st[1][2], st[1][3])};
      const uint32_t da[4] = {pack2(dp[0][0], dp[0][1]), pack2(dp[0][2], dp[0][3]), pack2(dp[1][0], dp[1][1]),
                              pack2(dp[1][2], dp[1][3])};
      mma_rows_nn<HD>(dv, pa, gs_s, np * 16, lane);
      mma_rows_nn<HD>(dk, da, qs_s, np * 16, lane);
    }
    __syncthreads();
  }
  bf16* dkb = (bf16*)a.dk + (int64_t)b * a.Nk * a.lddk + h * HD;
  bf16* dvb = (bf16*)a.dv + (int64_t)b * a.Nk * a.lddv + h * HD;
  store_rows<HD>(dkb, a.lddk, k0 + warp * 16, a.Nk, dk, a.scale, a.scale, lane);
  store_rows<HD>(dvb, a.lddv, k0 + warp * 16, a.Nk, dv, 1.f, 1.f, lane);
}

// --------------------------------------------------------------------------------------- backward, queries-owner pass (dQ)
template <int HD>
__global__ void __launch_bounds__(128) gattn_dq_mma_kernel(GArgs a) {
  constexpr int TILE = 64 * HD * 2;
  __shared__ __align__(128) unsigned char sm[4 * TILE];
  unsigned char* Ks = sm;
  unsigned char* Vs = sm + 2 * TILE;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const bf16* qb = (const bf16*)a.q + (int64_t)b * a.Nq * a.ldq + h * HD;
  const bf16* kb = (const bf16*)a.k + (int64_t)b * a.Nk * a.ldk + h * HD;
  const bf16* vb = (const bf16*)a.v + (int64_t)b * a.Nk * a.ldv + h * HD;
  const bf16* gb = (const bf16*)a.dout + (int64_t)b * a.Nq * a.lddo + h * HD;
  const float* lse = a.lse + ((int64_t)b * a.nH + h) * a.Nq;
  const float* dvec = a.dvec + ((int64_t)b * a.nH + h) * a.Nq;
  const int nkb = (a.Nk + 63) / 64;
  const float sl2 = a.scale * kLog2e;

  load_tile<HD>(Ks + TILE, qb, a.ldq, q0, a.Nq);
  load_tile<HD>(Vs + TILE, gb, a.lddo, q0, a.Nq);
  load_tile<HD>(Ks, kb, a.ldk, 0, a.Nk);
  load_tile<HD>(Vs, vb, a.ldv, 0, a.Nk);
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[HD / 16][4], gf[HD / 16][4];
  load_a_frags<HD>(qf, ptx::smem_u32(Ks + TILE), warp * 16, lane);
  load_a_frags<HD>(gf, ptx::smem_u32(Vs + TILE), warp * 16, lane);
  __syncthreads();
  const int ra = q0 + warp * 16 + g, rb = ra + 8;
  const float la = ra < a.Nq ? lse[ra] * kLog2e : INFINITY, lb = rb < a.Nq ? lse[rb] * kLog2e : INFINITY;
  const float da_ = ra < a.Nq ? dvec[ra] : 0.f, db_ = rb < a.Nq ? dvec[rb] : 0.f;

  float dq[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
  for (int ib = 0; ib < nkb; ++ib) {
    const int stg = ib & 1;
    if (ib + 1 < nkb) {
      load_tile<HD>(Ks + (stg ^ 1) * TILE, kb, a.ldk, (ib + 1) * 64, a.Nk);
      load_tile<HD>(Vs + (stg ^ 1) * TILE, vb, a.ldv, (ib + 1) * 64, a.Nk);
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait<1>();
    __syncthreads();
    const uint32_t ks_s = ptx::smem_u32(Ks + stg * TILE), vs_s = ptx::smem_u32(Vs + stg * TILE);
    const bool tail = (ib + 1) * 64 > a.Nk;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      mma_rows_nt<HD>(s, qf, ks_s, np * 16, lane);
      mma_rows_nt<HD>(dp, gf, vs_s, np * 16, lane);
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        float p0 = ex2f(s[jj][0] * sl2 - la), p1 = ex2f(s[jj][1] * sl2 - la);
        float p2 = ex2f(s[jj][2] * sl2 - lb), p3 = ex2f(s[jj][3] * sl2 - lb);
        if (tail) {
          const int key = ib * 64 + np * 16 + jj * 8 + 2 * t;
          if (key >= a.Nk) p0 = p2 = 0.f;
          if (key + 1 >= a.Nk) p1 = p3 = 0.f;
        }
        dp[jj][0] = p0 * (dp[jj][0] - da_); dp[jj][1] = p1 * (dp[jj][1] - da_);
        dp[jj][2] = p2 * (dp[jj][2] - db_); dp[jj][3] = p3 * (dp[jj][3] - db_);
      }
      const uint32_t dsa[4] = {pack2(dp[0][0], dp[0][1]), pack2(dp[0][2], dp[0][3]), pack2(dp[1][0], dp[1][1]),
                               pack2(dp[1][2], dp[1][3])};
      mma_rows_nn<HD>(dq, dsa, ks_s, np * 16, lane);
    }
    __syncthreads();
  }
  bf16* dqb = (bf16*)a.dq + (int64_t)b * a.Nq * a.lddq + h * HD;
  store_rows<HD>(dqb, a.lddq, q0 + warp * 16, a.Nq, dq, a.scale, a.scale, lane);
}

// ------------------------------------------------------------------------------------------------ D = <dO, O> per (row, head)
template <typename T, int HD>
__global__ void gattn_prep_kernel(GArgs a) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // (b, row, h)
  const int64_t total = (int64_t)a.B * a.Nq * a.nH;
  if (idx >= total) return;
  const int h = (int)(idx % a.nH);
  const int64_t br = idx / a.nH;
  const int b = (int)(br / a.Nq), r = (int)(br % a.Nq);
  const T* o = (const T*)a.out + br * a.ldo + h * HD;
  const T* g = (const T*)a.dout + br * a.lddo + h * HD;
  float acc = 0.f;
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    float x[4], y[4];
    ld4(o + c, x);
    ld4(g + c, y);
    acc = fmaf(x[0], y[0], acc); acc = fmaf(x[1], y[1], acc); acc = fmaf(x[2], y[2], acc); acc = fmaf(x[3], y[3], acc);
  }
  a.dvec[((int64_t)b * a.nH + h) * a.Nq + r] = acc;
}

// ------------------------------------------------------------------------------------------- CUDA-core kernels (any dtype)
constexpr int KC = 32;                     // streamed rows per shared-memory chunk

template <typename T, int HD>
__device__ __forceinline__ void load_row(const T* p, float (&v)[HD]) {
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    float x[4];
    ld4(p + c, x);
    v[c] = x[0]; v[c + 1] = x[1]; v[c + 2] = x[2]; v[c + 3] = x[3];
  }
}
template <typename T, int HD>
__device__ __forceinline__ void store_row(T* p, const float (&v)[HD], float f) {
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float x[4] = {v[c] * f, v[c + 1] * f, v[c + 2] * f, v[c + 3] * f};
    st4(p + c, x);
  }
}
// rows r0 .. r0+kc-1 of one head of a [n, ld] matrix -> fp32 chunk [kc][HD] in shared memory
template <typename T, int HD>
__device__ __forceinline__ void load_chunk(float* dst, const T* base, int64_t ld, int r0, int kc) {
  for (int idx = threadIdx.x; idx < kc * (HD / 4); idx += blockDim.x) {
    const int r = idx / (HD / 4), c = (idx % (HD / 4)) * 4;
    float x[4];
    ld4(base + (int64_t)(r0 + r) * ld + c, x);
    *reinterpret_cast<float4*>(dst + r * HD + c) = make_float4(x[0], x[1], x[2], x[3]);
  }
}
template <int HD>
__device__ __forceinline__ float dot_row(const float (&a)[HD], const float* b) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(b + c);
    s0 = fmaf(a[c], t.x, s0); s1 = fmaf(a[c + 1], t.y, s1);
    s0 = fmaf(a[c + 2], t.z, s0); s1 = fmaf(a[c + 3], t.w, s1);
  }
  return s0 + s1;
}
template <int HD>
__device__ __forceinline__ void axpy_row(float (&acc)[HD], float w, const float* b) {
#pragma unroll
  for (int c = 0; c < HD; c += 4) {
    const float4 t = *reinterpret_cast<const float4*>(b + c);
    acc[c] = fmaf(w, t.x, acc[c]); acc[c + 1] = fmaf(w, t.y, acc[c + 1]);
    acc[c + 2] = fmaf(w, t.z, acc[c + 2]); acc[c + 3] = fmaf(w, t.w, acc[c + 3]);
  }
}

// thread = query row; K / V stream through shared memory in chunks of KC rows; online softmax in the log2 domain
template <typename T, int HD>
__global__ void __launch_bounds__(128) gattn_fwd_simt_kernel(GArgs a) {
  __shared__ __align__(16) float ks[KC * HD];
  __shared__ __align__(16) float vs[KC * HD];
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < a.Nq;
  const T* kb = (const T*)a.k + (int64_t)b * a.Nk * a.ldk + h * HD;
  const T* vb = (const T*)a.v + (int64_t)b * a.Nk * a.ldv + h * HD;
  const float sl2 = a.scale * kLog2e;
  float q[HD], o[HD];
#pragma unroll
  for (int c = 0; c < HD; ++c) { q[c] = 0.f; o[c] = 0.f; }
  if (active) load_row<T, HD>((const T*)a.q + ((int64_t)b * a.Nq + i) * a.ldq + h * HD, q);
  float m = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < a.Nk; j0 += KC) {
    const int kc = min(KC, a.Nk - j0);
    if (j0 > 0) __syncthreads();
    load_chunk<T, HD>(ks, kb, a.ldk, j0, kc);
    load_chunk<T, HD>(vs, vb, a.ldv, j0, kc);
    __syncthreads();
    for (int jj = 0; jj < kc; jj += 8) {
      float s[8];
      float mx = -INFINITY;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s[u] = (jj + u < kc) ? dot_row<HD>(q, ks + (jj + u) * HD) * sl2 : -INFINITY;
        mx = fmaxf(mx, s[u]);
      }
      const float mn = fmaxf(m, mx);
      const float corr = exp2f(m - mn);
      m = mn;
      l *= corr;
#pragma unroll
      for (int c = 0; c < HD; ++c) o[c] *= corr;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (jj + u < kc) {
          const float p = exp2f(s[u] - mn);
          l += p;
          axpy_row<HD>(o, p, vs + (jj + u) * HD);
        }
      }
    }
  }
  if (active) {
    store_row<T, HD>((T*)a.o + ((int64_t)b * a.Nq + i) * a.ldo + h * HD, o, 1.f / l);
    a.lse[((int64_t)b * a.nH + h) * a.Nq + i] = (m + log2f(l)) * kLn2;
  }
}

// Backward with one thread per OWNER row and the other side streamed through shared memory.
//   MODE 0 (dQ): owner = query i (x = q_i, y = dO_i; lse_i, D_i scalars), streams K, V:   dq_i = scale * sum_j dS_ij k_j
//   MODE 1 (dK): owner = key j   (x = k_j, y = v_j),  streams Q, dO (+ lse, D per row):    dk_j = scale * sum_i dS_ij q_i
//   MODE 2 (dV): owner = key j   (x = k_j),           streams Q, dO (+ lse):               dv_j = sum_i P_ij dO_i
// with P = exp(scale * x.X_r - lse), dS = P * (y.Y_r - D).
template <typename T, int HD, int MODE>
__global__ void __launch_bounds__(128) gattn_bwd_simt_kernel(GArgs a) {
  __shared__ __align__(16) float xs[KC * HD];
  __shared__ __align__(16) float ys[KC * HD];
  __shared__ float lses[KC], dss[KC];
  const int h = blockIdx.y, b = blockIdx.z;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_own = MODE == 0 ? a.Nq : a.Nk, n_str = MODE == 0 ? a.Nk : a.Nq;
  const bool active = i < n_own;
  const float sl2 = a.scale * kLog2e;
  const T* xo = MODE == 0 ? (const T*)a.q + ((int64_t)b * a.Nq + i) * a.ldq : (const T*)a.k + ((int64_t)b * a.Nk + i) * a.ldk;
  const T* yo = MODE == 0 ? (const T*)a.dout + ((int64_t)b * a.Nq + i) * a.lddo : (const T*)a.v + ((int64_t)b * a.Nk + i) * a.ldv;
  const T* xsb = MODE == 0 ? (const T*)a.k + (int64_t)b * a.Nk * a.ldk : (const T*)a.q + (int64_t)b * a.Nq * a.ldq;
  const T* ysb = MODE == 0 ? (const T*)a.v + (int64_t)b * a.Nk * a.ldv : (const T*)a.dout + (int64_t)b * a.Nq * a.lddo;
  const int64_t ldxs = MODE == 0 ? a.ldk : a.ldq, ldys = MODE == 0 ? a.ldv : a.lddo;
  const float* lse = a.lse + ((int64_t)b * a.nH + h) * a.Nq;
  const float* dvec = a.dvec + ((int64_t)b * a.nH + h) * a.Nq;
  float x[HD], y[HD], acc[HD];                    // y is dead code in MODE 2
#pragma unroll
  for (int c = 0; c < HD; ++c) { x[c] = 0.f; y[c] = 0.f; acc[c] = 0.f; }
  float lse_own = INFINITY, d_own = 0.f;
  if (active) {
    load_row<T, HD>(xo + h * HD, x);
    if (MODE != 2) load_row<T, HD>(yo + h * HD, y);
    if (MODE == 0) { lse_own = lse[i] * kLog2e; d_own = dvec[i]; }
  }
  for (int j0 = 0; j0 < n_str; j0 += KC) {
    const int kc = min(KC, n_str - j0);
    if (j0 > 0) __syncthreads();
    load_chunk<T, HD>(xs, xsb + h * HD, ldxs, j0, kc);
    load_chunk<T, HD>(ys, ysb + h * HD, ldys, j0, kc);
    if (MODE != 0 && (int)threadIdx.x < kc) {
      lses[threadIdx.x] = lse[j0 + threadIdx.x] * kLog2e;
      dss[threadIdx.x] = dvec[j0 + threadIdx.x];
    }
    __syncthreads();
    if (active) {
      for (int r = 0; r < kc; ++r) {
        const float l2 = MODE == 0 ? lse_own : lses[r];
        const float p = exp2f(dot_row<HD>(x, xs + r * HD) * sl2 - l2);
        if (MODE == 2) {
          axpy_row<HD>(acc, p, ys + r * HD);
        } else {
          const float dd = MODE == 0 ? d_own : dss[r];
          const float ds = p * (dot_row<HD>(y, ys + r * HD) - dd);
          axpy_row<HD>(acc, ds, xs + r * HD);
        }
      }
    }
  }
  if (active) {
    if (MODE == 0) store_row<T, HD>((T*)a.dq + ((int64_t)b * a.Nq + i) * a.lddq + h * HD, acc, a.scale);
    if (MODE == 1) store_row<T, HD>((T*)a.dk + ((int64_t)b * a.Nk + i) * a.lddk + h * HD, acc, a.scale);
    if (MODE == 2) store_row<T, HD>((T*)a.dv + ((int64_t)b * a.Nk + i) * a.lddv + h * HD, acc, 1.f);
  }
}

// attention weights averaged over the heads (what nn.MultiheadAttention returns with need_weights=True):
//   w[b, i, j] = 1/nH * sum_h exp(scale * q_ih . k_jh - lse[b, h, i]);   one thread per (i, j) of a 16 x 16 tile
template <typename T, int HD>
__global__ void __launch_bounds__(256) gattn_avg_weights_kernel(GArgs a, T* w) {
  __shared__ float qs[16][HD + 1], ks[16][HD + 1];
  const int b = blockIdx.z, i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
  const int i = i0 + ti, j = j0 + tj;
  const float sl2 = a.scale * kLog2e;
  float acc = 0.f;
  for (int h = 0; h < a.nH; ++h) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 16 * HD; idx += 256) {
      const int r = idx / HD, c = idx % HD;
      qs[r][c] = (i0 + r < a.Nq) ? Io<T>::ld((const T*)a.q + ((int64_t)b * a.Nq + i0 + r) * a.ldq + h * HD + c) : 0.f;
      ks[r][c] = (j0 + r < a.Nk) ? Io<T>::ld((const T*)a.k + ((int64_t)b * a.Nk + j0 + r) * a.ldk + h * HD + c) : 0.f;
    }
    __syncthreads();
    if (i < a.Nq && j < a.Nk) {
      float s = 0.f;
#pragma unroll 8
      for (int c = 0; c < HD; ++c) s = fmaf(qs[ti][c], ks[tj][c], s);
      acc += exp2f(s * sl2 - a.lse[((int64_t)b * a.nH + h) * a.Nq + i] * kLog2e);
    }
  }
  if (i < a.Nq && j < a.Nk) Io<T>::st(w + ((int64_t)b * a.Nq + i) * a.Nk + j, acc / (float)a.nH);
}

int check_common(const char* what, const GArgs& a, int head_dim, int dtype) {
  BSW_REQUIRE(a.B > 0 && a.Nq > 0 && a.Nk > 0 && a.nH > 0, "%s: non-positive dimension", what);
  BSW_REQUIRE(head_dim == 32 || head_dim == 64, "%s: head_dim %d not built (32 and 64 are)", what, head_dim);
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "%s: bad dtype %d", what, dtype);
  BSW_REQUIRE(a.B <= 65535 && a.nH <= 65535, "%s: batch / head count beyond the launch grid", what);
  return B200SWIN_OK;
}
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int HD>
int fwd_hd(const GArgs& a, int dtype, cudaStream_t st) {
  if (dtype == B200SWIN_BF16) {
    dim3 grid((a.Nq + 63) / 64, a.nH, a.B);
    gattn_fwd_mma_kernel<HD><<<grid, 128, 0, st>>>(a);
  } else {
    dim3 grid((a.Nq + 127) / 128, a.nH, a.B);
    gattn_fwd_simt_kernel<float, HD><<<grid, 128, 0, st>>>(a);
  }
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

template <int HD>
int bwd_hd(const GArgs& a, int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)a.B * a.Nq * a.nH;
  if (dtype == B200SWIN_BF16) {
    gattn_prep_kernel<bf16, HD><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    BSW_LAUNCH_CHECK();
    gattn_dkv_mma_kernel<HD><<<dim3((a.Nk + 63) / 64, a.nH, a.B), 128, 0, st>>>(a);
    BSW_LAUNCH_CHECK();
    gattn_dq_mma_kernel<HD><<<dim3((a.Nq + 63) / 64, a.nH, a.B), 128, 0, st>>>(a);
    BSW_LAUNCH_CHECK();
  } else {
    gattn_prep_kernel<float, HD><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a);
    BSW_LAUNCH_CHECK();
    gattn_bwd_simt_kernel<float, HD, 0><<<dim3((a.Nq + 127) / 128, a.nH, a.B), 128, 0, st>>>(a);
    BSW_LAUNCH_CHECK();
    gattn_bwd_simt_kernel<float, HD, 1><<<dim3((a.Nk + 127) / 128, a.nH, a.B), 128, 0, st>>>(a);
    BSW_LAUNCH_CHECK();
    gattn_bwd_simt_kernel<float, HD, 2><<<dim3((a.Nk + 127) / 128, a.nH, a.B), 128, 0, st>>>(a);
    BSW_LAUNCH_CHECK();
  }
  return B200SWIN_OK;
}

}  // namespace
}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_mha_fwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, void* out,
                                int64_t ldo, float* lse, int B, int Nq, int Nk, int nH, int head_dim, float scale, int dtype,
                                void* stream) {
  BSW_REQUIRE(q && k && v && out && lse, "mha_fwd: null pointer");
  GArgs a = {};
  a.q = q; a.k = k; a.v = v; a.o = out; a.lse = lse;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo;
  a.B = B; a.Nq = Nq; a.Nk = Nk; a.nH = nH; a.scale = scale;
  int rc = check_common("mha_fwd", a, head_dim, dtype);
  if (rc) return rc;
  const int64_t width = (int64_t)nH * head_dim;
  BSW_REQUIRE(ldq >= width && ldk >= width && ldv >= width && ldo >= width, "mha_fwd: row stride smaller than nH * head_dim");
  BSW_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out) && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
                  ldo % 8 == 0, "mha_fwd: tensors must be 16-byte aligned with row strides that are multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  return head_dim == 64 ? fwd_hd<64>(a, dtype, st) : fwd_hd<32>(a, dtype, st);
}

extern "C" size_t b200swin_mha_bwd_workspace_bytes(int B, int Nq, int nH) {
  return (size_t)((int64_t)B * Nq * nH) * sizeof(float);
}

extern "C" int b200swin_mha_bwd(const void* q, const void* k, const void* v, int64_t ldq, int64_t ldk, int64_t ldv, const void* out,
                                int64_t ldo, const void* dout, int64_t lddo, const float* lse, void* dq, void* dk, void* dv,
                                int64_t lddq, int64_t lddk, int64_t lddv, int B, int Nq, int Nk, int nH, int head_dim,
                                float scale, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  BSW_REQUIRE(q && k && v && out && dout && lse && dq && dk && dv, "mha_bwd: null pointer");
  GArgs a = {};
  a.q = q; a.k = k; a.v = v; a.out = out; a.dout = dout; a.lse = const_cast<float*>(lse);
  a.dq = dq; a.dk = dk; a.dv = dv; a.dvec = (float*)workspace;
  a.ldq = ldq; a.ldk = ldk; a.ldv = ldv; a.ldo = ldo; a.lddo = lddo; a.lddq = lddq; a.lddk = lddk; a.lddv = lddv;
  a.B = B; a.Nq = Nq; a.Nk = Nk; a.nH = nH; a.scale = scale;
  int rc = check_common("mha_bwd", a, head_dim, dtype);
  if (rc) return rc;
  BSW_REQUIRE(workspace && workspace_bytes >= b200swin_mha_bwd_workspace_bytes(B, Nq, nH),
              "mha_bwd: workspace too small (see b200swin_mha_bwd_workspace_bytes)");
  const int64_t width = (int64_t)nH * head_dim;
  BSW_REQUIRE(ldq >= width && ldk >= width && ldv >= width && ldo >= width && lddo >= width && lddq >= width && lddk >= width &&
                  lddv >= width, "mha_bwd: row stride smaller than nH * head_dim");
  BSW_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(out) && aligned16(dout) && aligned16(dq) &&
                  aligned16(dk) && aligned16(dv) && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 &&
                  lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0,
              "mha_bwd: tensors must be 16-byte aligned with row strides that are multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  return head_dim == 64 ? bwd_hd<64>(a, dtype, st) : bwd_hd<32>(a, dtype, st);
}

extern "C" int b200swin_mha_avg_weights(const void* q, const void* k, int64_t ldq, int64_t ldk, const float* lse, void* weights,
                                        int B, int Nq, int Nk, int nH, int head_dim, float scale, int dtype, void* stream) {
  BSW_REQUIRE(q && k && lse && weights, "mha_avg_weights: null pointer");
  GArgs a = {};
  a.q = q; a.k = k; a.lse = const_cast<float*>(lse);
  a.ldq = ldq; a.ldk = ldk;
  a.B = B; a.Nq = Nq; a.Nk = Nk; a.nH = nH; a.scale = scale;
  int rc = check_common("mha_avg_weights", a, head_dim, dtype);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((Nk + 15) / 16, (Nq + 15) / 16, B);
  if (dtype == B200SWIN_BF16) {
    if (head_dim == 64) gattn_avg_weights_kernel<bf16, 64><<<grid, 256, 0, st>>>(a, (bf16*)weights);
    else gattn_avg_weights_kernel<bf16, 32><<<grid, 256, 0, st>>>(a, (bf16*)weights);
  } else {
    if (head_dim == 64) gattn_avg_weights_kernel<float, 64><<<grid, 256, 0, st>>>(a, (float*)weights);
    else gattn_avg_weights_kernel<float, 32><<<grid, 256, 0, st>>>(a, (float*)weights);
  }
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
