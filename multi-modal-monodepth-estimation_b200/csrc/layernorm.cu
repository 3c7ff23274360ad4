// LayerNorm fused with the DropPath scale and the residual add of the post-norm Swin-V2 block.
// Replaces LayerNormFP32.forward (models/swin_transformer_v2.py:41-47) and
//   x = shortcut + drop_path(norm1(x))   (:472-474),   x = x + drop_path(norm2(mlp(x)))   (:482-483)
// HBM-bound: fwd reads x (+residual) and writes y once; statistics in fp32 (two-pass, from registers).
// One warp per row, the row cached in registers for C <= 1024 (NV*128 columns).
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

constexpr int kLnThreads = 256;
constexpr int kLnWarps = kLnThreads / 32;

// R rows per warp and iteration: all loads of the R rows (x and the residual) are issued before any arithmetic, so a
// warp keeps R * NV * (8 or 16) bytes per lane in flight instead of one 8-byte load (the C = 128 stage is pure HBM
// streaming: one row is only 256 bytes).
template <typename T, int NV, int R>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ residual, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ row_scale, int64_t rows_per_scale,
              T* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int C,
              float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  const float inv_c = 1.0f / (float)C;
  for (int64_t row0 = warp0 * R; row0 < rows; row0 += nwarps * R) {
    float v[R][NV][4], rs[R][NV][4];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + k * 128;
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[r][k][e] = 0.f; rs[r][k][e] = 0.f; }
        if (row < rows && c < C) {
          ld4(x + row * C + c, v[r][k]);
          if (residual) ld4(residual + row * C + c, rs[r][k]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
      if (row >= rows) break;                       // warp-uniform
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) s += (v[r][k][0] + v[r][k][1]) + (v[r][k][2] + v[r][k][3]);
      const float mean = warp_sum(s) * inv_c;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (lane * 4 + k * 128 < C) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float d = v[r][k][e] - mean; q = fmaf(d, d, q); }
        }
      }
      const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
      const float sc = row_scale ? row_scale[row / rows_per_scale] : 1.0f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + k * 128;
        if (c < C) {
          float g[4], b[4], o[4];
          ld4(gamma + c, g);                       // L1-resident after the first row
          ld4(beta + c, b);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = ((v[r][k][e] - mean) * rstd * g[e] + b[e]) * sc + rs[r][k][e];
          st4(y + row * C + c, o);
        }
      }
      if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
  }
}

// Backward: dx per row, and per-block partial sums of dgamma/dbeta (reduced by ln_param_reduce_kernel).
template <typename T, int NV, int R>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ gamma,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
              const float* __restrict__ row_scale, int64_t rows_per_scale, T* __restrict__ dx,
              float* __restrict__ part /*[grid][2][C]*/, int64_t rows, int C) {
  extern __shared__ float sh[];   // [2][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  const float inv_c = 1.0f / (float)C;
  float dg[NV][4], db[NV][4], gm[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    int c = lane * 4 + k * 128;
#pragma unroll
    for (int e = 0; e < 4; ++e) { dg[k][e] = 0.f; db[k][e] = 0.f; gm[k][e] = 0.f; }
    if (c < C) ld4(gamma + c, gm[k]);
  }
  for (int64_t row0 = warp0 * R; row0 < rows; row0 += nwarps * R) {
    float xv[R][NV][4], g[R][NV][4], mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
      mean[r] = row < rows ? mean_in[row] : 0.f;
      rstd[r] = row < rows ? rstd_in[row] : 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + k * 128;
#pragma unroll
        for (int e = 0; e < 4; ++e) { xv[r][k][e] = 0.f; g[r][k][e] = 0.f; }
        if (row < rows && c < C) {
          ld4(x + row * C + c, xv[r][k]);
          ld4(dy + row * C + c, g[r][k]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
      if (row >= rows) break;                       // warp-uniform
      const float sc = row_scale ? row_scale[row / rows_per_scale] : 1.0f;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xh = (xv[r][k][e] - mean[r]) * rstd[r];
          xv[r][k][e] = xh;
          g[r][k][e] *= sc;
          dg[k][e] = fmaf(g[r][k][e], xh, dg[k][e]);
          db[k][e] += g[r][k][e];
          const float gg = g[r][k][e] * gm[k][e];
          s1 += gg;
          s2 = fmaf(gg, xh, s2);
        }
      }
      const float m1 = warp_sum(s1) * inv_c, m2 = warp_sum(s2) * inv_c;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + k * 128;
        if (c < C) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = rstd[r] * (g[r][k][e] * gm[k][e] - m1 - xv[r][k][e] * m2);
          st4(dx + row * C + c, o);
        }
      }
    }
  }
  // fixed-order cross-warp reduction in shared memory -> one partial row per block
  for (int c = threadIdx.x; c < 2 * C; c += kLnThreads) sh[c] = 0.f;
  __syncthreads();
  for (int w = 0; w < kLnWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        int c = lane * 4 + k * 128;
        if (c < C) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { sh[c + e] += dg[k][e]; sh[C + c + e] += db[k][e]; }
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < 2 * C; c += kLnThreads) part[(int64_t)blockIdx.x * 2 * C + c] = sh[c];
}

// dgamma / dbeta = fixed-order sum of the per-block partial rows: 32 columns x 8 part lanes per block
__global__ void __launch_bounds__(256)
ln_param_reduce_kernel(const float* __restrict__ part, int nparts, int C, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;              // over 2*C
  float s = 0.f;
  if (c < 2 * C)
    for (int p = py; p < nparts; p += 8) s += part[(int64_t)p * 2 * C + c];
  sh[py][cx] = s;
  __syncthreads();
  if (py == 0 && c < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][cx];
    if (c < C) dgamma[c] = t; else dbeta[c - C] = t;
  }
}

static int ln_grid(int64_t rows) {
  int64_t blocks = (rows + kLnWarps - 1) / kLnWarps;
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}
static int ln_bwd_grid(int64_t rows) {
  int64_t blocks = (rows + 4 * kLnWarps - 1) / (4 * kLnWarps);   // >= 4 rows per warp before a partial row is paid
  int64_t cap = (int64_t)sm_count() * 4;
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

template <typename T>
static int ln_fwd_launch(const void* x, const void* residual, const float* gamma, const float* beta,
                         const float* row_scale, int64_t rps, void* y, float* mean, float* rstd, int64_t rows, int C,
                         float eps, cudaStream_t st) {
  int grid = ln_grid(rows);
#define LN_FWD(NV, R)                                                                                         \
  ln_fwd_kernel<T, NV, R><<<grid, kLnThreads, 0, st>>>((const T*)x, (const T*)residual, gamma, beta, row_scale, \
                                                       rps, (T*)y, mean, rstd, rows, C, eps)
  if (C <= 128) LN_FWD(1, 4);
  else if (C <= 256) LN_FWD(2, 2);
  else if (C <= 512) LN_FWD(4, 1);
  else if (C <= 1024) LN_FWD(8, 1);
  else if (C <= 2048) LN_FWD(16, 1);
  else LN_FWD(24, 1);
#undef LN_FWD
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

template <typename T>
static int ln_bwd_launch(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                         const float* row_scale, int64_t rps, void* dx, float* part, int grid, int64_t rows, int C,
                         cudaStream_t st) {
  size_t smem = (size_t)2 * C * sizeof(float);
#define LN_BWD(NV, R)                                                                                           \
  ln_bwd_kernel<T, NV, R><<<grid, kLnThreads, smem, st>>>((const T*)dy, (const T*)x, gamma, mean, rstd, row_scale, \
                                                          rps, (T*)dx, part, rows, C)
  if (C <= 128) LN_BWD(1, 4);
  else if (C <= 256) LN_BWD(2, 2);
  else if (C <= 512) LN_BWD(4, 1);
  else if (C <= 1024) LN_BWD(8, 1);
  else if (C <= 2048) LN_BWD(16, 1);
  else LN_BWD(24, 1);
#undef LN_BWD
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}


// ------------------------------------------------------------------------------------------------------------------
// bf16 streaming kernels (C % 8 == 0, C <= 1536): the row stays PACKED (uint4 = 8 bf16) in registers until it is used,
// every lane moves 16 bytes per load / store, and U passes (a pass = 32 / LPR rows per warp) are requested before
// the first is reduced, so a warp keeps ~4 KB in flight with 32 registers.  gamma / beta live in shared memory.
// LPR = lanes per row (16 for C <= 128: two rows per pass), NV = 16-byte chunks per lane and row.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
    f[2 * e] = t.x; f[2 * e + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
    w[e] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

template <int NV, int LPR, int U>
__global__ void __launch_bounds__(kLnThreads, NV <= 3 ? 3 : 2)
ln_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ residual,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ row_scale, int64_t rows_per_scale, __nv_bfloat16* __restrict__ y,
                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int C, float eps) {
  extern __shared__ float sh[];   // gamma[C] | beta[C]
  for (int c = threadIdx.x; c < C; c += kLnThreads) { sh[c] = gamma[c]; sh[C + c] = beta[c]; }
  __syncthreads();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int lr = lane % LPR, sub = lane / LPR;
  const int chunks = C >> 3;
  const float inv_c = 1.0f / (float)C;
  const int64_t npass = (rows + RPW - 1) / RPW;
  const int64_t gwarp = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  for (int64_t p0 = gwarp * U; p0 < npass; p0 += nwarps * U) {
    uint4 xv[U][NV], rv[U][NV];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = (p0 + u) * RPW + sub;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        xv[u][k] = make_uint4(0u, 0u, 0u, 0u);
        rv[u][k] = make_uint4(0u, 0u, 0u, 0u);
        if (row < rows && ch < chunks) {
          xv[u][k] = ldg16(x + row * C + ch * 8);
          if (residual) rv[u][k] = ldg16(residual + row * C + ch * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if ((p0 + u) * RPW >= rows) break;            // warp-uniform
      const int64_t row = (p0 + u) * RPW + sub;
      const bool live = row < rows;
      float f[NV][8];
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        unpack8(xv[u][k], f[k]);
#pragma unroll
        for (int e = 0; e < 8; e += 2) s += f[k][e] + f[k][e + 1];
      }
      const float mean = row_sum<LPR>(s) * inv_c;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (k * LPR + lr < chunks) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { const float d = f[k][e] - mean; q = fmaf(d, d, q); }
        }
      }
      const float rstd = rsqrtf(row_sum<LPR>(q) * inv_c + eps);
      const float sc = (row_scale && live) ? row_scale[row / rows_per_scale] : 1.0f;
      const float a = rstd * sc, b0 = -mean * a;    // ((x - mean) rstd g + b) sc = (x a + b0) g + b sc
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        if (live && ch < chunks) {
          float r[8], o[8];
          unpack8(rv[u][k], r);
          const float4 g0 = *reinterpret_cast<const float4*>(sh + ch * 8), g1 = *reinterpret_cast<const float4*>(sh + ch * 8 + 4);
          const float4 b0v = *reinterpret_cast<const float4*>(sh + C + ch * 8), b1v = *reinterpret_cast<const float4*>(sh + C + ch * 8 + 4);
          const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float b[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = fmaf(fmaf(f[k][e], a, b0), g[e], fmaf(b[e], sc, r[e]));
          *reinterpret_cast<uint4*>(y + row * C + ch * 8) = pack8(o);
        }
      }
      if (lr == 0 && live) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
  }
}

// Backward.  Phase A accumulates dgamma / dbeta and the two row sums from the packed registers, phase B unpacks again
// and writes dx (recomputing x_hat costs three FMAs per element and saves 16 * NV live registers).  COLSUM adds the
// column sums of dx -- the bias gradient of the Linear that produced x (proj / fc2) -- to the partial rows.
template <int NV, int LPR, int U, bool COLSUM>
__global__ void __launch_bounds__(kLnThreads, NV <= 3 ? 2 : 1)
ln_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const float* __restrict__ gamma, const float* __restrict__ mean_in,
                   const float* __restrict__ rstd_in, const float* __restrict__ row_scale, int64_t rows_per_scale,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ part /*[grid][NP][C]*/, int64_t rows, int C) {
  constexpr int NP = COLSUM ? 3 : 2;
  extern __shared__ float sh[];   // gamma[C] | reduction scratch [NP][C]
  for (int c = threadIdx.x; c < C; c += kLnThreads) sh[c] = gamma[c];
  for (int c = threadIdx.x; c < NP * C; c += kLnThreads) sh[C + c] = 0.f;
  __syncthreads();
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lr = lane % LPR, sub = lane / LPR;
  const int chunks = C >> 3;
  const float inv_c = 1.0f / (float)C;
  const int64_t npass = (rows + RPW - 1) / RPW;
  const int64_t gwarp = (int64_t)blockIdx.x * kLnWarps + warp;
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  float dg[NV][8], db[NV][8], dc[COLSUM ? NV : 1][8];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) { dg[k][e] = 0.f; db[k][e] = 0.f; if (COLSUM) dc[k][e] = 0.f; }

  for (int64_t p0 = gwarp * U; p0 < npass; p0 += nwarps * U) {
    uint4 xv[U][NV], gv[U][NV];
    float mean[U], rstd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t row = (p0 + u) * RPW + sub;
      mean[u] = row < rows ? mean_in[row] : 0.f;
      rstd[u] = row < rows ? rstd_in[row] : 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        xv[u][k] = make_uint4(0u, 0u, 0u, 0u);
        gv[u][k] = make_uint4(0u, 0u, 0u, 0u);
        if (row < rows && ch < chunks) {
          xv[u][k] = ldg16(x + row * C + ch * 8);
          gv[u][k] = ldg16(dy + row * C + ch * 8);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if ((p0 + u) * RPW >= rows) break;            // warp-uniform
      const int64_t row = (p0 + u) * RPW + sub;
      const bool live = row < rows;
      const float sc = (row_scale && live) ? row_scale[row / rows_per_scale] : 1.0f;
      const float a = rstd[u], b0 = -mean[u] * rstd[u];             // x_hat = x a + b0
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        if (ch < chunks) {                          // dead rows carry zeros
          float xf[8], gf[8];
          unpack8(xv[u][k], xf);
          unpack8(gv[u][k], gf);
          const float4 g0 = *reinterpret_cast<const float4*>(sh + ch * 8), g1 = *reinterpret_cast<const float4*>(sh + ch * 8 + 4);
          const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xh = live ? fmaf(xf[e], a, b0) : 0.f;
            const float g = gf[e] * sc;
            dg[k][e] = fmaf(g, xh, dg[k][e]);
            db[k][e] += g;
            const float gg = g * gm[e];
            s1 += gg;
            s2 = fmaf(gg, xh, s2);
          }
        }
      }
      const float m1 = row_sum<LPR>(s1) * inv_c, m2 = row_sum<LPR>(s2) * inv_c;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        if (live && ch < chunks) {
          float xf[8], gf[8], o[8];
          unpack8(xv[u][k], xf);
          unpack8(gv[u][k], gf);
          const float4 g0 = *reinterpret_cast<const float4*>(sh + ch * 8), g1 = *reinterpret_cast<const float4*>(sh + ch * 8 + 4);
          const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float xh = fmaf(xf[e], a, b0);
            o[e] = a * (gf[e] * sc * gm[e] - m1 - xh * m2);
            if (COLSUM) dc[k][e] += o[e];
          }
          *reinterpret_cast<uint4*>(dx + row * C + ch * 8) = pack8(o);
        }
      }
    }
  }
  // two rows per pass: lanes l and l + 16 own the same columns
  if (LPR == 16) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dg[k][e] += __shfl_xor_sync(0xffffffffu, dg[k][e], 16);
        db[k][e] += __shfl_xor_sync(0xffffffffu, db[k][e], 16);
        if (COLSUM) dc[k][e] += __shfl_xor_sync(0xffffffffu, dc[k][e], 16);
      }
  }
  // fixed-order cross-warp reduction in shared memory -> one partial row set per block
  float* red = sh + C;
  for (int w = 0; w < kLnWarps; ++w) {
    if (warp == w && sub == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        if (ch < chunks) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            red[ch * 8 + e] += dg[k][e];
            red[C + ch * 8 + e] += db[k][e];
            if (COLSUM) red[2 * C + ch * 8 + e] += dc[k][e];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < NP * C; c += kLnThreads) part[(int64_t)blockIdx.x * NP * C + c] = red[c];
}

// dgamma / dbeta (/ dcolsum) = fixed-order sum of the per-block partial rows: 32 columns x 8 part lanes per block
__global__ void __launch_bounds__(256)
ln_param_reduce3_kernel(const float* __restrict__ part, int nparts, int C, int NP, float* __restrict__ o0,
                        float* __restrict__ o1, float* __restrict__ o2) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;              // over NP*C
  float s = 0.f;
  if (c < NP * C)
    for (int p = py; p < nparts; p += 8) s += part[(int64_t)p * NP * C + c];
  sh[py][cx] = s;
  __syncthreads();
  if (py == 0 && c < NP * C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][cx];
    if (c < C) o0[c] = t; else if (c < 2 * C) o1[c - C] = t; else o2[c - 2 * C] = t;
  }
}

static bool ln_fast_ok(int C, int dtype) { return dtype == B200SWIN_BF16 && C % 8 == 0 && C <= 1536; }
static int ln_fast_bwd_grid(int64_t rows, int C) {
  int64_t blocks = (rows + 4 * kLnWarps - 1) / (4 * kLnWarps);
  int64_t cap = (int64_t)sm_count() * (C <= 768 ? 2 : 1);         // exactly the resident blocks
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}

// (NV, LPR, U) by row width: ~4 KB (two tensors) in flight per warp
#define LN_FAST_DISPATCH(C, X)                   \
  do {                                           \
    if ((C) <= 128) X(1, 16, 4);                 \
    else if ((C) <= 256) X(1, 32, 4);            \
    else if ((C) <= 512) X(2, 32, 2);            \
    else if ((C) <= 768) X(3, 32, 1);            \
    else if ((C) <= 1024) X(4, 32, 1);           \
    else X(6, 32, 1);                            \
  } while (0)

static int ln_fwd_fast(const void* x, const void* residual, const float* gamma, const float* beta,
                       const float* row_scale, int64_t rps, void* y, float* mean, float* rstd, int64_t rows, int C,
                       float eps, cudaStream_t st) {
  const size_t smem = (size_t)2 * C * sizeof(float);
#define X(NV, LPR, U)                                                                                                \
  do {                                                                                                               \
    const int64_t npass = (rows + (32 / LPR) - 1) / (32 / LPR);                                                      \
    int64_t blocks = (npass + (int64_t)U * kLnWarps - 1) / ((int64_t)U * kLnWarps);                                  \
    const int64_t cap = (int64_t)sm_count() * (NV <= 3 ? 3 : 2);     /* exactly the resident blocks: no second wave */ \
    const int grid = (int)(blocks < cap ? blocks : cap);                                                             \
    ln_fwd_bf16_kernel<NV, LPR, U><<<grid, kLnThreads, smem, st>>>(                                                  \
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, gamma, beta, row_scale, rps, (__nv_bfloat16*)y, mean, \
        rstd, rows, C, eps);                                                                                         \
  } while (0)
  LN_FAST_DISPATCH(C, X);
#undef X
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

static int ln_bwd_fast(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const float* row_scale, int64_t rps, void* dx, float* part, int grid, bool colsum, int64_t rows,
                       int C, cudaStream_t st) {
  const size_t smem = (size_t)(1 + (colsum ? 3 : 2)) * C * sizeof(float);
#define X(NV, LPR, U)                                                                                                \
  do {                                                                                                               \
    if (colsum)                                                                                                      \
      ln_bwd_bf16_kernel<NV, LPR, U, true><<<grid, kLnThreads, smem, st>>>(                                          \
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, mean, rstd, row_scale, rps, (__nv_bfloat16*)dx,  \
          part, rows, C);                                                                                            \
    else                                                                                                             \
      ln_bwd_bf16_kernel<NV, LPR, U, false><<<grid, kLnThreads, smem, st>>>(                                         \
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, mean, rstd, row_scale, rps, (__nv_bfloat16*)dx,  \
          part, rows, C);                                                                                            \
  } while (0)
  LN_FAST_DISPATCH(C, X);
#undef X
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_ln_fwd(const void* x, const void* residual, const float* gamma, const float* beta,
                               const float* row_scale, int64_t rows_per_scale, void* y, float* mean, float* rstd,
                               int64_t rows, int C, float eps, int dtype, void* stream) {
  BSW_REQUIRE(x && gamma && beta && y && mean && rstd, "ln_fwd: null pointer");
  BSW_REQUIRE(rows >= 0 && C > 0 && C % 4 == 0 && C <= 3072, "ln_fwd: C=%d must be a multiple of 4, <= 3072", C);
  BSW_REQUIRE(!row_scale || rows_per_scale > 0, "ln_fwd: rows_per_scale must be > 0 with row_scale");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "ln_fwd: bad dtype %d", dtype);
  if (rows == 0) return B200SWIN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (ln_fast_ok(C, dtype))
    return ln_fwd_fast(x, residual, gamma, beta, row_scale, rows_per_scale, y, mean, rstd, rows, C, eps, st);
  if (dtype == B200SWIN_F32)
    return ln_fwd_launch<float>(x, residual, gamma, beta, row_scale, rows_per_scale, y, mean, rstd, rows, C, eps, st);
  return ln_fwd_launch<__nv_bfloat16>(x, residual, gamma, beta, row_scale, rows_per_scale, y, mean, rstd, rows, C,
                                      eps, st);
}

extern "C" size_t b200swin_ln_bwd_workspace_bytes(int64_t rows, int C) {
  const int g0 = ln_bwd_grid(rows), g1 = ln_fast_bwd_grid(rows, C);
  return (size_t)(g0 > g1 ? g0 : g1) * 3 * (size_t)C * sizeof(float);
}

extern "C" int b200swin_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                               const float* rstd, const float* row_scale, int64_t rows_per_scale, void* dx,
                               float* dgamma, float* dbeta, float* dcolsum, int64_t rows, int C, int dtype,
                               void* workspace, size_t workspace_bytes, void* stream) {
  BSW_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "ln_bwd: null pointer");
  BSW_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= 3072, "ln_bwd: bad rows/C");
  BSW_REQUIRE(!row_scale || rows_per_scale > 0, "ln_bwd: rows_per_scale must be > 0 with row_scale");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "ln_bwd: bad dtype %d", dtype);
  BSW_REQUIRE(workspace_bytes >= b200swin_ln_bwd_workspace_bytes(rows, C), "ln_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (ln_fast_ok(C, dtype)) {
    const int grid = ln_fast_bwd_grid(rows, C);
    const int NP = dcolsum ? 3 : 2;
    rc = ln_bwd_fast(dy, x, gamma, mean, rstd, row_scale, rows_per_scale, dx, (float*)workspace, grid, dcolsum != nullptr,
                     rows, C, st);
    if (rc) return rc;
    ln_param_reduce3_kernel<<<(NP * C + 31) / 32, 256, 0, st>>>((const float*)workspace, grid, C, NP, dgamma, dbeta,
                                                                dcolsum);
    BSW_LAUNCH_CHECK();
    return B200SWIN_OK;
  }
  BSW_REQUIRE(!dcolsum, "ln_bwd: the fused column sum of dx is built for bf16 tensors with C %% 8 == 0, C <= 1536");
  int grid = ln_bwd_grid(rows);
  if (dtype == B200SWIN_F32)
    rc = ln_bwd_launch<float>(dy, x, gamma, mean, rstd, row_scale, rows_per_scale, dx, (float*)workspace, grid, rows,
                              C, st);
  else
    rc = ln_bwd_launch<__nv_bfloat16>(dy, x, gamma, mean, rstd, row_scale, rows_per_scale, dx, (float*)workspace,
                                      grid, rows, C, st);
  if (rc) return rc;
  ln_param_reduce_kernel<<<(2 * C + 31) / 32, 256, 0, st>>>((const float*)workspace, grid, C, dgamma, dbeta);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
