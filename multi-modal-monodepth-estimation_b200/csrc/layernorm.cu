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
  if (dtype == B200SWIN_F32)
    return ln_fwd_launch<float>(x, residual, gamma, beta, row_scale, rows_per_scale, y, mean, rstd, rows, C, eps, st);
  return ln_fwd_launch<__nv_bfloat16>(x, residual, gamma, beta, row_scale, rows_per_scale, y, mean, rstd, rows, C,
                                      eps, st);
}

extern "C" size_t b200swin_ln_bwd_workspace_bytes(int64_t rows, int C) {
  return (size_t)ln_bwd_grid(rows) * 2 * (size_t)C * sizeof(float);
}

extern "C" int b200swin_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean,
                               const float* rstd, const float* row_scale, int64_t rows_per_scale, void* dx,
                               float* dgamma, float* dbeta, int64_t rows, int C, int dtype, void* workspace,
                               size_t workspace_bytes, void* stream) {
  BSW_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "ln_bwd: null pointer");
  BSW_REQUIRE(rows > 0 && C > 0 && C % 4 == 0 && C <= 3072, "ln_bwd: bad rows/C");
  BSW_REQUIRE(!row_scale || rows_per_scale > 0, "ln_bwd: rows_per_scale must be > 0 with row_scale");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "ln_bwd: bad dtype %d", dtype);
  int grid = ln_bwd_grid(rows);
  BSW_REQUIRE(workspace_bytes >= (size_t)grid * 2 * C * sizeof(float), "ln_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
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
