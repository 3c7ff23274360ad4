// SiLog depth loss, forward + backward (replaces utils/criterion.py:15-21 and its autograd).
//
// HBM-bound: fwd reads pred+target (8 B/px in fp32), bwd reads both and writes grad (12 B/px).
// fwd = one grid-wide masked reduction of (sum d, sum d^2, count) with per-thread fp32 partials,
// warp-shuffle + double block partials, and a fixed-order final reduce (deterministic, no atomics,
// no host sync - the reference's boolean indexing forces a nonzero() round trip per call).
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

constexpr int kSilogThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kSilogThreads)
silog_partial_kernel(const T* __restrict__ pred, const float* __restrict__ target, int64_t n,
                     double* __restrict__ partials /*[grid][3]*/) {
  float s1 = 0.f, s2 = 0.f, cnt = 0.f;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
  int64_t n4 = vec_ok ? (n >> 2) : 0;
  for (int64_t i = tid; i < n4; i += nthreads) {
    float p[4], t[4];
    ld4(pred + 4 * i, p);
    ld4(target + 4 * i, t);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (t[k] > 0.f) {
        float d = logf(t[k]) - logf(p[k]);
        s1 += d;
        s2 = fmaf(d, d, s2);
        cnt += 1.f;
      }
    }
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nthreads) {
    float t = target[i];
    if (t > 0.f) {
      float d = logf(t) - logf(Io<T>::ld(pred + i));
      s1 += d;
      s2 = fmaf(d, d, s2);
      cnt += 1.f;
    }
  }
  double d1 = warp_sum_d((double)s1), d2 = warp_sum_d((double)s2), dc = warp_sum_d((double)cnt);
  __shared__ double sh[3][kSilogThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = d1; sh[1][warp] = d2; sh[2][warp] = dc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
#pragma unroll
    for (int w = 0; w < kSilogThreads / 32; ++w) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
    partials[3 * blockIdx.x + 0] = a;
    partials[3 * blockIdx.x + 1] = b;
    partials[3 * blockIdx.x + 2] = c;
  }
}

__global__ void __launch_bounds__(256)
silog_final_kernel(const double* __restrict__ partials, int nblocks, float lambd,
                   float* __restrict__ loss, float* __restrict__ stats) {
  // fixed-order tree: thread t sums partials t, t+256, ...; then shared-memory tree.
  double a = 0, b = 0, c = 0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    a += partials[3 * i]; b += partials[3 * i + 1]; c += partials[3 * i + 2];
  }
  __shared__ double sh[3][256];
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b; sh[2][threadIdx.x] = c;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
      sh[2][threadIdx.x] += sh[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double n = sh[2][0];
    double mean = sh[0][0] / n;            // n == 0 -> NaN, like the reference's empty mean()
    double mean2 = sh[1][0] / n;
    double l = sqrt(mean2 - (double)lambd * mean * mean);
    *loss = (float)l;
    stats[0] = (float)mean; stats[1] = (float)n; stats[2] = (float)l; stats[3] = (float)mean2;
  }
}

template <typename T>
__global__ void __launch_bounds__(kSilogThreads)
silog_bwd_kernel(const T* __restrict__ pred, const float* __restrict__ target, int64_t n, float lambd,
                 const float* __restrict__ stats, const float* __restrict__ grad_out, T* __restrict__ grad) {
  const float mean = stats[0], cnt = stats[1], l = stats[2];
  // dL/dpred_i = -(d_i - lambd*mean) / (n * L * pred_i)   
  const float coef = -(*grad_out) / (cnt * l);
  const float lm = lambd * mean;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) |
                        reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  int64_t n4 = vec_ok ? (n >> 2) : 0;
  for (int64_t i = tid; i < n4; i += nthreads) {
    float p[4], t[4], g[4];
    ld4(pred + 4 * i, p);
    ld4(target + 4 * i, t);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      g[k] = (t[k] > 0.f) ? coef * (logf(t[k]) - logf(p[k]) - lm) / p[k] : 0.f;
    st4(grad + 4 * i, g);
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nthreads) {
    float t = target[i], p = Io<T>::ld(pred + i);
    Io<T>::st(grad + i, (t > 0.f) ? coef * (logf(t) - logf(p) - lm) / p : 0.f);
  }
}

static int silog_grid(int64_t n) {
  int64_t want = (n / 4 + kSilogThreads - 1) / kSilogThreads;      // one float4 per thread at least
  int64_t cap = (int64_t)sm_count() * 8;                           // 8 x 256 threads resident per SM
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace b200swin

using namespace b200swin;

extern "C" size_t b200swin_silog_workspace_bytes(int64_t n) {
  // exactly what b200swin_silog_fwd checks: one (sum d, sum d^2, count) triple of doubles per CTA of ITS grid on the
  // current device (floor of one CTA so that n == 0 still hands out a valid pointer)
  return (size_t)silog_grid(n < 0 ? 0 : n) * 3 * sizeof(double);
}

extern "C" int b200swin_silog_fwd(const void* pred, int pred_dtype, const float* target, int64_t n, float lambd,
                                  float* loss, float* stats, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  BSW_REQUIRE(pred && target && loss && stats && workspace, "silog_fwd: null pointer");
  BSW_REQUIRE(n >= 0, "silog_fwd: n < 0");
  BSW_REQUIRE(pred_dtype == B200SWIN_F32 || pred_dtype == B200SWIN_BF16, "silog_fwd: bad dtype %d", pred_dtype);
  int grid = silog_grid(n);
  BSW_REQUIRE(workspace_bytes >= (size_t)grid * 3 * sizeof(double), "silog_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  double* part = (double*)workspace;
  if (pred_dtype == B200SWIN_F32)
    silog_partial_kernel<float><<<grid, kSilogThreads, 0, st>>>((const float*)pred, target, n, part);
  else
    silog_partial_kernel<__nv_bfloat16><<<grid, kSilogThreads, 0, st>>>((const __nv_bfloat16*)pred, target, n, part);
  BSW_LAUNCH_CHECK();
  silog_final_kernel<<<1, 256, 0, st>>>(part, grid, lambd, loss, stats);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

extern "C" int b200swin_silog_bwd(const void* pred, int pred_dtype, const float* target, int64_t n, float lambd,
                                  const float* stats, const float* grad_out, void* grad_pred, void* stream) {
  BSW_REQUIRE(pred && target && stats && grad_out && grad_pred, "silog_bwd: null pointer");
  BSW_REQUIRE(pred_dtype == B200SWIN_F32 || pred_dtype == B200SWIN_BF16, "silog_bwd: bad dtype %d", pred_dtype);
  if (n == 0) return B200SWIN_OK;
  int grid = silog_grid(n);
  cudaStream_t st = (cudaStream_t)stream;
  if (pred_dtype == B200SWIN_F32)
    silog_bwd_kernel<float><<<grid, kSilogThreads, 0, st>>>((const float*)pred, target, n, lambd, stats, grad_out,
                                                          (float*)grad_pred);
  else
    silog_bwd_kernel<__nv_bfloat16><<<grid, kSilogThreads, 0, st>>>((const __nv_bfloat16*)pred, target, n, lambd,
                                                                  stats, grad_out, (__nv_bfloat16*)grad_pred);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
