// Small HBM-bound helpers around the tensor-core kernels: fp32 -> bf16 casts (single or hi/lo split for the
// fp32-accurate GEMM mode), and column sums for bias gradients.
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

// dst_hi = bf16(x); dst_lo = bf16(x - float(dst_hi))  (dst_lo optional).  x ~ hi + lo to ~2^-17 relative.
__global__ void __launch_bounds__(256)
split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                  int64_t n4, int64_t n) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n4; i += nth) {
    float v[4], h[4], l[4];
    ld4(src + 4 * i, v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
      l[e] = v[e] - h[e];
    }
    st4(hi + 4 * i, h);
    if (lo) st4(lo + 4 * i, l);
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nth) {
    float v = src[i];
    __nv_bfloat16 hb = __float2bfloat16_rn(v);
    hi[i] = hb;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(hb));
  }
}

// out[n] = sum_m x[m, col_offset + n] over a [M, ld] matrix; two-stage, fixed order.
// A block covers 128 columns with 16-byte loads (8 bf16 / 4 fp32 per thread) and 256 / (128 / VEC) row lanes; the row
// loop is unrolled four times with independent accumulators so that 64 bytes per thread are in flight (HBM-bound).
template <typename T>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, int64_t M, int64_t ld, int64_t col0, int ncols,
                      float* __restrict__ part, int rows_per_block) {
  constexpr int VEC = 16 / sizeof(T);              // elements per 16-byte load
  constexpr int TPR = 128 / VEC;                   // threads per row
  constexpr int RL = 256 / TPR;                    // row lanes
  const int cq = threadIdx.x % TPR, rl = threadIdx.x / TPR;
  const int c = blockIdx.x * 128 + cq * VEC;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float acc[4][VEC];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[u][e] = 0.f;
  auto add = [&](int u, const uint4& w) {
    if constexpr (sizeof(T) == 2) {
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[e]));
        acc[u][2 * e] += f.x;
        acc[u][2 * e + 1] += f.y;
      }
    } else {
      acc[u][0] += __uint_as_float(w.x); acc[u][1] += __uint_as_float(w.y);
      acc[u][2] += __uint_as_float(w.z); acc[u][3] += __uint_as_float(w.w);
    }
  };
  if (c < ncols) {
    const T* base = x + col0 + c;
    int64_t r = r0 + rl;
    for (; r + 3 * RL < r1; r += 4 * RL) {
      uint4 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = *reinterpret_cast<const uint4*>(base + (r + u * RL) * ld);
#pragma unroll
      for (int u = 0; u < 4; ++u) add(u, w[u]);
    }
    for (; r < r1; r += RL) add(0, *reinterpret_cast<const uint4*>(base + r * ld));
  }
  __shared__ float sh[RL][128 + 4];
#pragma unroll
  for (int e = 0; e < VEC; ++e) sh[rl][cq * VEC + e] = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]);
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < RL; ++w) s += sh[w][threadIdx.x];
    int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < ncols) part[(int64_t)blockIdx.y * ncols + cc] = s;
  }
}

// fixed-order sum of the row-block partials: 32 columns x 8 part lanes per block
__global__ void __launch_bounds__(256)
colsum_final_kernel(const float* __restrict__ part, int nparts, int ncols, const float* __restrict__ extra,
                    float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s = 0.f;
  if (c < ncols)
    for (int p = py; p < nparts; p += 8) s += part[(int64_t)p * ncols + c];
  sh[py][cx] = s;
  __syncthreads();
  if (py == 0 && c < ncols) {
    float t = extra ? extra[c] : 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[w][cx];
    out[c] = t;
  }
}

static int colsum_rows_per_block(int64_t M, int ncols) {
  int64_t colblocks = (ncols + 127) / 128;
  int64_t want = ((int64_t)sm_count() * 4 + colblocks - 1) / colblocks;   // row blocks to fill the machine
  int64_t rpb = (M + want - 1) / want;
  if (rpb < 64) rpb = 64;
  return (int)rpb;
}

}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_split_bf16(const float* src, void* hi, void* lo, int64_t n, void* stream) {
  BSW_REQUIRE(src && hi, "split_bf16: null pointer");
  if (n <= 0) return B200SWIN_OK;
  bool al = ((reinterpret_cast<uintptr_t>(src) & 15) | (reinterpret_cast<uintptr_t>(hi) & 7) |
             (reinterpret_cast<uintptr_t>(lo) & 7)) == 0;
  int64_t n4 = al ? n / 4 : 0;
  int64_t blocks = (n / 4 + 255) / 256 + 1;
  int64_t cap = (int64_t)sm_count() * 8;
  split_bf16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
      src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n4, n);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

extern "C" size_t b200swin_colsum_workspace_bytes(int64_t M, int ncols) {
  int rpb = colsum_rows_per_block(M, ncols);
  int64_t nparts = (M + rpb - 1) / rpb;
  return (size_t)nparts * (size_t)ncols * sizeof(float);
}

extern "C" int b200swin_colsum(const void* x, int dtype, int64_t M, int64_t ld, int64_t col0, int ncols,
                               const float* extra, float* out, void* workspace, size_t workspace_bytes,
                               void* stream) {
  BSW_REQUIRE(x && out && workspace, "colsum: null pointer");
  const int vec = dtype == B200SWIN_BF16 ? 8 : 4;      // 16-byte loads
  BSW_REQUIRE(M > 0 && ncols > 0 && ncols % vec == 0 && col0 % vec == 0 && ld % vec == 0 && col0 + ncols <= ld,
              "colsum: columns / leading dimension must be multiples of %d (16-byte loads)", vec);
  BSW_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "colsum: x must be 16-byte aligned");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "colsum: bad dtype %d", dtype);
  int rpb = colsum_rows_per_block(M, ncols);
  int nparts = (int)((M + rpb - 1) / rpb);
  BSW_REQUIRE(workspace_bytes >= (size_t)nparts * ncols * sizeof(float), "colsum: workspace too small");
  BSW_REQUIRE(nparts <= 65535, "colsum: too many row blocks");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((ncols + 127) / 128, nparts);
  if (dtype == B200SWIN_F32)
    colsum_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)x, M, ld, col0, ncols, (float*)workspace, rpb);
  else
    colsum_partial_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, M, ld, col0, ncols,
                                                              (float*)workspace, rpb);
  BSW_LAUNCH_CHECK();
  colsum_final_kernel<<<(ncols + 31) / 32, 256, 0, st>>>((const float*)workspace, nparts, ncols, extra, out);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
