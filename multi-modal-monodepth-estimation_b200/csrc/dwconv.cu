// Depthwise 3x3 convolution on the token layout [B, H, W, C] (stride 1, zero padding 1, no bias): the `conv_proj` of the
// reference's ConvMlp (models/swin_transformer_v2.py:98-104, :107-111), which permutes [B, L, C] to NCHW, calls cuDNN and
// permutes back.  Here the tensor stays where it is:
//   y[b,i,j,c]  = sum_{u,v} w[c,u,v] * x[b, i+u-1, j+v-1, c]
//   dx[b,i,j,c] = sum_{u,v} w[c,u,v] * dy[b, i-u+1, j-v+1, c]          (same kernel, taps flipped)
//   dw[c,u,v]   = sum_{b,i,j} dy[b,i,j,c] * x[b, i+u-1, j+v-1, c]      (per-CTA partials + fixed-order reduce)
// HBM-bound: 9 reads per output hit L1/L2 (neighbouring tokens), algorithmic traffic = read x + write y.
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {
namespace {

template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
                                                        int B, int H, int W, int C, int flip) {
  const int G = C / 4;
  const int64_t total = (int64_t)B * H * W * G;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(idx % G);
    const int64_t tok = idx / G;
    const int j = (int)(tok % W);
    const int i = (int)((tok / W) % H);
    const int64_t b = tok / ((int64_t)W * H);
    const int c = cg * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int ii = i + u - 1;
      if (ii < 0 || ii >= H) continue;
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const int jj = j + v - 1;
        if (jj < 0 || jj >= W) continue;
        const int tap = flip ? (2 - u) * 3 + (2 - v) : u * 3 + v;
        float xv[4];
        ld4(x + ((b * H + ii) * W + jj) * C + c, xv);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(w[(c + e) * 9 + tap], xv[e], acc[e]);
      }
    }
    st4(y + tok * C + c, acc);
  }
}

// blockDim = 256 = 64 channel groups (of 4) x 4 token lanes; grid = (token chunks, ceil(C / 256))
template <typename T>
__global__ void __launch_bounds__(256) dwconv3x3_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                              float* __restrict__ partial, int B, int H, int W, int C) {
  __shared__ float red[4][64][36];
  const int cgl = threadIdx.x & 63, lane = threadIdx.x >> 6;
  const int c = (blockIdx.y * 64 + cgl) * 4;
  const int64_t ntok = (int64_t)B * H * W;
  const int64_t per = (ntok + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = blockIdx.x * per, t1 = min(ntok, t0 + per);
  float acc[4][9];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[e][t] = 0.f;
  if (c < C) {
    for (int64_t tok = t0 + lane; tok < t1; tok += 4) {
      const int j = (int)(tok % W);
      const int i = (int)((tok / W) % H);
      const int64_t b = tok / ((int64_t)W * H);
      float g[4];
      ld4(dy + tok * C + c, g);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int ii = i + u - 1;
        if (ii < 0 || ii >= H) continue;
#pragma unroll
        for (int v = 0; v < 3; ++v) {
          const int jj = j + v - 1;
          if (jj < 0 || jj >= W) continue;
          float xv[4];
          ld4(x + ((b * H + ii) * W + jj) * C + c, xv);
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e][u * 3 + v] = fmaf(g[e], xv[e], acc[e][u * 3 + v]);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int t = 0; t < 9; ++t) red[lane][cgl][e * 9 + t] = acc[e][t];
  __syncthreads();
  // 64 x 36 sums over the four token lanes, fixed order
  for (int idx = threadIdx.x; idx < 64 * 36; idx += 256) {
    const int g = idx / 36, r = idx % 36;
    const int cc = (blockIdx.y * 64 + g) * 4 + r / 9;
    if (cc < C) partial[((int64_t)blockIdx.x * C + cc) * 9 + r % 9] = red[0][g][r] + red[1][g][r] + red[2][g][r] + red[3][g][r];
  }
}

__global__ void dwconv3x3_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nparts, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * n + idx];
  dw[idx] = s;
}

int wgrad_parts(int64_t ntok) {
  int64_t p = (ntok + 255) / 256;
  const int64_t cap = 2 * (int64_t)sm_count();
  return (int)(p < 1 ? 1 : (p > cap ? cap : p));
}

}  // namespace
}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_dwconv3x3(const void* x, const float* weight, void* y, int B, int H, int W, int C, int dtype,
                                  int transpose, void* stream) {
  BSW_REQUIRE(x && weight && y, "dwconv3x3: null pointer");
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "dwconv3x3: bad shape (C must be a multiple of 4)");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "dwconv3x3: bad dtype %d", dtype);
  const int64_t total = (int64_t)B * H * W * (C / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = 16 * (int64_t)sm_count();
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200SWIN_F32)
    dwconv3x3_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, weight, (float*)y, B, H, W, C, transpose ? 1 : 0);
  else
    dwconv3x3_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, weight, (__nv_bfloat16*)y, B, H, W,
                                                                      C, transpose ? 1 : 0);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

extern "C" size_t b200swin_dwconv3x3_wgrad_workspace_bytes(int B, int H, int W, int C) {
  return (size_t)wgrad_parts((int64_t)B * H * W) * (size_t)C * 9 * sizeof(float);
}

extern "C" int b200swin_dwconv3x3_wgrad(const void* x, const void* dy, float* dweight, int B, int H, int W, int C, int dtype,
                                        void* workspace, size_t workspace_bytes, void* stream) {
  BSW_REQUIRE(x && dy && dweight, "dwconv3x3_wgrad: null pointer");
  BSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "dwconv3x3_wgrad: bad shape (C must be a multiple of 4)");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "dwconv3x3_wgrad: bad dtype %d", dtype);
  BSW_REQUIRE(workspace && workspace_bytes >= b200swin_dwconv3x3_wgrad_workspace_bytes(B, H, W, C),
              "dwconv3x3_wgrad: workspace too small (see b200swin_dwconv3x3_wgrad_workspace_bytes)");
  const int parts = wgrad_parts((int64_t)B * H * W);
  dim3 grid(parts, (C + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200SWIN_F32)
    dwconv3x3_wgrad_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)dy, (float*)workspace, B, H, W, C);
  else
    dwconv3x3_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy,
                                                                (float*)workspace, B, H, W, C);
  BSW_LAUNCH_CHECK();
  dwconv3x3_wgrad_reduce_kernel<<<(C * 9 + 255) / 256, 256, 0, st>>>((const float*)workspace, dweight, parts, C * 9);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
