// Continuous position bias table of Swin-V2 (cpb_mlp / rpe_mlp), forward + backward.
// Replaces WindowAttention.forward's bias-table part, models/swin_transformer_v2.py:304-313 with
// rpe_output_type='sigmoid':
//     hid   = relu(coords @ W0^T + b0)          coords [T, 2] (log-spaced offsets, T = (2ws-1)^2), W0 [HID, 2], b0 [HID]
//     table = 16 * sigmoid(hid @ W2^T)          W2 [nH, HID] (no bias)  ->  table [T, nH] fp32
// A few hundred kFLOP per block -- in PyTorch that is ~5 launches forward and ~10 backward per block, all of them
// latency-bound; here it is one launch each way.  fp32 throughout (the reference runs this branch in fp32, :50-56).
//   fwd: one CTA per table row t: the hidden vector goes to shared memory, then one warp per group of heads.
//   bwd: one warp per hidden unit k (8 per CTA sharing the staged dz tiles): lanes stride over t, recompute hid[t,k] and
//        accumulate dW2[:,k], dW0[k,:], db0[k] in registers; fixed-order warp reduction (deterministic, no atomics).
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {

constexpr int kCpbThreads = 128;
constexpr int kCpbMaxHeads = 64;
constexpr float kLogitMax = 4.605170185988092f;   // ln(1 / 0.01)

__global__ void __launch_bounds__(kCpbThreads)
cpb_fwd_kernel(const float* __restrict__ coords, const float* __restrict__ w0, const float* __restrict__ b0,
               const float* __restrict__ w2, float* __restrict__ table, const float* __restrict__ logit_scale,
               float* __restrict__ scale, int T, int HID, int nH) {
  extern __shared__ float hid[];   // [HID]
  const int t = blockIdx.x;
  // per-head temperature of the cosine attention: exp(min(logit_scale, ln 100))   (swin_transformer_v2.py:294)
  if (t == 0 && logit_scale && threadIdx.x < nH) scale[threadIdx.x] = __expf(fminf(logit_scale[threadIdx.x], kLogitMax));
  const float c0 = coords[2 * t], c1 = coords[2 * t + 1];
  for (int k = threadIdx.x; k < HID; k += kCpbThreads)
    hid[k] = fmaxf(fmaf(c0, w0[2 * k], fmaf(c1, w0[2 * k + 1], b0[k])), 0.f);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int h = warp; h < nH; h += kCpbThreads / 32) {
    float s = 0.f;
#pragma unroll 8
    for (int k = lane; k < HID; k += 32) s = fmaf(hid[k], w2[(int64_t)h * HID + k], s);   // 8 loads in flight: pure latency
    s = warp_sum(s);
    if (lane == 0) table[(int64_t)t * nH + h] = 16.0f / (1.0f + __expf(-s));
  }
}

// dz[t,h] = dtable[t,h] * y (1 - y/16),  y = table[t,h]   (16 sigmoid' = y (1 - y/16)).
// One WARP per hidden unit k, kCpbUnits warps per CTA: the CTA stages `tile` table rows of dz TRANSPOSED in shared memory
// ([h][t], odd row stride: coalesced global reads, conflict-free compute reads) together with their coordinates, and
// every warp reuses them, lanes striding over the rows.  The host sizes the tile to cover the whole table whenever it
// fits 96 KB (every Swin-V2 configuration: 529 x 32 heads = 68 KB), so a launch is ONE round of global loads (16-byte,
// all issued before the first use), one barrier, arithmetic out of shared memory and the reductions: the kernel is pure
// latency, and the former 128-row tiles cost five dependent rounds (37 us per launch cold, x 24 blocks per step).
// NH = compile-time bucket of the head count (nH <= NH): the per-row loops over heads are fully unrolled with no branch
// -- rows of dz and columns of W2 beyond nH are zero -- because 64 uniform `if (h < nH)` branches per table row, with
// two warps per scheduler to hide them, WERE the kernel's time (25 us per launch whatever nH).
constexpr int kCpbUnits = 8;
constexpr int kCpbBwdSmemFloats = 24 * 1024;   // 96 KB
template <int NH>
__global__ void __launch_bounds__(kCpbUnits * 32)
cpb_bwd_kernel(const float* __restrict__ coords, const float* __restrict__ w0, const float* __restrict__ b0,
               const float* __restrict__ w2, const float* __restrict__ table, const float* __restrict__ dtable,
               float* __restrict__ dw0, float* __restrict__ db0, float* __restrict__ dw2,
               const float* __restrict__ logit_scale, const float* __restrict__ dscale, float* __restrict__ dlogit, int T,
               int HID, int nH, int tile) {
  extern __shared__ float sdz[];     // [NH][tile + 1] then coordinates [tile][2]
  const int stride = tile + 1;
  float* scoord = sdz + NH * stride;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // d logit_scale = d scale * scale where the clamp is inactive (torch.clamp passes the gradient for x <= max)
  if (blockIdx.x == 0 && logit_scale && threadIdx.x < nH) {
    const float ls = logit_scale[threadIdx.x];
    dlogit[threadIdx.x] = ls <= kLogitMax ? dscale[threadIdx.x] * __expf(ls) : 0.f;
  }
  const int k = blockIdx.x * kCpbUnits + warp;
  const bool kvalid = k < HID;
  const float wa = kvalid ? w0[2 * k] : 0.f, wb = kvalid ? w0[2 * k + 1] : 0.f, bb = kvalid ? b0[k] : 0.f;
  float acc[NH];                     // dW2[h, k] partial sums (compile-time indexed below)
  float w2k[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) { acc[h] = 0.f; w2k[h] = (kvalid && h < nH) ? w2[(int64_t)h * HID + k] : 0.f; }
  float g0 = 0.f, g1 = 0.f, gb = 0.f;
  for (int e = nH * stride + threadIdx.x; e < NH * stride; e += kCpbUnits * 32) sdz[e] = 0.f;   // heads beyond nH
  const bool vec = (nH & 3) == 0 && (tile & 3) == 0 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(dtable)) & 15) == 0;
  for (int t0 = 0; t0 < T; t0 += tile) {
    __syncthreads();
    const int nrow = min(tile, T - t0);
    for (int e = threadIdx.x; e < 2 * nrow; e += kCpbUnits * 32) scoord[e] = coords[2 * t0 + e];
    if (vec) {                       // 4 heads of one table row per load
      const float4* y4 = reinterpret_cast<const float4*>(table + (int64_t)t0 * nH);
      const float4* d4 = reinterpret_cast<const float4*>(dtable + (int64_t)t0 * nH);
      const int n4 = nrow * nH / 4;
#pragma unroll 4
      for (int e4 = threadIdx.x; e4 < n4; e4 += kCpbUnits * 32) {
        const float4 y = y4[e4], d = d4[e4];
        const int e = 4 * e4, tl = e / nH, h = e - tl * nH;
        float* dst = sdz + h * stride + tl;
        dst[0] = d.x * y.x * (1.0f - y.x * 0.0625f);
        dst[stride] = d.y * y.y * (1.0f - y.y * 0.0625f);
        dst[2 * stride] = d.z * y.z * (1.0f - y.z * 0.0625f);
        dst[3 * stride] = d.w * y.w * (1.0f - y.w * 0.0625f);
      }
    } else {
#pragma unroll 4
      for (int e = threadIdx.x; e < nrow * nH; e += kCpbUnits * 32) {
        const int tl = e / nH, h = e - tl * nH;
        const float y = table[(int64_t)t0 * nH + e];
        sdz[h * stride + tl] = dtable[(int64_t)t0 * nH + e] * y * (1.0f - y * 0.0625f);
      }
    }
    __syncthreads();
    for (int tl = lane; tl < nrow; tl += 32) {
      const float c0 = scoord[2 * tl], c1 = scoord[2 * tl + 1];
      const float pre = fmaf(c0, wa, fmaf(c1, wb, bb));
      const float hv = fmaxf(pre, 0.f);
      float dh = 0.f;
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float dz = sdz[h * stride + tl];
        acc[h] = fmaf(dz, hv, acc[h]);
        dh = fmaf(dz, w2k[h], dh);
      }
      if (pre > 0.f) {
        g0 = fmaf(dh, c0, g0);
        g1 = fmaf(dh, c1, g1);
        gb += dh;
      }
    }
  }
  // fixed-order warp reductions (deterministic)
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const float s = warp_sum(acc[h]);
    if (lane == 0 && kvalid && h < nH) dw2[(int64_t)h * HID + k] = s;
  }
  g0 = warp_sum(g0); g1 = warp_sum(g1); gb = warp_sum(gb);
  if (lane == 0 && kvalid) { dw0[2 * k] = g0; dw0[2 * k + 1] = g1; db0[k] = gb; }
}

}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_cpb_fwd(const float* coords, const float* w0, const float* b0, const float* w2, float* table,
                                const float* logit_scale, float* scale, int T, int HID, int nH, void* stream) {
  BSW_REQUIRE(coords && w0 && b0 && w2 && table, "cpb_fwd: null pointer");
  BSW_REQUIRE((logit_scale == nullptr) == (scale == nullptr), "cpb_fwd: logit_scale and scale go together");
  BSW_REQUIRE(T > 0 && HID > 0 && HID <= 8192 && nH > 0 && nH <= kCpbMaxHeads, "cpb_fwd: T=%d HID=%d nH=%d out of range",
              T, HID, nH);
  cpb_fwd_kernel<<<T, kCpbThreads, (size_t)HID * sizeof(float), (cudaStream_t)stream>>>(coords, w0, b0, w2, table, logit_scale,
                                                                                     scale, T, HID, nH);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

extern "C" int b200swin_cpb_bwd(const float* coords, const float* w0, const float* b0, const float* w2,
                                const float* table, const float* dtable, float* dw0, float* db0, float* dw2,
                                const float* logit_scale, const float* dscale, float* dlogit, int T, int HID, int nH,
                                void* stream) {
  BSW_REQUIRE(coords && w0 && b0 && w2 && table && dtable && dw0 && db0 && dw2, "cpb_bwd: null pointer");
  BSW_REQUIRE((logit_scale == nullptr) == (dscale == nullptr) && (logit_scale == nullptr) == (dlogit == nullptr),
              "cpb_bwd: logit_scale, dscale and dlogit go together");
  BSW_REQUIRE(T > 0 && HID > 0 && HID <= 8192 && nH > 0 && nH <= kCpbMaxHeads, "cpb_bwd: T=%d HID=%d nH=%d out of range",
              T, HID, nH);
  // rows per tile: the whole table when nH * (tile + 1) + 2 * tile floats fit, else as many rows (a multiple of 4) as do
  const int NH = nH <= 4 ? 4 : nH <= 8 ? 8 : nH <= 16 ? 16 : nH <= 24 ? 24 : nH <= 32 ? 32 : nH <= 48 ? 48 : 64;
  int tile = (kCpbBwdSmemFloats - NH) / (NH + 2);
  tile = tile >= T ? (T + 3) / 4 * 4 : tile / 4 * 4;
  const size_t smem = ((size_t)NH * (tile + 1) + 2 * (size_t)tile) * sizeof(float);
#define CPB_BWD(N)                                                                                                    \
  case N: {                                                                                                           \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      BSW_CUDA(cudaFuncSetAttribute(cpb_bwd_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
                                    kCpbBwdSmemFloats * (int)sizeof(float)));                                         \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    cpb_bwd_kernel<N><<<(HID + kCpbUnits - 1) / kCpbUnits, kCpbUnits * 32, smem, (cudaStream_t)stream>>>(              \
        coords, w0, b0, w2, table, dtable, dw0, db0, dw2, logit_scale, dscale, dlogit, T, HID, nH, tile);            \
  } break
  switch (NH) {
    CPB_BWD(4); CPB_BWD(8); CPB_BWD(16); CPB_BWD(24); CPB_BWD(32); CPB_BWD(48);
    default: CPB_BWD(64);
  }
#undef CPB_BWD
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
