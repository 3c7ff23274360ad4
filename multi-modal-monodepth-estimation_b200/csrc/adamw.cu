// Fused multi-tensor AdamW over ONE flat parameter buffer with per-tensor learning-rate scale and weight decay.
//
// Replaces the optimizer step the reference builds with SwinLayerDecayOptimizerConstructor (models/optimizer.py:36-104:
// ~60 parameter groups = layer-decay depth x {decay, no_decay}) and steps with torch.optim.AdamW after rewriting every
// group's lr (train.py:195-203): here all parameters, gradients and both moments live in flat fp32 buffers, every tensor
// occupies a whole number of 1024-element chunks, and ONE launch walks the chunks.  A chunk looks up its tensor's
// lr_scale / weight_decay; the step's base learning rate and the step count are read from device memory, so the launch
// is CUDA-graph capturable and a schedule only rewrites one float.  The same pass writes the bf16 copy of the updated
// weights that the tensor-core GEMMs read (no separate cast launches per weight and step).
//
// Math = torch.optim.AdamW (decoupled weight decay, bias correction), fp32:
//   p *= 1 - lr*wd;  m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: 16 B read + 12 B (+2 B bf16) written per element.
#include "common.cuh"
#include "../../include/b200swin.h"

namespace b200swin {
namespace {
constexpr int kChunk = 1024;

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             __nv_bfloat16* __restrict__ p16, const int* __restrict__ chunk_tensor, const float* __restrict__ lr_scale,
             const float* __restrict__ wd, const float* __restrict__ lr_ptr, const float* __restrict__ step_ptr,
             double beta1d, double beta2d, float eps, float grad_scale, int64_t nchunks) {
  // bias corrections in double like torch (1 - 0.999^t in fp32 loses 1e-5 relative to cancellation)
  const float lr0 = *lr_ptr;
  const double t = (double)*step_ptr;
  const double bc1 = 1.0 - pow(beta1d, t);
  const float bc2s = (float)sqrt(1.0 - pow(beta2d, t));
  const float w1 = (float)(1.0 - beta1d), beta2 = (float)beta2d, w2 = (float)(1.0 - beta2d);
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int ti = chunk_tensor[c];
    if (ti < 0) continue;                                   // padding chunk
    const float lr = lr0 * lr_scale[ti];
    const float decay = 1.0f - lr * wd[ti];
    const float step_size = (float)((double)lr / bc1);
    const int64_t i = c * kChunk + threadIdx.x * 4;
    float pv[4], gv[4], mv[4], vv[4];
    ld4(p + i, pv); ld4(g + i, gv); ld4(m + i, mv); ld4(v + i, vv);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float ge = gv[e] * grad_scale;
      pv[e] *= decay;
      mv[e] = mv[e] + (ge - mv[e]) * w1;                    // lerp, as torch does
      vv[e] = vv[e] * beta2 + w2 * ge * ge;
      const float denom = sqrtf(vv[e]) / bc2s + eps;
      pv[e] -= step_size * (mv[e] / denom);
    }
    st4(p + i, pv); st4(m + i, mv); st4(v + i, vv);
    if (p16) st4(p16 + i, pv);
  }
}
}  // namespace
}  // namespace b200swin

using namespace b200swin;

extern "C" int b200swin_adamw_chunk(void) { return kChunk; }

extern "C" int b200swin_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                   void* params_bf16, const int* chunk_tensor, const float* lr_scale,
                                   const float* weight_decay, const float* lr, const float* step, double beta1,
                                   double beta2, float eps, float grad_scale, int64_t nchunks, void* stream) {
  BSW_REQUIRE(params && grads && exp_avg && exp_avg_sq && chunk_tensor && lr_scale && weight_decay && lr && step,
              "adamw_step: null pointer");
  BSW_REQUIRE(nchunks >= 0, "adamw_step: negative chunk count");
  BSW_REQUIRE(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) |
                reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(params_bf16) & 7) == 0,
              "adamw_step: buffers must be 16-byte aligned");
  BSW_REQUIRE(beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.f, "adamw_step: bad hyper-parameters");
  if (nchunks == 0) return B200SWIN_OK;
  int64_t grid = (int64_t)sm_count() * 8;
  if (grid > nchunks) grid = nchunks;
  adamw_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, (__nv_bfloat16*)params_bf16, chunk_tensor, lr_scale, weight_decay, lr, step,
      beta1, beta2, eps, grad_scale, nchunks);
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}
