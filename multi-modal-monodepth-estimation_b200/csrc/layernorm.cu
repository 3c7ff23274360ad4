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
      const float sc = row_scale ? row_scale[(uint32_t)row / (uint32_t)rows_per_scale] : 1.0f;
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

// Post-norm transformer layer:  y = LN(x + xadd) * gamma + beta   (Transformer_Encoder.forward,
// models/cnn_transformer.py:202-203 and :208-209: `x = v + attn(x); x = norm1(x)` / `x = x + ffn(x); x = norm2(x)`).
// x is the fp32 stream, xadd the branch output (fp32 or bf16: the GEMM's output type under autocast).  One pass instead of
// an elementwise add, a LayerNorm and the cast of the result for the next GEMM: y (fp32) and, optionally, y16 = bf16(y)
// and xsum = x + xadd (what the backward normalises; written only when a backward will run).
template <typename A, int NV, int R>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_sum_kernel(const float* __restrict__ x, const A* __restrict__ xadd, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float* __restrict__ y, __nv_bfloat16* __restrict__ y16,
                  float* __restrict__ xsum, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int C,
                  float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kLnWarps;
  const float inv_c = 1.0f / (float)C;
  for (int64_t row0 = warp0 * R; row0 < rows; row0 += nwarps * R) {
    float v[R][NV][4];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + k * 128;
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 4; ++e) v[r][k][e] = 0.f;
        if (row < rows && c < C) {
          ld4(x + row * C + c, v[r][k]);
          ld4(xadd + row * C + c, a);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) v[r][k][e] += a[e];
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
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = lane * 4 + k * 128;
        if (c < C) {
          float g[4], b[4], o[4];
          ld4(gamma + c, g);
          ld4(beta + c, b);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = (v[r][k][e] - mean) * rstd * g[e] + b[e];
          st4(y + row * C + c, o);
          if (y16) st4(y16 + row * C + c, o);
          if (xsum) st4(xsum + row * C + c, v[r][k]);
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
      const float sc = row_scale ? row_scale[(uint32_t)row / (uint32_t)rows_per_scale] : 1.0f;
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
// bf16 streaming kernels (C % 8 == 0, C <= 1536).  HBM-bound, so the design goal is bytes in flight, not math:
//   * every warp owns a private ring of kRing slots in shared memory; a slot receives one CHUNK of PASSES * (32 / LPR)
//     consecutive rows of both input tensors -- rows are contiguous in memory, so a chunk is ONE 1-D bulk copy per
//     tensor (cp.async.bulk, 1.5 - 3 KB), issued by lane 0 and signalled through an mbarrier;
//   * while a slot is processed the other kRing - 1 slots of the warp are in flight: 16 warps x 2 slots x ~4 KB =
//     128 KB per SM without a single register holding in-flight data;
//   * the math reads the slot with conflict-free 16-byte LDS (8 bf16 per lane), keeps statistics in fp32 and writes
//     the result straight to global memory (16 bytes per lane, coalesced).
// LPR = lanes per row (16 for C <= 128: two rows per pass), NV = 16-byte chunks per lane and row.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kRing = 3;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void ring_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ring_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void ring_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(a), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) break;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(w[e] << 16);
    f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
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
// the same eight values as four (even, odd) pairs for the packed fp32x2 instructions (FFMA2 / FMUL2 / FADD2)
__device__ __forceinline__ void unpack8p(const uint4& u, float2 (&f)[4]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) f[e] = make_float2(__uint_as_float(w[e] << 16), __uint_as_float(w[e] & 0xffff0000u));
}
__device__ __forceinline__ uint4 pack8p(const float2 (&f)[4]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[e].x, f[e].y);
    w[e] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void lds_f8p(const float* p, float2 (&f)[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = make_float2(a.x, a.y); f[1] = make_float2(a.z, a.w); f[2] = make_float2(b.x, b.y); f[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void lds_f8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Ring bookkeeping shared by the forward and the backward kernel.  T0 / T1: the two streamed tensors (T1 may be null).
struct RingWarp {
  unsigned char* base;       // this warp's kRing slots, each (1 + t1_mul) * slot_bytes
  uint64_t* bars;            // this warp's kRing mbarriers
  uint32_t slot_bytes;       // bytes of the FIRST tensor's chunk (full chunk, bf16)
  int64_t rows, chunk_rows, nchunks, first, stride;
  int C;
  int t1_mul = 1;            // element size of the second tensor in units of bf16 (2: an fp32 residual stream)
  const __nv_bfloat16* t0;
  const void* t1;

  __device__ __forceinline__ int64_t my_chunks() const { return first < nchunks ? (nchunks - first + stride - 1) / stride : 0; }
  __device__ __forceinline__ size_t slot_stride() const { return (size_t)(1 + t1_mul) * slot_bytes; }
  __device__ __forceinline__ void issue(int64_t it, int lane) const {
    if (lane != 0) return;
    const int s = (int)(it % kRing);
    const int64_t row0 = (first + it * stride) * chunk_rows;
    const int64_t nr = min(chunk_rows, rows - row0);
    const uint32_t bytes = (uint32_t)(nr * C * 2);
    unsigned char* dst = base + (size_t)s * slot_stride();
    ring_expect(&bars[s], t1 ? (1 + t1_mul) * bytes : bytes);
    ring_bulk_load(dst, t0 + row0 * C, bytes, &bars[s]);
    if (t1) ring_bulk_load(dst + slot_bytes, static_cast<const unsigned char*>(t1) + (size_t)row0 * C * 2 * t1_mul, bytes * t1_mul, &bars[s]);
  }
};

// STREAM32: the residual stream is kept in fp32 beside its bf16 copy (what torch.autocast does: LayerNorm outputs and the
// residual adds stay fp32) -- `residual` is then an fp32 tensor and the result is written twice, fp32 to y32 (the stream)
// and bf16 to y (the operand of the next GEMM).  With the reference's 1e-5 block-norm initialisation a block's
// contribution is far below a bf16 half-ulp of the stream and would otherwise be rounded away in every block.
template <int NV, int LPR, int PASSES, bool STREAM32>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, const void* __restrict__ residual,
                   const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ row_scale, int64_t rows_per_scale, __nv_bfloat16* __restrict__ y,
                   float* __restrict__ y32, float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows,
                   int C, float eps) {
  extern __shared__ __align__(128) unsigned char smem_ln[];
  __shared__ __align__(8) uint64_t bars[kLnWarps][kRing];
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lr = lane % LPR, sub = lane / LPR;
  const int chunks16 = C >> 3;
  const uint32_t slot_bytes = (uint32_t)(PASSES * RPW) * C * 2;
  float* gb = reinterpret_cast<float*>(smem_ln);                     // gamma[C] | beta[C]
  unsigned char* ring0 = smem_ln + (((size_t)2 * C * 4 + 127) & ~(size_t)127);
  for (int c = threadIdx.x; c < C; c += kLnThreads) { gb[c] = gamma[c]; gb[C + c] = beta[c]; }
  if (lane == 0)
    for (int s = 0; s < kRing; ++s) ring_mbar_init(&bars[warp][s]);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  RingWarp rw;
  rw.t1_mul = STREAM32 ? 2 : 1;
  rw.slot_bytes = slot_bytes;
  rw.base = ring0 + (size_t)warp * kRing * rw.slot_stride();
  rw.bars = bars[warp];
  rw.rows = rows; rw.chunk_rows = PASSES * RPW; rw.C = C;
  rw.nchunks = (rows + rw.chunk_rows - 1) / rw.chunk_rows;
  rw.first = (int64_t)blockIdx.x * kLnWarps + warp;
  rw.stride = (int64_t)gridDim.x * kLnWarps;
  rw.t0 = x; rw.t1 = residual;
  const int64_t n = rw.my_chunks();
  const float inv_c = 1.0f / (float)C;
  for (int64_t it = 0; it < n && it < kRing; ++it) rw.issue(it, lane);

  for (int64_t it = 0; it < n; ++it) {
    const int s = (int)(it % kRing);
    ring_wait(&rw.bars[s], (uint32_t)((it / kRing) & 1));
    const unsigned char* xs = rw.base + (size_t)s * rw.slot_stride();
    const unsigned char* rs = xs + slot_bytes;
    const int64_t row0 = (rw.first + it * rw.stride) * rw.chunk_rows;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int rl = p * RPW + sub;                                  // row inside the chunk
      const int64_t row = row0 + rl;
      if (row0 + p * RPW >= rows) break;                             // warp-uniform
      const bool live = row < rows;
      // packed fp32x2 math (pairs of even / odd columns): the kernel is issue-bound once the loads are decoupled
      float2 f[NV][4];
      float2 sm2 = f2(0.f);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (live && ch < chunks16) v = *reinterpret_cast<const uint4*>(xs + (size_t)rl * C * 2 + ch * 16);
        unpack8p(v, f[k]);
#pragma unroll
        for (int e = 0; e < 4; ++e) sm2 = __fadd2_rn(sm2, f[k][e]);
      }
      const float mean = row_sum<LPR>(sm2.x + sm2.y) * inv_c;
      const float2 nmean = f2(-mean);
      float2 q2 = f2(0.f);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (k * LPR + lr < chunks16) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 d = __fadd2_rn(f[k][e], nmean); q2 = __ffma2_rn(d, d, q2); }
        }
      }
      const float rstd = rsqrtf(row_sum<LPR>(q2.x + q2.y) * inv_c + eps);
      const float sc = (row_scale && live) ? row_scale[(uint32_t)row / (uint32_t)rows_per_scale] : 1.0f;
      const float a = rstd * sc, b0 = -mean * a;    // ((x - mean) rstd g + b) sc + r = (x a + b0) g + (b sc + r)
      const float2 a2 = f2(a), b02 = f2(b0), sc2 = f2(sc);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        if (live && ch < chunks16) {
          float2 r[4], g[4], b[4], o[4];
          if (residual) {
            if constexpr (STREAM32) lds_f8p(reinterpret_cast<const float*>(rs + (size_t)rl * C * 4) + ch * 8, r);
            else unpack8p(*reinterpret_cast<const uint4*>(rs + (size_t)rl * C * 2 + ch * 16), r);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) r[e] = f2(0.f);
          }
          lds_f8p(gb + ch * 8, g);
          lds_f8p(gb + C + ch * 8, b);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = __ffma2_rn(__ffma2_rn(f[k][e], a2, b02), g[e], __ffma2_rn(b[e], sc2, r[e]));
          *reinterpret_cast<uint4*>(y + row * C + ch * 8) = pack8p(o);
          if constexpr (STREAM32) {
            float4* d32 = reinterpret_cast<float4*>(y32 + row * C + ch * 8);
            d32[0] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
            d32[1] = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
          }
        }
      }
      if (lr == 0 && live) { mean_out[row] = mean; rstd_out[row] = rstd; }
    }
    __syncwarp();                                                    // every lane has finished reading the slot
    if (it + kRing < n) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      rw.issue(it + kRing, lane);
    }
  }
}

// Backward.  Phase A accumulates dgamma / dbeta and the two row sums, phase B re-reads the slot and writes dx.
// COLSUM adds the column sums of dx -- the bias gradient of the Linear that produced x (proj / fc2) -- to the partial
// rows.
template <int NV, int LPR, int PASSES, bool COLSUM>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                   const float* __restrict__ gamma, const float* __restrict__ mean_in,
                   const float* __restrict__ rstd_in, const float* __restrict__ row_scale, int64_t rows_per_scale,
                   __nv_bfloat16* __restrict__ dx, float* __restrict__ part /*[grid][NP][C]*/, int64_t rows, int C) {
  constexpr int NP = COLSUM ? 3 : 2;
  extern __shared__ __align__(128) unsigned char smem_ln[];
  __shared__ __align__(8) uint64_t bars[kLnWarps][kRing];
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lr = lane % LPR, sub = lane / LPR;
  const int chunks16 = C >> 3;
  const uint32_t slot_bytes = (uint32_t)(PASSES * RPW) * C * 2;
  float* gm_s = reinterpret_cast<float*>(smem_ln);                   // gamma[C] | reduction scratch [NP][C]
  float* red = gm_s + C;
  unsigned char* ring0 = smem_ln + (((size_t)(1 + NP) * C * 4 + 127) & ~(size_t)127);
  for (int c = threadIdx.x; c < C; c += kLnThreads) gm_s[c] = gamma[c];
  for (int c = threadIdx.x; c < NP * C; c += kLnThreads) red[c] = 0.f;
  if (lane == 0)
    for (int s = 0; s < kRing; ++s) ring_mbar_init(&bars[warp][s]);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  RingWarp rw;
  rw.slot_bytes = slot_bytes;
  rw.base = ring0 + (size_t)warp * kRing * rw.slot_stride();
  rw.bars = bars[warp];
  rw.rows = rows; rw.chunk_rows = PASSES * RPW; rw.C = C;
  rw.nchunks = (rows + rw.chunk_rows - 1) / rw.chunk_rows;
  rw.first = (int64_t)blockIdx.x * kLnWarps + warp;
  rw.stride = (int64_t)gridDim.x * kLnWarps;
  rw.t0 = x; rw.t1 = dy;
  const int64_t n = rw.my_chunks();
  const float inv_c = 1.0f / (float)C;
  float2 dg[NV][4], db[NV][4], dc[COLSUM ? NV : 1][4];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) { dg[k][e] = f2(0.f); db[k][e] = f2(0.f); if (COLSUM) dc[k][e] = f2(0.f); }
  for (int64_t it = 0; it < n && it < kRing; ++it) rw.issue(it, lane);

  for (int64_t it = 0; it < n; ++it) {
    const int s = (int)(it % kRing);
    const int64_t row0 = (rw.first + it * rw.stride) * rw.chunk_rows;
    // per-row statistics: requested before the wait, they arrive with the slot
    float mean[PASSES], rstd[PASSES];
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int64_t row = row0 + p * RPW + sub;
      mean[p] = row < rows ? mean_in[row] : 0.f;
      rstd[p] = row < rows ? rstd_in[row] : 0.f;
    }
    ring_wait(&rw.bars[s], (uint32_t)((it / kRing) & 1));
    const unsigned char* xs = rw.base + (size_t)s * rw.slot_stride();
    const unsigned char* gs = xs + slot_bytes;
#pragma unroll
    for (int p = 0; p < PASSES; ++p) {
      const int rl = p * RPW + sub;
      const int64_t row = row0 + rl;
      if (row0 + p * RPW >= rows) break;                             // warp-uniform
      const bool live = row < rows;
      const float sc = (row_scale && live) ? row_scale[(uint32_t)row / (uint32_t)rows_per_scale] : 1.0f;
      const float a = rstd[p], b0 = -mean[p] * rstd[p];             // x_hat = x a + b0
      // packed fp32x2 math (the kernel is issue-bound once the loads are decoupled): pairs (even, odd column)
      const float2 a2 = f2(a), b2 = f2(b0), sc2 = f2(sc);
      float2 s1 = f2(0.f), s2 = f2(0.f);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        if (live && ch < chunks16) {
          float2 xf[4], gf[4], gm[4];
          unpack8p(*reinterpret_cast<const uint4*>(xs + (size_t)rl * C * 2 + ch * 16), xf);
          unpack8p(*reinterpret_cast<const uint4*>(gs + (size_t)rl * C * 2 + ch * 16), gf);
          lds_f8p(gm_s + ch * 8, gm);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 xh = __ffma2_rn(xf[e], a2, b2);
            const float2 g = __fmul2_rn(gf[e], sc2);
            dg[k][e] = __ffma2_rn(g, xh, dg[k][e]);
            db[k][e] = __fadd2_rn(db[k][e], g);
            const float2 gg = __fmul2_rn(g, gm[e]);
            s1 = __fadd2_rn(s1, gg);
            s2 = __ffma2_rn(gg, xh, s2);
          }
        }
      }
      const float m1 = row_sum<LPR>(s1.x + s1.y) * inv_c, m2 = row_sum<LPR>(s2.x + s2.y) * inv_c;
      const float2 nm1 = f2(-m1), nm2 = f2(-m2);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int ch = k * LPR + lr;
        if (live && ch < chunks16) {
          float2 xf[4], gf[4], gm[4], o[4];
          unpack8p(*reinterpret_cast<const uint4*>(xs + (size_t)rl * C * 2 + ch * 16), xf);
          unpack8p(*reinterpret_cast<const uint4*>(gs + (size_t)rl * C * 2 + ch * 16), gf);
          lds_f8p(gm_s + ch * 8, gm);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 xh = __ffma2_rn(xf[e], a2, b2);
            const float2 gg = __fmul2_rn(__fmul2_rn(gf[e], sc2), gm[e]);
            o[e] = __fmul2_rn(a2, __fadd2_rn(__ffma2_rn(xh, nm2, gg), nm1));
            if (COLSUM) dc[k][e] = __fadd2_rn(dc[k][e], o[e]);
          }
          *reinterpret_cast<uint4*>(dx + row * C + ch * 8) = pack8p(o);
        }
      }
    }
    __syncwarp();
    if (it + kRing < n) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      rw.issue(it + kRing, lane);
    }
  }
  // two rows per pass: lanes l and l + 16 own the same columns
  if (LPR == 16) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dg[k][e].x += __shfl_xor_sync(0xffffffffu, dg[k][e].x, 16);
        dg[k][e].y += __shfl_xor_sync(0xffffffffu, dg[k][e].y, 16);
        db[k][e].x += __shfl_xor_sync(0xffffffffu, db[k][e].x, 16);
        db[k][e].y += __shfl_xor_sync(0xffffffffu, db[k][e].y, 16);
        if (COLSUM) {
          dc[k][e].x += __shfl_xor_sync(0xffffffffu, dc[k][e].x, 16);
          dc[k][e].y += __shfl_xor_sync(0xffffffffu, dc[k][e].y, 16);
        }
      }
  }
  // fixed-order cross-warp reduction -> one partial row set per block.  Every warp parks its sums in its own row of a
  // [warps][NP * C] matrix laid over the (now idle) ring, ONE barrier, then each thread adds the eight rows of its
  // columns in warp order.  (Eight warps taking turns on one row cost eight barriers at the end of a 40 us kernel.)
  __syncthreads();                                   // every warp has left its ring slots
  float* stage = reinterpret_cast<float*>(ring0);    // ring >= kLnWarps * 3 slots * 4 KB >= kLnWarps * NP * C * 4
  if (sub == 0) {
    float* mine = stage + (size_t)warp * NP * C;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int ch = k * LPR + lr;
      if (ch < chunks16) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          mine[ch * 8 + 2 * e] = dg[k][e].x;
          mine[ch * 8 + 2 * e + 1] = dg[k][e].y;
          mine[C + ch * 8 + 2 * e] = db[k][e].x;
          mine[C + ch * 8 + 2 * e + 1] = db[k][e].y;
          if (COLSUM) {
            mine[2 * C + ch * 8 + 2 * e] = dc[k][e].x;
            mine[2 * C + ch * 8 + 2 * e + 1] = dc[k][e].y;
          }
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < NP * C; c += kLnThreads) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) t += stage[(size_t)w * NP * C + c];
    part[(int64_t)blockIdx.x * NP * C + c] = t;
  }
}

// dgamma / dbeta (/ dcolsum) = fixed-order sum of the per-block partial rows: 32 columns x 32 part lanes per block, so a
// thread has ~10 independent loads in flight instead of a chain of ~40 (the kernel is pure latency: 6 us -> 3 us).
constexpr int kRedLanes = 32;
__global__ void __launch_bounds__(32 * kRedLanes)
ln_param_reduce3_kernel(const float* __restrict__ part, int nparts, int C, int NP, float* __restrict__ o0,
                        float* __restrict__ o1, float* __restrict__ o2) {
  __shared__ float sh[kRedLanes][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;              // over NP*C
  float s = 0.f;
  if (c < NP * C) {
#pragma unroll 8
    for (int p = py; p < nparts; p += kRedLanes) s += part[(int64_t)p * NP * C + c];
  }
  sh[py][cx] = s;
  __syncthreads();
  if (py == 0 && c < NP * C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kRedLanes; ++w) t += sh[w][cx];
    if (c < C) o0[c] = t; else if (c < 2 * C) o1[c - C] = t; else o2[c - 2 * C] = t;
  }
}

static bool ln_fast_ok(int C, int dtype) { return dtype == B200SWIN_BF16 && C % 8 == 0 && C <= 1536; }

// (NV, LPR, PASSES) by row width: chunks of 1.5 - 3 KB per tensor
#define LN_FAST_DISPATCH(C, X)                   \
  do {                                           \
    if ((C) <= 128) X(1, 16, 4);                 \
    else if ((C) <= 256) X(1, 32, 4);            \
    else if ((C) <= 512) X(2, 32, 2);            \
    else if ((C) <= 768) X(3, 32, 1);            \
    else if ((C) <= 1024) X(4, 32, 1);           \
    else X(6, 32, 1);                            \
  } while (0)

static int ln_chunk_rows(int C) { return C <= 128 ? 8 : C <= 256 ? 4 : C <= 512 ? 2 : 1; }
static size_t ln_ring_bytes(int C, int tensors = 2) { return (size_t)kLnWarps * kRing * tensors * (size_t)ln_chunk_rows(C) * C * 2; }
static size_t ln_fwd_smem(int C, bool stream32 = false) {
  return (((size_t)2 * C * 4 + 127) & ~(size_t)127) + ln_ring_bytes(C, stream32 ? 3 : 2);
}
static size_t ln_bwd_smem(int C, int NP) { return (((size_t)(1 + NP) * C * 4 + 127) & ~(size_t)127) + ln_ring_bytes(C); }
static int ln_blocks_per_sm(size_t smem) {
  const size_t avail = 227 * 1024;
  int b = (int)(avail / (smem + 1024));
  return b < 1 ? 1 : (b > 4 ? 4 : b);
}
// one resident wave; every warp gets at least two chunks when there is that much work
static int ln_fast_grid(int64_t rows, int C, size_t smem) {
  const int64_t nchunks = (rows + ln_chunk_rows(C) - 1) / ln_chunk_rows(C);
  int64_t blocks = (nchunks + 2 * kLnWarps - 1) / (2 * kLnWarps);
  const int64_t cap = (int64_t)sm_count() * ln_blocks_per_sm(smem);
  if (blocks < 1) blocks = 1;
  return (int)(blocks < cap ? blocks : cap);
}
static int ln_fast_bwd_grid(int64_t rows, int C) { return ln_fast_grid(rows, C, ln_bwd_smem(C, 3)); }

static int ln_fwd_fast(const void* x, const void* residual, const float* gamma, const float* beta,
                       const float* row_scale, int64_t rps, void* y, float* y32, float* mean, float* rstd, int64_t rows,
                       int C, float eps, cudaStream_t st) {
  const bool stream32 = y32 != nullptr;
  const size_t smem = ln_fwd_smem(C, stream32);
  const int grid = ln_fast_grid(rows, C, smem);
#define X(NV, LPR, PASSES)                                                                                           \
  do {                                                                                                               \
    if (stream32) {                                                                                                  \
      BSW_CUDA(cudaFuncSetAttribute(ln_fwd_bf16_kernel<NV, LPR, PASSES, true>,                                       \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
      ln_fwd_bf16_kernel<NV, LPR, PASSES, true><<<grid, kLnThreads, smem, st>>>(                                     \
          (const __nv_bfloat16*)x, residual, gamma, beta, row_scale, rps, (__nv_bfloat16*)y, y32, mean, rstd, rows,  \
          C, eps);                                                                                                   \
    } else {                                                                                                         \
      BSW_CUDA(cudaFuncSetAttribute(ln_fwd_bf16_kernel<NV, LPR, PASSES, false>,                                      \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
      ln_fwd_bf16_kernel<NV, LPR, PASSES, false><<<grid, kLnThreads, smem, st>>>(                                    \
          (const __nv_bfloat16*)x, residual, gamma, beta, row_scale, rps, (__nv_bfloat16*)y, nullptr, mean, rstd,    \
          rows, C, eps);                                                                                             \
    }                                                                                                                \
  } while (0)
  LN_FAST_DISPATCH(C, X);
#undef X
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

static int ln_bwd_fast(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                       const float* row_scale, int64_t rps, void* dx, float* part, int grid, bool colsum, int64_t rows,
                       int C, cudaStream_t st) {
  const size_t smem = ln_bwd_smem(C, colsum ? 3 : 2);
#define X(NV, LPR, PASSES)                                                                                           \
  do {                                                                                                               \
    if (colsum) {                                                                                                    \
      BSW_CUDA(cudaFuncSetAttribute(ln_bwd_bf16_kernel<NV, LPR, PASSES, true>,                                       \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
      ln_bwd_bf16_kernel<NV, LPR, PASSES, true><<<grid, kLnThreads, smem, st>>>(                                     \
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, mean, rstd, row_scale, rps, (__nv_bfloat16*)dx,  \
          part, rows, C);                                                                                            \
    } else {                                                                                                         \
      BSW_CUDA(cudaFuncSetAttribute(ln_bwd_bf16_kernel<NV, LPR, PASSES, false>,                                      \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
      ln_bwd_bf16_kernel<NV, LPR, PASSES, false><<<grid, kLnThreads, smem, st>>>(                                    \
          (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, gamma, mean, rstd, row_scale, rps, (__nv_bfloat16*)dx,  \
          part, rows, C);                                                                                            \
    }                                                                                                                \
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
  BSW_REQUIRE(!row_scale || (rows < (1ll << 32) && rows_per_scale < (1ll << 32)), "ln_fwd: row_scale needs rows < 2^32");
  BSW_REQUIRE(dtype == B200SWIN_F32 || dtype == B200SWIN_BF16, "ln_fwd: bad dtype %d", dtype);
  if (rows == 0) return B200SWIN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (ln_fast_ok(C, dtype))
    return ln_fwd_fast(x, residual, gamma, beta, row_scale, rows_per_scale, y, nullptr, mean, rstd, rows, C, eps, st);
  if (dtype == B200SWIN_F32)
    return ln_fwd_launch<float>(x, residual, gamma, beta, row_scale, rows_per_scale, y, mean, rstd, rows, C, eps, st);
  return ln_fwd_launch<__nv_bfloat16>(x, residual, gamma, beta, row_scale, rows_per_scale, y, mean, rstd, rows, C,
                                      eps, st);
}

template <typename A>
static int ln_fwd_sum_launch(const float* x, const void* xadd, const float* gamma, const float* beta, float* y, void* y16,
                             float* xsum, float* mean, float* rstd, int64_t rows, int C, float eps, cudaStream_t st) {
  const int grid = ln_grid(rows);
#define LN_SUM(NV, R)                                                                                                  \
  ln_fwd_sum_kernel<A, NV, R><<<grid, kLnThreads, 0, st>>>(x, (const A*)xadd, gamma, beta, y, (__nv_bfloat16*)y16, xsum, \
                                                           mean, rstd, rows, C, eps)
  if (C <= 128) LN_SUM(1, 4);
  else if (C <= 256) LN_SUM(2, 2);
  else if (C <= 512) LN_SUM(4, 2);
  else LN_SUM(8, 1);
#undef LN_SUM
  BSW_LAUNCH_CHECK();
  return B200SWIN_OK;
}

extern "C" int b200swin_ln_fwd_sum(const float* x, const void* xadd, int xadd_dtype, const float* gamma, const float* beta,
                                   float* y, void* y16, float* xsum, float* mean, float* rstd, int64_t rows, int C, float eps,
                                   void* stream) {
  BSW_REQUIRE(x && xadd && gamma && beta && y && mean && rstd, "ln_fwd_sum: null pointer");
  BSW_REQUIRE(rows >= 0 && C > 0 && C % 4 == 0 && C <= 1024, "ln_fwd_sum: C=%d must be a multiple of 4, <= 1024", C);
  BSW_REQUIRE(xadd_dtype == B200SWIN_F32 || xadd_dtype == B200SWIN_BF16, "ln_fwd_sum: bad dtype %d", xadd_dtype);
  if (rows == 0) return B200SWIN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (xadd_dtype == B200SWIN_F32)
    return ln_fwd_sum_launch<float>(x, xadd, gamma, beta, y, y16, xsum, mean, rstd, rows, C, eps, st);
  return ln_fwd_sum_launch<__nv_bfloat16>(x, xadd, gamma, beta, y, y16, xsum, mean, rstd, rows, C, eps, st);
}

extern "C" int b200swin_ln_fwd_stream32(const void* x, const float* residual32, const float* gamma, const float* beta,
                                        const float* row_scale, int64_t rows_per_scale, void* y, float* y32,
                                        float* mean, float* rstd, int64_t rows, int C, float eps, void* stream) {
  BSW_REQUIRE(x && gamma && beta && y && y32 && mean && rstd, "ln_fwd_stream32: null pointer");
  // three tensors per ring slot (bf16 x, fp32 residual): the per-warp rings of 1536-wide rows no longer fit in 227 KB
  BSW_REQUIRE(rows >= 0 && C > 0 && C % 8 == 0 && C <= 1024, "ln_fwd_stream32: needs C %% 8 == 0, C <= 1024 (C=%d)", C);
  BSW_REQUIRE(!row_scale || rows_per_scale > 0, "ln_fwd_stream32: rows_per_scale must be > 0 with row_scale");
  BSW_REQUIRE(!row_scale || (rows < (1ll << 32) && rows_per_scale < (1ll << 32)), "ln_fwd_stream32: row_scale needs rows < 2^32");
  if (rows == 0) return B200SWIN_OK;
  return ln_fwd_fast(x, residual32, gamma, beta, row_scale, rows_per_scale, y, y32, mean, rstd, rows, C, eps,
                     (cudaStream_t)stream);
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
  BSW_REQUIRE(!row_scale || (rows < (1ll << 32) && rows_per_scale < (1ll << 32)), "ln_bwd: row_scale needs rows < 2^32");
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
    ln_param_reduce3_kernel<<<(NP * C + 31) / 32, 32 * kRedLanes, 0, st>>>((const float*)workspace, grid, C, NP, dgamma, dbeta,
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
