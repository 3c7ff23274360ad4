// Probe: what the legacy warp-level tensor path (mma.sync m16n8k16 bf16 -> HMMA) sustains on B200, alone and mixed with
// the per-element work of a softmax (ex2, fma, cvt), plus ldmatrix / movmatrix rates.  Window attention with 144-row
// windows and head_dim 32 needs ~1100 FLOP/clk/SM of tensor work if the rest of the kernel runs at the issue floor; this
// tells whether the m16 granularity (144 = 9 x 16 rows, no 128-row tail) is affordable.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int ACC>
__global__ void hmma_only(float* out, int iters) {
  float c[ACC][4] = {};
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
  uint32_t b0 = threadIdx.x * 5u, b1 = 11u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) mma16816(c[j], a, b0 + j, b1);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < ACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// one "softmax-like" element stream next to the MMAs: per accumulator element one fma, one ex2, one add; per pair one cvt
template <int ACC>
__global__ void hmma_softmax(float* out, int iters) {
  float c[ACC][4] = {};
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
  uint32_t b0 = threadIdx.x * 5u, b1 = 11u;
  float sum = 0.f;
  uint32_t pk = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < ACC; ++j) mma16816(c[j], a, b0 + j, b1);
#pragma unroll
    for (int j = 0; j < ACC; ++j) {
      float e[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float x = fmaf(c[j][q], 1.0009f, -3.f);
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[q]) : "f"(x));
        sum += e[q];
        c[j][q] = 0.f;
      }
      uint32_t p0, p1;
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(e[1]), "f"(e[0]));
      asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(e[3]), "f"(e[2]));
      pk ^= p0 + p1;
    }
    a[0] ^= pk;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum + (float)pk;
}

__global__ void ldmatrix_rate(float* out, int iters) {
  extern __shared__ __align__(128) uint8_t smem[];
  for (int i = threadIdx.x; i < 36864 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i;
  __syncthreads();
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  int lane = threadIdx.x & 31;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 18; ++j) {
      // 8 rows of 64 bytes, 64B-swizzled chunks: conflict-free
      int row = j * 8 + (lane & 7), chunk = (lane >> 3) ^ ((row >> 1) & 3);
      uint32_t addr = base + row * 64 + chunk * 16 + ((it & 3) * 9216);
      uint32_t r0, r1, r2, r3;
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
      acc += r0 ^ r1 ^ r2 ^ r3;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}

__global__ void movmatrix_rate(float* out, int iters) {
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * (j + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(v[j]) : "r"(v[j]));
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc ^= v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, 1 << 24);
  const int iters = 20000;
  printf("SMs %d, nominal clock %.0f MHz (rates below use it)\n", sms, khz / 1e3);
  for (int warps : {4, 8, 9, 16, 18, 32}) {
    float ms = time_ms([&] { hmma_only<8><<<sms, warps * 32>>>(out, iters); });
    double flop = 2.0 * 16 * 8 * 16 * 8 * (double)iters * warps * sms;
    printf("hmma only      %2d warps/SM: %8.3f ms  %7.1f TFLOP/s  %6.0f FLOP/clk/SM\n", warps, ms, flop / ms / 1e9,
           flop / sms / (ms * 1e-3 * khz * 1e3));
  }
  for (int warps : {8, 9, 16, 18, 32}) {
    float ms = time_ms([&] { hmma_softmax<8><<<sms, warps * 32>>>(out, iters / 4); });
    double flop = 2.0 * 16 * 8 * 16 * 8 * (double)(iters / 4) * warps * sms;
    double elems = 32.0 * 8 * 4 * (double)(iters / 4) * warps;
    printf("hmma + softmax %2d warps/SM: %8.3f ms  %7.1f TFLOP/s  %6.0f FLOP/clk/SM  %5.2f elem/clk/SM (MUFU peak 16)\n", warps, ms,
           flop / ms / 1e9, flop / sms / (ms * 1e-3 * khz * 1e3), elems / (ms * 1e-3 * khz * 1e3));
  }
  cudaFuncSetAttribute(ldmatrix_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  for (int warps : {9, 18}) {
    float ms = time_ms([&] { ldmatrix_rate<<<sms, warps * 32, 40960>>>(out, iters / 4); });
    double bytes = 512.0 * 18 * (iters / 4) * warps;
    printf("ldmatrix.x4    %2d warps/SM: %8.3f ms  %6.1f B/clk/SM\n", warps, ms, bytes / (ms * 1e-3 * khz * 1e3));
  }
  for (int warps : {9, 18}) {
    float ms = time_ms([&] { movmatrix_rate<<<sms, warps * 32>>>(out, iters); });
    double n = 8.0 * iters * warps;
    printf("movmatrix      %2d warps/SM: %8.3f ms  %6.3f per clk/SM\n", warps, ms, n / (ms * 1e-3 * khz * 1e3));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
