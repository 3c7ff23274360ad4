// Probe (run on a B200): how fast can one warp issue tcgen05.mma?  Same 32 small MMAs (M128 N32 K16) issued
//   (a) from `if (lane == 0)` (divergent region, what the kernels did),
//   (b) by a converged warp with an elect.sync predicate,
//   (c) like (b), fully unrolled with precomputed descriptors.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../multi-modal-monodepth-estimation_b200/csrc/tc_ptx.cuh"
using namespace b200swin;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 100 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_dyn)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    constexpr uint32_t kSw64 = 4;
    const uint64_t pk = ptx::make_smem_desc(base + 65536, 16, 1024, 2);
    const uint64_t mn64 = ptx::make_smem_desc(base + 32768, 512, 512, kSw64);
    const uint32_t id_dq = ptx::make_idesc_bf16(128, 32, 0, 1);
    uint32_t phase = 0;
    for (int variant = 0; variant < 3; ++variant) {
      for (int rep = 0; rep < 2; ++rep) {
        __syncwarp();
        const long long t0 = clock64();
        if (variant == 0) {
          if (lane == 0) {
            for (int i = 0; i < 32; ++i) ptx::mma_bf16_ss(tm + 288, pk + 2 * (i & 3), mn64 + 64 * (i & 7), id_dq, i != 0);
          }
        } else if (variant == 1) {
          for (int i = 0; i < 32; ++i) {
            if (elect_one()) ptx::mma_bf16_ss(tm + 288, pk + 2 * (i & 3), mn64 + 64 * (i & 7), id_dq, i != 0);
          }
        } else {
          if (elect_one()) {
#pragma unroll
            for (int i = 0; i < 32; ++i) ptx::mma_bf16_ss(tm + 288, pk + 2 * (i & 3), mn64 + 64 * (i & 7), id_dq, i != 0);
          }
        }
        __syncwarp();
        const long long t1 = clock64();
        if (lane == 0) ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, phase);
        phase ^= 1;
        const long long t2 = clock64();
        if (lane == 0) {
          out[(variant * 2 + rep) * 2] = t1 - t0;
          out[(variant * 2 + rep) * 2 + 1] = t2 - t0;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tm, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  cudaMemset(d, 0, 64 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
  probe<<<1, 128, 110 * 1024>>>(d);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  long long h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[3] = {"if (lane == 0) loop", "converged warp, elect.sync per MMA (loop)", "elect.sync once, 32 MMAs unrolled"};
  for (int f = 0; f < 3; ++f)
    printf("%-45s: issue %6lld cyc, complete %6lld cyc for 32 MMAs (second rep: %lld / %lld)\n", names[f], h[f * 4], h[f * 4 + 1],
           h[f * 4 + 2], h[f * 4 + 3]);
  return 0;
}
