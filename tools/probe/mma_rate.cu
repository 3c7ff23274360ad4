// Probe (run on a B200): issue cost of the tcgen05.mma flavours the attention kernels use.  One CTA, one issuing
// thread; operands are whatever is in shared memory (timing only).  For each flavour: R back-to-back MMAs + commit +
// wait, cycles measured with clock64 around (a) the issue loop alone and (b) until the commit barrier completes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../multi-modal-monodepth-estimation_b200/csrc/tc_ptx.cuh"
using namespace b200swin;

__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t base = (ptx::smem_u32(smem_dyn) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_dyn)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t kSw64 = 4;
    const uint64_t k64 = ptx::make_smem_desc(base, 16, 512, kSw64);              // K-major 64 B rows
    const uint64_t mn64 = ptx::make_smem_desc(base + 32768, 512, 512, kSw64);    // MN-major 64 B rows (N = 32)
    const uint64_t pk = ptx::make_smem_desc(base + 65536, 16, 1024, 2);          // panel K-major (128 B swizzle)
    const uint64_t pmn = ptx::make_smem_desc(base + 65536, 16384, 1024, 2);      // panel MN-major
    const uint32_t id_s = ptx::make_idesc_bf16(128, 144, 0, 0);
    const uint32_t id_dq = ptx::make_idesc_bf16(128, 32, 0, 1);
    const uint32_t id_t = ptx::make_idesc_bf16(128, 32, 1, 1);
    const uint32_t id_t64 = ptx::make_idesc_bf16(64, 32, 1, 1);
    const uint32_t id_pv = ptx::make_idesc_bf16(128, 32, 0, 1);
    uint32_t phase = 0;
    const int R = 32;
    for (int flavour = 0; flavour < 7; ++flavour) {
      for (int rep = 0; rep < 2; ++rep) {
        const long long t0 = clock64();
        for (int i = 0; i < R; ++i) {
          const uint32_t acc = i != 0;
          switch (flavour) {
            case 0: ptx::mma_bf16_ss(tm, k64 + 2 * (i & 1), k64 + 576 + 2 * (i & 1), id_s, acc); break;          // S
            case 1: ptx::mma_bf16_ss(tm + 288, pk + 2 * (i & 3), mn64 + 64 * (i & 7), id_dq, acc); break;       // dQ
            case 2: ptx::mma_bf16_ss(tm + 352, pmn + 128 * (i & 7), mn64 + 64 * (i & 7), id_t, acc); break;     // dV main
            case 3: ptx::mma_bf16_ss(tm + 416, pmn + 2048 + 128 * (i & 7), mn64 + 64 * (i & 7), id_t64, acc); break;  // dV tail M=64
            case 4: ptx::mma_bf16_ts(tm + 320, tm + 8 * (i & 7), mn64 + 64 * (i & 7), id_pv, acc); break;       // PV (A in TMEM)
            case 5:                                                                                              // dV/dK interleaved on 4 accumulators
              ptx::mma_bf16_ss(tm + 352 + 32 * (i & 1), pmn + 128 * ((i >> 2) & 7), mn64 + 64 * ((i >> 2) & 7),
                               (i & 2) ? id_t64 : id_t, i >= 4);
              break;
            default: ptx::mma_bf16_ss(tm + 352, pmn + 128 * (i & 7), k64 + 2 * (i & 1), ptx::make_idesc_bf16(128, 32, 1, 0), acc); break;  // A MN-major, B K-major
          }
        }
        const long long t1 = clock64();
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, phase);
        phase ^= 1;
        const long long t2 = clock64();
        out[(flavour * 2 + rep) * 2] = t1 - t0;
        out[(flavour * 2 + rep) * 2 + 1] = t2 - t0;
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
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  probe<<<1, 128, 170 * 1024>>>(d);
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  long long h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[7] = {"S  M128 N144 K-major/K-major sw64", "dQ A K-major sw128, B MN-major sw64 N32", "dV A MN-major sw128 M128, B MN sw64",
                          "dV tail A MN-major M64", "PV A in TMEM, B MN sw64", "dV/dK interleaved 4 accumulators M128/M64", "A MN-major M128, B K-major"};
  for (int f = 0; f < 7; ++f)
    printf("%-45s: issue %6lld cyc, complete %6lld cyc for 32 MMAs (second rep: %lld / %lld)\n", names[f], h[f * 4], h[f * 4 + 1],
           h[f * 4 + 2], h[f * 4 + 3]);
  return 0;
}
