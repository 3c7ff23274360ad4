// Probe (run on a B200): which (lane, column) of tensor memory lands in which thread/register for the
// 16-lane tcgen05.ld / tcgen05.st shapes.  TMEM is filled with (lane << 8 | column) through the well-understood
// 32x32b shape; the 16x256b / 16x128b loads are then dumped, and a 16x128b STORE is read back through 32x32b.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = slot;
  const uint32_t t_lane = base + ((uint32_t)(warp * 32) << 16);
  // fill 32 columns: value = absolute lane << 8 | column
  for (int c = 0; c < 32; c += 8) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = ((uint32_t)(warp * 32 + lane) << 8) | (uint32_t)(c + i);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(t_lane + c),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
  }
  asm volatile("tcgen05.wait::st.sync.aligned;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  if (warp == 1) {   // quarter 1: lanes 32..63
    uint32_t r[8];
    // 16x256b.x1 at lane offset 0 of this quarter, column 8
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(t_lane + 8));
    // same shape at lane offset 16
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(t_lane + (16u << 16) + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 8; ++i) out[lane * 8 + i] = r[i];
    uint32_t q[4];
    asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0, %1}, [%2];" : "=r"(q[0]), "=r"(q[1]) : "r"(t_lane + 4));
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(t_lane + 16));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    out[256 + lane * 2] = q[0];
    out[256 + lane * 2 + 1] = q[1];
    for (int i = 0; i < 8; ++i) out[320 + lane * 8 + i] = r[i];
    // store probe: 16x128b.x1 at columns 40..43, value = 0xS0000 | lane << 4 | reg
    uint32_t s0 = 0x50000u | (lane << 4) | 0u, s1 = 0x50000u | (lane << 4) | 1u;
    asm volatile("tcgen05.st.sync.aligned.16x128b.x1.b32 [%0], {%1, %2};" ::"r"(t_lane + 40), "r"(s0), "r"(s1));
    asm volatile("tcgen05.wait::st.sync.aligned;");
    uint32_t b[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(t_lane + 40));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int i = 0; i < 4; ++i) out[576 + lane * 4 + i] = b[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64));
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 4096 * 4);
  cudaMemset(d, 0, 4096 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[4096];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("16x256b.x1 @lane+0,col 8 (r0..r3) and @lane+16 (r4..r7): thread -> (lane,col)\n");
  for (int t = 0; t < 32; ++t) {
    printf("t%2d:", t);
    for (int i = 0; i < 8; ++i) printf(" (%u,%u)", h[t * 8 + i] >> 8, h[t * 8 + i] & 255);
    printf("\n");
  }
  printf("16x128b.x1 @col 4: thread -> (lane,col) x2\n");
  for (int t = 0; t < 32; ++t) printf("t%2d: (%u,%u) (%u,%u)\n", t, h[256 + 2 * t] >> 8, h[256 + 2 * t] & 255, h[257 + 2 * t] >> 8, h[257 + 2 * t] & 255);
  printf("16x256b.x2 @col 16: thread -> (lane,col) x8\n");
  for (int t = 0; t < 32; ++t) {
    printf("t%2d:", t);
    for (int i = 0; i < 8; ++i) printf(" (%u,%u)", h[320 + t * 8 + i] >> 8, h[320 + t * 8 + i] & 255);
    printf("\n");
  }
  printf("16x128b.x1 STORE @col 40 read back with 32x32b: lane -> 4 cols (thread<<4|reg)\n");
  for (int t = 0; t < 32; ++t) printf("lane%2d: %05x %05x %05x %05x\n", t, h[576 + 4 * t], h[577 + 4 * t], h[578 + 4 * t], h[579 + 4 * t]);
  return 0;
}
