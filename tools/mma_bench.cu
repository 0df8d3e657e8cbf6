// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16) issued back to back from one thread,
// operands in shared memory (K-major, 128B / 32B swizzle), for several N.  One CTA per SM, all SMs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_bench tools/mma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N, int ROWB>  // ROWB: bytes per operand row (128: SW128 with 4 k-steps per row, 32: SW32 with 1)
__global__ void __launch_bounds__(128, 1) mma_bench(int iters, long long *out, int a_stride_rows, int distinct_a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    const unsigned layout = (ROWB == 128) ? 2u : (ROWB == 64) ? 4u : 6u;
    const unsigned sbo = (unsigned)(8 * ROWB) >> 4;
    const uint64_t desc_hi = (uint64_t)(sbo | (1u << 14) | (layout << 29)) << 32;
    const uint64_t desc0 = desc_hi | (uint64_t)(1u << 16);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
    const uint32_t a16 = (base & 0x3FFFFu) >> 4;
    const uint32_t b16 = ((base + 128 * 1024) & 0x3FFFFu) >> 4;
    // 8 operand positions precomputed (emulates tap shifts / fresh tiles); the issue loop is fully unrolled by 8
    uint64_t da[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) da[q] = desc0 | (uint64_t)(a16 + ((uint32_t)((q % distinct_a) * a_stride_rows * ROWB) >> 4));
    const uint64_t db = desc0 | (uint64_t)b16;
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(da[q]), "l"(db), "r"(idesc), "r"(1u));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    long long t1 = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

template <int N, int ROWB>
void run(const char *name, int stride_rows, int distinct) {
  long long *d, h[2];
  cudaMalloc(&d, 16);
  auto k = mma_bench<N, ROWB>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2048;
  k<<<148, 128, 200 * 1024>>>(iters, d, stride_rows, distinct);
  k<<<148, 128, 200 * 1024>>>(iters, d, stride_rows, distinct);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d rowB=%3d stride_rows=%3d distinct=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", name, N, ROWB,
         stride_rows, distinct, (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<16, 32>("sw32 same A", 0, 1);
  run<16, 32>("sw32 shifted A (taps)", 1, 9);
  run<16, 32>("sw32 distinct A blocks", 128, 8);
  run<64, 32>("sw32 N=64", 0, 1);
  run<16, 128>("sw128 N=16", 0, 1);
  run<64, 128>("sw128 N=64 same A", 0, 1);
  run<64, 128>("sw128 N=64 shifted A", 1, 9);
  run<64, 128>("sw128 N=64 43-row shifts", 43, 3);
  run<128, 128>("sw128 N=128", 0, 1);
  run<256, 128>("sw128 N=256", 0, 1);
  run<80, 32>("sw32 N=80", 0, 1);
  return 0;
}
