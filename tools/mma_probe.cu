// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) as a function
// of N, of how many independent accumulators the stream rotates over, and of the operand swizzle.
// Operands are whatever is in shared memory (zeros); only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe tools/mma_probe.cu && ./mma_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int sw) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sw == 128 ? 1024 : 512) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(sw == 128 ? 2 : 4) << 61;
  return d;
}

__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

template <int NACC, int ROT_A>
__global__ void __launch_bounds__(128, 1) probe(int N, int sw, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)raw)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    long long t0 = 0, t1 = 0;
    if (pred) {
      const uint64_t dB = make_desc(base + 96 * 1024, sw);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 12; ++j) {
          const uint64_t dA = make_desc(base + (ROT_A ? (j % 6) * 8192 : 0), sw) + (uint64_t)((j & 1) * 2);
          mma(tm + (j % NACC) * N, dA, dB + (uint64_t)((j & 1) * 2), idesc, 1u);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    t1 = clock64();
    if (pred && blockIdx.x == 0) out[0] = t1 - t0;
    (void)lane;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

template <int NACC, int ROT_A>
static void run(int N, int sw, int grid) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe<NACC, ROT_A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  probe<NACC, ROT_A><<<grid, 128, 200 * 1024>>>(N, sw, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d sw=%3d nacc=%d rotA=%d grid=%3d : %7.1f cyc/MMA (floor %d)%s\n", N, sw, NACC, ROT_A, grid,
         (double)h / (12.0 * iters), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int Ns[4] = {32, 64, 128, 256};
  for (int sw : {64, 128})
    for (int i = 0; i < 4; ++i) {
      const int N = Ns[i];
      run<1, 0>(N, sw, 1);
      if (2 * N <= 512) run<2, 0>(N, sw, 1);
      if (4 * N <= 512) run<4, 0>(N, sw, 1);
      run<1, 1>(N, sw, 1);
      if (4 * N <= 512) run<4, 1>(N, sw, 1);
    }
  run<4, 1>(32, 64, 148);
  run<2, 1>(256, 64, 148);
  run<2, 1>(256, 128, 148);
  return 0;
}
