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

template <int NACC, int ROT_A, int commit_every, int fence_every, int wait_every, int DELAY, int SPIN>
__global__ void __launch_bounds__(320, 1) probe(int N, int sw, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) stop_flag = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)raw)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])) : "memory");
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
          if (DELAY && (j + 1) % 6 == 0) {
            const long long d0 = clock64();
            while (clock64() - d0 < DELAY) {}
          }
          if (commit_every && (j + 1) % commit_every == 0)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[0])) : "memory");
          if (fence_every && (j + 1) % fence_every == 0) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (wait_every && (j + 1) % wait_every == 0) {
            uint32_t dn;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(dn) : "r"(smem_u32(&bar2[1])), "r"(1u) : "memory");
          }
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
    stop_flag = 1;
  } else if (SPIN && warp >= 2) {
    // spinning pollers, like epilogue warps waiting for an accumulator
    while (!stop_flag) {
      uint32_t dn;
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(dn) : "r"(smem_u32(&bar2[1])), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

template <int NACC, int ROT_A, int ce = 0, int fe = 0, int we = 0, int DELAY = 0, int SPIN = 0>
static void run(int N, int sw, int grid) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe<NACC, ROT_A, ce, fe, we, DELAY, SPIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  probe<NACC, ROT_A, ce, fe, we, DELAY, SPIN><<<grid, 320, 200 * 1024>>>(N, sw, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d nacc=%d commit/%d fence/%d wait/%d delay=%d spin=%d : %7.1f cyc/MMA (floor %d)%s\n", N, NACC, ce, fe, we, DELAY, SPIN,
         (double)h / (12.0 * iters), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int N : {32, 256}) {
    run<1, 1, 6, 6, 6, 0, 0>(N, 64, 148);
    run<1, 1, 6, 6, 6, 0, 1>(N, 64, 148);
    run<1, 1, 6, 6, 6, 100, 0>(N, 64, 148);
    run<1, 1, 6, 6, 6, 200, 0>(N, 64, 148);
    run<1, 1, 6, 6, 6, 400, 0>(N, 64, 148);
    run<1, 1, 6, 6, 6, 800, 0>(N, 64, 148);
    run<1, 1, 6, 6, 6, 400, 1>(N, 64, 148);
  }
  return 0;
}
