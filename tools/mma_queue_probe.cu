// Micro-benchmark: depth of the asynchronous tcgen05.mma issue queue.  One thread issues 48 MMAs
// back to back into an idle tensor pipe and records the clock after every issue: the issue interval
// jumps from the instruction cost to the MMA execution time once the queue is full.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_queue_probe tools/mma_queue_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

template <int COMMIT_EVERY>
__global__ void __launch_bounds__(128, 1) probe(int N, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar, bar2;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)raw)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)) : "memory");
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
    if (pred) {
      const uint64_t dA = make_desc(base), dB = make_desc(base + 96 * 1024);
      long long t[49];
      t[0] = clock64();
#pragma unroll
      for (int j = 0; j < 48; ++j) {
        mma(tm, dA + (uint64_t)((j & 1) * 2), dB + (uint64_t)((j & 1) * 2), idesc);
        if (COMMIT_EVERY && (j + 1) % COMMIT_EVERY == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        t[j + 1] = clock64();
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      if (blockIdx.x == 0)
        for (int j = 0; j <= 48; ++j) out[j] = t[j] - t[0];
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

template <int CE>
static void run(int N) {
  long long* d;
  cudaMalloc(&d, 49 * 8);
  cudaFuncSetAttribute(probe<CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<CE><<<1, 128, 200 * 1024>>>(N, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[49];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%d commit_every=%d %s\n  ", N, CE, e == cudaSuccess ? "" : cudaGetErrorString(e));
  for (int j = 1; j <= 48; ++j) printf("%lld ", h[j] - h[j - 1]);
  printf("\n");
  cudaFree(d);
}

int main() {
  run<0>(256); run<0>(32); run<6>(256); run<6>(32); run<1>(32);
  return 0;
}
