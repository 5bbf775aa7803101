// Micro-benchmark: cost of a tcgen05.mma (kind::f16, M=128, K=16, cta_group::1, 64-byte swizzle) as a function of WHERE its A
// tile starts and how its 8-row groups are strided - the shared-activation-box layouts of conv_tc.cuh read A tiles that start
// (dx + dxmax) * 64 bytes into a box row and whose 8-pixel groups are box_w * 64 bytes apart (640 B for 3x3 layers, 896 B for
// the phase-lattice MSBlock tail).  Operands are zeros; only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_align_probe tools/mma_align_probe.cu && tools/mma_align_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

// pattern 0: N alone; pattern 1: the wide pair of a narrow layer (N2 = 2N then N) on one accumulator
template <int NACC>
__global__ void __launch_bounds__(128, 1) probe(int N, int pair, uint32_t a_shift, uint32_t sbo, int iters, long long* out, uint32_t acc_stride) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
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
    // a_shift bit 0 set (never a real shift: shifts are multiples of 64): M = 64 instead of 128 (the transposed formulation
    // of the MSBlock tail would stack 2 x 32 weight rows as the M operand)
    const uint32_t M = (a_shift & 1u) ? 64u : 128u;
    a_shift &= ~1u;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((M >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * N) >> 3) << 17) | ((M >> 4) << 24);
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    long long t0 = 0, t1 = 0;
    if (pred) {
      const uint64_t dB = make_desc(base + 128 * 1024, 512);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j0 = 0; j0 + NACC <= 12; j0 += NACC) {
          // NACC independent accumulators (128 TMEM columns apart) are interleaved like the sub-tiles of the real schedule: all wide MMAs of the
          // group first, then the narrow ones; A rotates over four 32 KB regions and the two K steps
          if (pair) {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
              const int j = j0 + a;
              mma(tm + a * (acc_stride & ~7u), make_desc(base + (j & 3) * 32768 + a_shift, sbo) + (uint64_t)((j >> 2) * 2), dB, idesc2);
            }
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
              const int j = j0 + a;
              mma(tm + a * (acc_stride & ~7u), make_desc(base + (j & 3) * 32768 + a_shift, sbo) + (uint64_t)((j >> 2) * 2) + (16384 >> 4), dB, idesc);
            }
          } else {
#pragma unroll
            for (int a = 0; a < NACC; ++a) {
              const int j = j0 + a;
              mma(tm + a * (acc_stride & ~7u), make_desc(base + (j & 3) * 32768 + a_shift, sbo) + (uint64_t)((j >> 2) * 2), dB, idesc);
            }
          }
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    t1 = clock64();
    if (pred && blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
  }
}

template <int NACC>
static void run(long long* d, int N, int pair, uint32_t shift, uint32_t sbo, uint32_t acc_stride = 128) {
  const int iters = 2000;
  cudaFuncSetAttribute(probe<NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  probe<NACC><<<148, 128, 200 * 1024>>>(N, pair, shift, sbo, iters, d, acc_stride);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s M=%3d N=%3d accumulators=%d (%3u columns apart, boundary ops %u) sbo=%4u a_shift=%3u : %6.1f cycles%s\n", pair ? "pair" : "one ", (shift & 1u) ? 64 : 128, N, NACC, acc_stride & ~7u, acc_stride & 7u, sbo, shift & ~1u,
         (double)h / ((double)(12 / NACC * NACC) * iters), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  printf("# cycles per MMA (M=128, K=16); pair = wide N=2n MMA followed by the N=n MMA of a narrow layer (cycles per PAIR)\n");
  // M = 64 against M = 128 at N = 256 / 128 (three accumulators would need 768 columns at N = 256: two)
  run<2>(d, 256, 0, 0, 512, 256);
  run<2>(d, 256, 0, 1, 512, 256);
  run<3>(d, 128, 0, 0, 512, 128);
  run<3>(d, 128, 0, 1, 512, 128);
  // the merged VGG + MSBlock layers: N = 160 (one 160-column accumulator per buffer today) and N = 144 / 96
  for (int N : {160, 144, 96}) {
    run<1>(d, N, 0, 0, 512, 160);
    run<2>(d, N, 0, 0, 512, 160);
    run<3>(d, N, 0, 0, 512, 160);
  }
  for (int pair = 0; pair < 2; ++pair)
    for (int N : {32, 64, 128}) {
      if (pair && N == 128) continue;
      for (uint32_t sbo : {512u, 896u})
        for (uint32_t shift : {0u, 64u}) run<1>(d, N, pair, shift, sbo);
      run<2>(d, N, pair, 0, 512);
      run<3>(d, N, pair, 0, 512);
      run<4>(d, N, pair, 0, 512);     // accumulators sit 128 TMEM columns apart: 4 x 128 = 512
      if (N == 32) {
        // per 12 MMAs: +1 = a tcgen05.fence::after_thread_sync, +2 = an mbarrier try_wait (already complete), +4 = a tcgen05.commit
        for (uint32_t extra : {1u, 2u, 4u, 7u}) run<4>(d, N, pair, 0, 512, 64 + extra);
      }
      if (N == 32) {                  // the real layouts: sub-tiles / groups 64 columns apart (wide) or 32 (plain)
        run<3>(d, N, pair, 0, 512, 64);
        run<4>(d, N, pair, 0, 512, 64);
        if (!pair) { run<3>(d, N, pair, 0, 512, 32); run<4>(d, N, pair, 0, 512, 32); }
      }
    }
  return 0;
}
