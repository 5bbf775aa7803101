// Probe: does a tcgen05 shared-memory matrix descriptor (K-major, SWIZZLE_64B) whose start address is
// NOT aligned to the 512-byte swizzle atom address the same bytes TMA would have written, i.e. is the
// XOR swizzle a function of the absolute shared-memory address?  If so, one activation box can serve
// every horizontal tap offset (start += dx * 64 B) and an arbitrary row pitch (SBO = pitch * 64 B).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o desc_shift_probe tools/desc_shift_probe.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}

#define ROWS 640   // logical pixel rows staged (64 B each)
#define NB 32

// mode: 0 = base_offset 0, 1 = base_offset (start >> 7) & 7
__global__ void __launch_bounds__(128, 1) probe(int shift, int sbo, int mode, float* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  // A: ROWS x 32 channels; B: NB x 32 channels, both written with the absolute-address 64B swizzle
  for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) {
    const int R = i / 32, c = i % 32;
    const float v = (float)(((R * 7 + c * 3) % 13) - 6);
    uint32_t off = R * 64 + c * 2;
    off ^= ((off >> 7) & 3u) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sm + off) = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < NB * 32; i += blockDim.x) {
    const int n = i / 32, c = i % 32;
    const float v = (float)(((n * 5 + c) % 7) - 3);
    uint32_t off = ROWS * 64 + n * 64 + c * 2;
    off ^= ((off >> 7) & 3u) << 4;
    *reinterpret_cast<__nv_bfloat16*>(sm + off) = __float2bfloat16(v);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sa = base + shift * 64, sb = base + ROWS * 64;
    for (int k = 0; k < 2; ++k) {
      const uint64_t dA = make_desc(sa, sbo, mode ? (sa >> 7) : 0) + (uint64_t)(k * 2);
      const uint64_t dB = make_desc(sb, 512, 0) + (uint64_t)(k * 2);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tm), "l"(dA), "l"(dB), "r"(idesc), "r"((uint32_t)k)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 32; ++n) out[threadIdx.x * 32 + n] = __uint_as_float(r[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u) : "memory");
  }
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  static float h[128 * 32];
  const int shifts[] = {0, 8, 1, 2, 3, 4, 5, 7, 9, 12};
  const int sbos[] = {512, 640, 1024, 2560};
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : sbos)
      for (int shift : shifts) {
        cudaMemset(d, 0, sizeof(h));
        probe<<<1, 128, 64 * 1024>>>(shift, sbo, mode, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("shift %d sbo %d mode %d: %s\n", shift, sbo, mode, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int R = shift + (m / 8) * (sbo / 64) + (m % 8);
          for (int n = 0; n < 32; ++n) {
            float ref = 0;
            for (int c = 0; c < 32; ++c) ref += (float)(((R * 7 + c * 3) % 13) - 6) * (float)(((n * 5 + c) % 7) - 3);
            if (h[m * 32 + n] != ref) ++bad;
          }
        }
        printf("base_offset_mode %d  SBO %4d  start shift %2d rows (%4d B): %s (%d / 4096 mismatches)\n", mode, sbo, shift,
               shift * 64, bad ? "MISMATCH" : "exact", bad);
      }
  return 0;
}
