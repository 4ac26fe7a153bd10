// Micro-probe: how does tcgen05.mma (M=128, kind::f16, cta_group::1) issue/execute when consecutive
// MMAs hit the same TMEM accumulator vs. alternate between several?  One CTA per SM, operands are
// fixed (garbage) swizzled smem tiles; timing with clock64 around the issue loop + final commit.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int N, int NACC>
__global__ void __launch_bounds__(256, 1) probe(int n_mma, int same_operands, long long* out, int mode, const uint4* gsrc) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) * 4 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = (mode & 1) ? (0x3c003c00u ^ (i * 2654435761u & 0x03ff03ffu)) : 0;
  __shared__ uint64_t cbar;
  __shared__ volatile int stop_flag;
  if (threadIdx.x == 0) { stop_flag = 0; asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar))); }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 4 * 16384);
    long long t0 = clock64();
    const uint64_t ad0 = desc(a0), bd0 = desc(b0);
    const uint32_t stride_a = same_operands ? 0 : (16384 >> 4), stride_b = same_operands ? 0 : ((N * 128) >> 4);
    for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int acc = (NACC == 1) ? 0 : ((i / 4) & (NACC - 1));
        const uint32_t st = (i >> 2) & 3;
        mma(tmem + acc * N, ad0 + st * stride_a + 2 * k, bd0 + st * stride_b + 2 * k, idesc, i >= 4 * NACC);
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop_flag = 1;
  } else if ((mode & 2) && (warp == 2 || warp == 3 || warp == 4 || warp == 5)) {
    // epilogue-like TMEM readers on the upper columns
    uint32_t v[16]; float acc = 0.f;
    while (!stop_flag) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15])
        : "r"(tmem + (((uint32_t)(warp & 3) * 32u) << 16) + 448u) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += __uint_as_float(v[0]);
    }
    if (acc == 123.f) out[2] = 1;
  } else if ((mode & 4) && warp == 6 && lane == 0) {
    // TMA-like producer: 16 KB bulk copies global -> smem (a scratch region after the operands)
    uint8_t* dst = smem + 4 * 16384 + 4 * 256 * 128;
    uint32_t ph = 0;
    while (!stop_flag) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar)), "r"(16384) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_u32(dst)), "l"(gsrc + (blockIdx.x * 1024)), "r"(16384), "r"(smem_u32(&cbar)) : "memory");
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&cbar)), "r"(ph) : "memory");
      ph ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}
template <int N, int NACC>
void run(int same, int mode = 0) {
  const int n_acc = NACC;
  long long* out; cudaMalloc(&out, 16);
  const int smem = 4 * 16384 + 4 * 256 * 128 + 16384 + 2048;
  static uint4* gsrc = nullptr; if (!gsrc) { cudaMalloc(&gsrc, 148 * 16384); cudaMemset(gsrc, 1, 148 * 16384); }
  cudaFuncSetAttribute(probe<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n = 2048;
  probe<N, NACC><<<148, 256, smem>>>(n, same, out, mode, gsrc);
  probe<N, NACC><<<148, 256, smem>>>(n, same, out, mode, gsrc);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("mode=%d N=%3d acc=%d same_ops=%d : issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %d)  %s\n", mode, N, n_acc, same,
         (double)h[0] / n, (double)h[1] / n, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
}
int main() {
  for (int mode = 0; mode < 8; ++mode) { run<64, 1>(0, mode); run<32, 1>(0, mode); run<128, 1>(0, mode); run<256, 1>(0, mode); }
  return 0;
}
