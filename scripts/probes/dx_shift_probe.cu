// Micro-probe for the next step of conv_tc_kernel: can ONE TMA box of 16 (w) x 18 (h) pixels serve all NINE taps
// of a 3x3 convolution tile (8 x 16 pixels), by moving only the start address of the tcgen05 A descriptor?
//   dy taps: + dy * 2048 B (a whole 16-pixel image row = two 1 KB swizzle atoms)      -- the trick conv_tc already uses
//   dx taps: + dx * 128 B  (ONE pixel row inside a swizzle atom), descriptor "base offset" = dx (PTX: (addr >> 7) & 7
//            when the start address is not aligned to the 1 KB repeating pattern), stride between 8-row groups = 2048 B
// Today the kernel fetches three 8 x 18 boxes (one per dx): 3.375x the tile's input from L2 and three TMA instructions
// per K chunk; the 16-wide box would be 2.25x and one instruction, in 36 KB of shared memory instead of 55 KB.
//
// The probe loads X[18][16][64] (fp16) with TMA (128B swizzle) and W[64][64], runs for every (dy, dx) one K=64 MMA chain
// (M=128, N=64) with the shifted descriptor -- once with base_offset = dx, once with base_offset = 0 -- and compares
// D[m][n] = sum_c X[m/8 + dy][m%8 + dx][c] * W[n][c] with a host reference.  Prints the max error per tap and variant.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o dx_shift_probe dx_shift_probe.cu && ./dx_shift_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int H = 18, W = 16, C = 64, N = 64;
constexpr int A_BYTES = H * W * C * 2, B_BYTES = N * C * 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address
  d |= (uint64_t)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(sbo_bytes >> 4) << 32;              // stride between 8-row groups
  d |= (uint64_t)1 << 46;                             // descriptor version
  d |= (uint64_t)(base_offset & 7) << 49;             // matrix base offset
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (clock64() - t0 > 2000000000LL) { printf("probe: mbarrier timeout\n"); __trap(); }
  }
}

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;
  uint8_t* sb = smem + A_BYTES;                        // 36864 = 36 KB: still 1 KB aligned
  __shared__ uint64_t full, done;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&done)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full)), "r"(A_BYTES + B_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(sa)), "l"((uint64_t)&amap), "r"(smem_u32(&full)), "r"(0), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sb)), "l"((uint64_t)&bmap), "r"(smem_u32(&full)), "r"(0), "r"(0) : "memory");
  }
  mbar_wait(&full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant)
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
      if (threadIdx.x == 0) {
        const uint32_t a_addr = smem_u32(sa) + dy * (W * C * 2) + dx * (C * 2);
        const uint64_t ad = make_desc(a_addr, W * C * 2 /* 2048 B between image rows */, variant == 0 ? dx : 0);
        const uint64_t bd = make_desc(smem_u32(sb), 1024, 0);
        for (int k = 0; k < 4; ++k)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(k > 0 ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
      }
      mbar_wait(&done, phase);
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* o = out + ((size_t)(variant * 9 + tap) * 128 + warp * 32 + lane) * N;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) o[c0 + j] = __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();                                  // everyone has read the accumulator before the next tap overwrites it
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  std::vector<__half> x((size_t)H * W * C), w((size_t)N * C);
  std::vector<float> xf(x.size()), wf(w.size());
  srand(1);
  for (size_t i = 0; i < x.size(); ++i) { x[i] = __float2half((rand() % 2001 - 1000) / 1000.0f); xf[i] = __half2float(x[i]); }
  for (size_t i = 0; i < w.size(); ++i) { w[i] = __float2half((rand() % 2001 - 1000) / 4000.0f); wf[i] = __half2float(w[i]); }
  __half *dx, *dw;
  float* dout;
  cudaMalloc(&dx, x.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&dout, sizeof(float) * 18 * 128 * N);
  cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  EncodeFn enc = (EncodeFn)fp;
  CUtensorMap amap, bmap;
  cuuint64_t adims[3] = {C, W, H}, astr[2] = {C * 2, (cuuint64_t)W * C * 2};
  cuuint32_t abox[3] = {C, W, H}, one3[3] = {1, 1, 1};
  cuuint64_t bdims[2] = {C, N}, bstr[1] = {C * 2};
  cuuint32_t bbox[2] = {C, N}, one2[2] = {1, 1};
  CUresult r1 = enc(&amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, dx, adims, astr, abox, one3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&bmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, bdims, bstr, bbox, one2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
  const int smem = A_BYTES + B_BYTES + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(amap, bmap, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out((size_t)18 * 128 * N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  for (int variant = 0; variant < 2; ++variant) {
    printf("base_offset = %s:\n", variant == 0 ? "dx" : "0");
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dxx = tap % 3;
      double worst = 0;
      int bad_rows = 0;
      for (int m = 0; m < 128; ++m) {
        double row_worst = 0;
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int c = 0; c < C; ++c) ref += (double)xf[((size_t)(m / 8 + dy) * W + (m % 8 + dxx)) * C + c] * wf[(size_t)n * C + c];
          row_worst = fmax(row_worst, fabs(ref - out[((size_t)(variant * 9 + tap) * 128 + m) * N + n]));
        }
        worst = fmax(worst, row_worst);
        bad_rows += row_worst > 1e-2;
      }
      printf("  tap dy=%d dx=%d: max err %.3e, rows off %d/128 %s\n", dy, dxx, worst, bad_rows, worst < 1e-2 ? "OK" : "MISMATCH");
    }
  }
  return 0;
}
