// conv_tc.cu -- implicit-GEMM convolution / dense layer on the 5th-gen tensor cores (sm_100a).
//
// Replaces the TensorFlow Conv2D / Dense kernels behind keras.Model.predict for CRAFT
// (reference detection.py:65-103, 365-410) and the CRNN (recognition.py:217-290, 292-327).
//
//   D[pixel, cout] = sum_{tap, c} X[pixel + offset(tap), c] * Wt[cout, tap, c]
//
// * M = 128 output pixels per tile, chosen as a (BW x BH x BNI) box of the NHWC activation so one
//   4-D TMA load per (tap, 64-channel chunk) fetches the shifted A tile; TMA's out-of-bounds
//   zero fill implements "same" padding and dilation for free.
// * B = weights packed K-major [cout][tap*cin + c], loaded with a 2-D TMA box (64 x BLOCK_N).
// * both operands land in 128B-swizzled shared memory; one elected thread issues
//   tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16) x4 per stage; fp32 accumulators
//   live in TMEM (2 stages) so the epilogue of tile i overlaps the main loop of tile i+1.
// * warp roles: warp0 = TMA producer, warp1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
//   (tcgen05.ld -> scale/shift/ReLU/affine -> fp16 or fp32 NHWC stores, optionally into a channel
//   slice of a wider concat buffer).
// * persistent: grid = min(#tiles, #SMs); tiles are walked n-tile fastest so the CTAs that share
//   an A tile run concurrently and hit L2.
#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;               // fp16 elements per stage along K (= 128 B swizzle span)
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 192;
constexpr int SMEM_BUDGET = 200 * 1024;

struct TcParams {
  int N, H, W;
  int cin, cout, ksize, dil;
  int bw_log2, bh_log2, bn_log2;
  int tiles_w, tiles_h, tiles_n;          // M-tile grid
  int n_tiles;                            // cout / BLOCK_N
  int total_tiles;
  const float *s1, *t1, *s2, *t2;
  int relu;
  void* out;
  int out_ld;
  int out_f32;
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a broken pipeline traps (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("b2ocr conv_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (rows of 128 B, 8-row atoms 1024 B apart).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);   // start address      bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                    // leading byte off.  bits [16,30) (unused here)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;            // stride byte offset bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                    // descriptor version bits [46,48)
  d |= static_cast<uint64_t>(2) << 61;                    // SWIZZLE_128B       bits [61,64)
  return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <int CH>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int BLOCK_N>
struct TcConfig {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = (2 * BLOCK_N) < 32 ? 32 : (2 * BLOCK_N);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int CH = BLOCK_N >= 32 ? 32 : 16;     // accumulator columns per tcgen05.ld
};

// ------------------------------------------------------------------------------------------ kernel
template <int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
               const TcParams p) {
  using Cfg = TcConfig<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&amap);
    tma_prefetch_desc(&bmap);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_before_sync();
  __syncthreads();
  tcgen05_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int taps = p.ksize * p.ksize;
  const int kchunks = p.cin / BLOCK_K;
  const int k_iters = taps * kchunks;
  const int half_k = p.ksize >> 1;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int tn = m_tile / (p.tiles_w * p.tiles_h);
        const int w0 = tw << p.bw_log2, h0 = th << p.bh_log2, n0 = tn << p.bn_log2;
        for (int tap = 0; tap < taps; ++tap) {
          const int dy = (tap / p.ksize - half_k) * p.dil;
          const int dx = (tap % p.ksize - half_k) * p.dil;
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_4d(&amap, &full_bar[stage], smem_a + stage * A_STAGE_BYTES, kc * BLOCK_K, w0 + dx, h0 + dy, n0);
            tma_load_2d(&bmap, &full_bar[stage], smem_b + stage * Cfg::B_STAGE_BYTES, tap * p.cin + kc * BLOCK_K,
                        n_tile * BLOCK_N);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=f16, both K-major, N=BLOCK_N, M=128
      constexpr uint32_t idesc = (1u << 4) | (static_cast<uint32_t>(BLOCK_N >> 3) << 17) |
                                 (static_cast<uint32_t>(BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tcgen05_after_sync();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_after_sync();
          const uint64_t adesc = umma_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
          const uint64_t bdesc = umma_desc(smem_u32(smem_b + stage * Cfg::B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 fp16 = 32 B along K inside the swizzle atom: +2 in 16-byte units
            umma_f16(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                     (it | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                 // frees the smem slot when the MMAs retire
          if (it == k_iters - 1) umma_commit(&tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    constexpr int CH = Cfg::CH;
    const int quad = warp & 3;                            // TMEM lane quadrant this warp may touch
    const int row = quad * 32 + lane;                     // accumulator row = pixel inside the tile
    const int bw_mask = (1 << p.bw_log2) - 1, bh_mask = (1 << p.bh_log2) - 1;
    const int wi = row & bw_mask;
    const int hi = (row >> p.bw_log2) & bh_mask;
    const int ni = row >> (p.bw_log2 + p.bh_log2);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int tw = m_tile % p.tiles_w;
      const int th = (m_tile / p.tiles_w) % p.tiles_h;
      const int tn = m_tile / (p.tiles_w * p.tiles_h);
      const int w = (tw << p.bw_log2) + wi, h = (th << p.bh_log2) + hi, n = (tn << p.bn_log2) + ni;
      const bool valid = (w < p.W) && (h < p.H) && (n < p.N);
      const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
      const int c_base = n_tile * BLOCK_N;

      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_after_sync();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);
#pragma unroll 1
      for (int ch = 0; ch < BLOCK_N / CH; ++ch) {
        uint32_t v[CH];
        tmem_ld<CH>(taddr + static_cast<uint32_t>(ch * CH), v);
        tmem_ld_wait();
        const int c0 = c_base + ch * CH;
        float y[CH];
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p.s1 + c0 + j));
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.t1 + c0 + j));
          y[j + 0] = fmaf(__uint_as_float(v[j + 0]), a.x, b.x);
          y[j + 1] = fmaf(__uint_as_float(v[j + 1]), a.y, b.y);
          y[j + 2] = fmaf(__uint_as_float(v[j + 2]), a.z, b.z);
          y[j + 3] = fmaf(__uint_as_float(v[j + 3]), a.w, b.w);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < CH; ++j) y[j] = fmaxf(y[j], 0.0f);
        }
        if (p.s2 != nullptr) {
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.s2 + c0 + j));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.t2 + c0 + j));
            y[j + 0] = fmaf(y[j + 0], a.x, b.x);
            y[j + 1] = fmaf(y[j + 1], a.y, b.y);
            y[j + 2] = fmaf(y[j + 2], a.z, b.z);
            y[j + 3] = fmaf(y[j + 3], a.w, b.w);
          }
        }
        if (valid) {
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + pix * p.out_ld + c0;
#pragma unroll
            for (int j = 0; j < CH; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          } else {
            __half* o = reinterpret_cast<__half*>(p.out) + pix * p.out_ld + c0;
#pragma unroll
            for (int j = 0; j < CH; j += 8) {
              __half2 h0 = __floats2half2_rn(y[j + 0], y[j + 1]);
              __half2 h1 = __floats2half2_rn(y[j + 2], y[j + 3]);
              __half2 h2 = __floats2half2_rn(y[j + 4], y[j + 5]);
              __half2 h3 = __floats2half2_rn(y[j + 6], y[j + 7]);
              uint4 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(o + j) = pk;
            }
          }
        }
      }
      tcgen05_before_sync();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_before_sync();
  __syncthreads();
  if (warp == 1) {
    tcgen05_after_sync();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// Pick the (bw, bh, bn) power-of-two box with bw*bh*bn = 128 that wastes the fewest pixels.
void pick_box(int N, int H, int W, int* bw_l, int* bh_l, int* bn_l) {
  double best = 1e30;
  int b_w = 7, b_h = 0, b_n = 0;
  for (int lw = 0; lw <= 7; ++lw)
    for (int lh = 0; lw + lh <= 7; ++lh) {
      const int ln = 7 - lw - lh;
      const int bw = 1 << lw, bh = 1 << lh, bn = 1 << ln;
      if (ln > 0 && (bn > 2 * N)) continue;
      const double cover = double((W + bw - 1) / bw * bw) * double((H + bh - 1) / bh * bh) *
                           double((N + bn - 1) / bn * bn);
      // prefer wide rows (coalesced stores / fewer TMA rows) on ties; keep bw >= 8 when W allows
      const double score = cover * (1.0 + 0.001 * (7 - lw)) * ((bw < 8 && W >= 8) ? 1.05 : 1.0);
      if (score < best) { best = score; b_w = lw; b_h = lh; b_n = ln; }
    }
  *bw_l = b_w; *bh_l = b_h; *bn_l = b_n;
}

template <int BLOCK_N>
int launch(b2o_ctx* ctx, const CUtensorMap& amap, const ConvLayer& L, const TcParams& p, cudaStream_t st) {
  using Cfg = TcConfig<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    B2O_CUDA_CHECK(ctx, cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES));
    configured = true;
  }
  const int grid = p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->profile) {
    B2O_CUDA_CHECK(ctx, cudaEventCreate(&e0));
    B2O_CUDA_CHECK(ctx, cudaEventCreate(&e1));
    B2O_CUDA_CHECK(ctx, cudaEventRecord(e0, st));
  }
  conv_tc_kernel<BLOCK_N><<<grid, NUM_THREADS, Cfg::SMEM_BYTES, st>>>(amap, L.wmap, p);
  B2O_LAUNCH_CHECK(ctx);
  if (ctx->profile) {
    B2O_CUDA_CHECK(ctx, cudaEventRecord(e1, st));
    ctx->prof_events.push_back(e0);
    ctx->prof_events.push_back(e1);
    // algorithmic FLOPs of this launch: 2 * pixels * (taps * cin) * cout
    ctx->prof_flop += 2.0 * double(p.N) * p.H * p.W * double(p.ksize * p.ksize) * p.cin * p.cout;
  }
  return B2O_OK;
}

}  // namespace

int conv_tc_prepare(b2o_ctx* ctx, ConvLayer& L) {
  L.block_n = 0;
  if (L.cin % BLOCK_K != 0 || L.cout % 16 != 0) return B2O_OK;   // handled by the SIMT engine
  int bn = 256;
  while (bn > 16 && (L.cout % bn != 0)) bn >>= 1;
  if (L.cout % bn != 0) return B2O_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) { ctx->set_error("cuTensorMapEncodeTiled entry point not available"); return B2O_ERR_CUDA; }
  const cuuint64_t ktot = static_cast<cuuint64_t>(L.ksize) * L.ksize * L.cin;
  cuuint64_t dims[2] = {ktot, static_cast<cuuint64_t>(L.cout)};
  cuuint64_t strides[1] = {ktot * 2};
  cuuint32_t box[2] = {BLOCK_K, static_cast<cuuint32_t>(bn)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&L.wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, L.w_kmajor, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled(weights " + L.name + ") failed: " + std::to_string(static_cast<int>(r)));
    return B2O_ERR_CUDA;
  }
  L.block_n = bn;
  return B2O_OK;
}

int conv_tc_run(b2o_ctx* ctx, const ConvLayer& L, const TensorView& in, const TensorView& out, int out_f32,
                cudaStream_t st) {
  if (L.block_n == 0) { ctx->set_error("conv_tc_run: layer " + L.name + " not eligible"); return B2O_ERR_ARG; }
  if (in.c != L.cin || out.c != L.cout || in.n != out.n || in.h != out.h || in.w != out.w) {
    ctx->set_error("conv_tc_run: shape mismatch in " + L.name);
    return B2O_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(in.ptr) & 15) || (in.ld % 8) || (reinterpret_cast<uintptr_t>(out.ptr) & 15) ||
      (out.ld % (out_f32 ? 4 : 8))) {
    ctx->set_error("conv_tc_run: misaligned view in " + L.name);
    return B2O_ERR_ARG;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { ctx->set_error("cuTensorMapEncodeTiled entry point not available"); return B2O_ERR_CUDA; }

  TcParams p;
  p.N = in.n; p.H = in.h; p.W = in.w;
  p.cin = L.cin; p.cout = L.cout; p.ksize = L.ksize; p.dil = L.dil;
  pick_box(in.n, in.h, in.w, &p.bw_log2, &p.bh_log2, &p.bn_log2);
  p.tiles_w = (in.w + (1 << p.bw_log2) - 1) >> p.bw_log2;
  p.tiles_h = (in.h + (1 << p.bh_log2) - 1) >> p.bh_log2;
  p.tiles_n = (in.n + (1 << p.bn_log2) - 1) >> p.bn_log2;
  p.n_tiles = L.cout / L.block_n;
  const long long total = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_n * p.n_tiles;
  if (total > 0x7fffffffLL) { ctx->set_error("conv_tc_run: too many tiles"); return B2O_ERR_ARG; }
  p.total_tiles = static_cast<int>(total);
  p.s1 = L.s1; p.t1 = L.t1; p.s2 = L.s2; p.t2 = L.t2; p.relu = L.relu;
  p.out = out.ptr; p.out_ld = out.ld; p.out_f32 = out_f32;

  CUtensorMap amap;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(in.c), static_cast<cuuint64_t>(in.w), static_cast<cuuint64_t>(in.h),
                        static_cast<cuuint64_t>(in.n)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(in.ld) * 2, static_cast<cuuint64_t>(in.ld) * 2 * in.w,
                           static_cast<cuuint64_t>(in.ld) * 2 * in.w * in.h};
  cuuint32_t box[4] = {BLOCK_K, 1u << p.bw_log2, 1u << p.bh_log2, 1u << p.bn_log2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, in.ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->set_error("cuTensorMapEncodeTiled(activation for " + L.name + ") failed: " +
                   std::to_string(static_cast<int>(r)));
    return B2O_ERR_CUDA;
  }
  switch (L.block_n) {
    case 16: return launch<16>(ctx, amap, L, p, st);
    case 32: return launch<32>(ctx, amap, L, p, st);
    case 64: return launch<64>(ctx, amap, L, p, st);
    case 128: return launch<128>(ctx, amap, L, p, st);
    case 256: return launch<256>(ctx, amap, L, p, st);
  }
  ctx->set_error("conv_tc_run: bad block_n");
  return B2O_ERR_ARG;
}
