// tcgen05 GEMM for sm_100a:  C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual), bf16 in, fp32 accumulate.
//
//   * A (activations) and W (nn.Linear weight) are both K-major bf16, fetched by TMA
//     (cp.async.bulk.tensor.2d, 128-byte swizzle) into a multi-stage shared-memory ring;
//   * one elected thread issues tcgen05.mma (cta_group::1, kind::f16, M=128 x N=BN x K=16) with the
//     accumulator tile living in tensor memory (BN fp32 columns x 128 lanes);
//   * four epilogue warps read the accumulator back with tcgen05.ld (32x32b: one row per thread), fuse bias,
//     ReLU, residual add and the fp32 / bf16 down-conversion, stage 32 x 128-byte sub-tiles in swizzled
//     shared memory and write them with TMA stores (cp.async.bulk.tensor, M / N tails clipped in hardware).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (TMEM lane quadrant = warp % 4; the two warps of a quadrant take alternate column chunks,
// which halves the serial TMEM-load -> math -> staging -> TMA-store chain per tile).
// Pipelines: full[s]/empty[s] mbarriers between TMA and MMA, one tmem_full mbarrier MMA -> epilogue.
// Persistent CTAs (one per SM) with a double-buffered TMEM accumulator: the epilogue of one tile overlaps
// the main loop of the next.
#pragma once
#include <cuda.h>
#include <unordered_map>
#include "common.cuh"

namespace bofi {
namespace tc {

constexpr int kBM = 128;   // UMMA M (TMEM lanes)
constexpr int kBK = 64;    // 64 bf16 = one 128-byte swizzle row
constexpr int kUK = 16;    // UMMA K for 16-bit inputs
constexpr int kThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a mis-programmed pipeline traps (error surfaces on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
// ---- optional in-kernel stall accounting (build with -DBOFI_GEMM_PROF, `python -m boficap_b200.build --prof`) ----
// Per shape class (see prof_bucket) the roles add up where their cycles go: the MMA thread's waits on operands
// (full) and on the epilogue (tmem_empty), the TMA thread's waits on free stages, one epilogue warp's waits on the
// accumulator and on its staging tile.  Compiled out of the product library.
#ifdef BOFI_GEMM_PROF
__device__ unsigned long long g_gemm_prof[8][12];
__device__ __forceinline__ int prof_bucket(int N, int K, bool resid) {
  if (K == 512 && N == 2048) return 0;      // FFN1
  if (K == 2048) return 1;                  // FFN2
  if (K == 512 && N == 512) return resid ? 2 : 3;   // O projection / single projection
  if (K == 512 && (N == 1536 || N == 1024)) return 4;   // fused Q|K|V, K|V
  if (K == 512 && N > 4096) return 5;       // vocabulary
  return 6;
}
#define PROF_DECL(...) long long __VA_ARGS__
#define PROF_T() clock64()
#define PROF_WAIT(acc, stmt) do { const long long _t = clock64(); stmt; acc += clock64() - _t; } while (0)
#define PROF_ADD(b, i, v) atomicAdd(&g_gemm_prof[b][i], (unsigned long long)(v))
#else
#define PROF_DECL(...)
#define PROF_WAIT(acc, stmt) do { stmt; } while (0)
#endif

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile (rows of 64 bf16, 8-row groups 1024 B apart):
// start address >> 4 | LBO (ignored for swizzled K-major) = 1 | SBO = 1024 B | version 1 | SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major, 128-byte-swizzled operand tile: the source matrix is stored [contraction][MN] (MN contiguous), a TMA box of
// 64 MN-elements x 64 contraction rows lands as 64 rows of 128 B (8-row swizzle atoms 1024 B apart = SBO), and the
// 64-element MN blocks of a tile are 8192 B apart (= LBO).  Canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in
// 16-byte units (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t make_sw128_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(8192u >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, N >> 3, M >> 4; bits 15 / 16 = A / B is MN-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, int STAGES, int KPS = 1>
struct SmemLayout {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = KPS * (kABytes + kBBytes);   // KPS k-blocks per stage: [A kb .. A kb+KPS-1 | W kb .. W kb+KPS-1]
  static constexpr int kOutOffset = STAGES * kStageBytes;          // epilogue staging: 8 warps x 4 KB
  static constexpr int kOutBytes = 8 * 4096;
  static constexpr int kBarOffset = kOutOffset + kOutBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;   // barriers + alignment slack
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x, +gridDim.x, ... (N fastest, so
// the CTAs resident at one time share A panels through L2).  Two accumulator tiles live in TMEM
// (2 x BN columns): the epilogue of tile i overlaps the main loop of tile i+1.
// A_MN / B_MN: that operand is read from a [K, M] / [K, N] matrix (the contraction index is the ROW of the source):
// weight gradients dW = dY^T . X (both MN-major) and input gradients dX = dY . W (B MN-major) need no transposed copies.
// REDUCE (fp32 output only): the K range is cut into `k_splits` slices handled by different CTAs and every slice is
// ADDED to C with a TMA reduce store (cp.reduce.async.bulk.tensor ... .add): weight gradients have few output tiles and
// a very long contraction, and accumulate into the gradient buffer anyway.
// KPS > 1 (K-major operands, no K split): a ring stage carries KPS consecutive k-blocks, brought by ONE 3-D box per operand
// (cached_tmap_kblocks).  tools/tma_feed_bench.cu: a stage costs ~830 clk of fixed, non-overlapping latency whatever it
// carries, so the 8 (K = 512) or 32 (K = 2048) one-k-block stages of a bounding-loop GEMM -- a CTA does a single tile there,
// nothing hides the fill -- cost 4-17 us; with four k-blocks per stage it is a quarter of the rounds.
template <int BN, int STAGES, typename TOut, bool RELU, bool RESID, bool A_MN = false, bool B_MN = false, bool REDUCE = false, int KPS = 1>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, const float* residual, int ldr,
               int M, int N, int K, int relu, const int* live_rows, int k_splits, const int* rows_dev) {
  pdl_launch();
  static_assert(KPS == 1 || (!A_MN && !B_MN), "several k-blocks per stage: K-major operands only");
  using L = SmemLayout<BN, STAGES, KPS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* full_bar = bars;                      // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;   // [2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk_all = (K + kBK - 1) / kBK;      // a K tail is zero-filled by TMA (out-of-bounds box elements)
  const int nk_per = (REDUCE && KPS == 1) ? (nk_all + k_splits - 1) / k_splits : nk_all;   // (KPS > 1: launched with k_splits == 1)
  const int tiles_n = (N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(smem_u32(&tmem_full_bar[a]), 1);
        mbar_init(smem_u32(&tmem_empty_bar[a]), 8);   // one arrival per epilogue warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the previous kernel; global
  // memory is only touched after the programmatic-dependency wait.  A dead bounding step walks zero tiles.
  pdl_wait();
  // rows_dev: the number of valid A rows lives on the device (compacted SAIC step); tiles beyond it are skipped
  const int m_eff = rows_dev ? min(M, *rows_dev) : M;
  const int ntiles_mn = ((m_eff + kBM - 1) / kBM) * tiles_n;
  const int ntiles = step_is_dead(live_rows) ? 0 : ntiles_mn * (REDUCE ? k_splits : 1);   // work items: (tile, K slice)

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      PROF_DECL(w_slot = 0);
      for (int item = blockIdx.x; item < ntiles; item += gridDim.x) {
        const int tile = item % ntiles_mn, kb0 = (item / ntiles_mn) * nk_per, kb1 = min(nk_all, kb0 + nk_per);
        const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * BN;
        for (int kb = kb0; kb < kb1; kb += KPS, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          PROF_WAIT(w_slot, mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1));
          const uint32_t fb = smem_u32(&full_bar[s]);
          mbar_expect_tx(fb, L::kStageBytes);
          const uint32_t a_dst = smem_u32(smem + s * L::kStageBytes);
          if constexpr (KPS > 1) {          // one 3-D box per operand: KPS consecutive 128B-swizzled k-block tiles
            tma_load_3d(a_dst, &tmA, fb, 0, m0, kb);
            tma_load_3d(a_dst + KPS * L::kABytes, &tmB, fb, 0, n0, kb);
            continue;
          }
          if constexpr (!A_MN) {
            tma_load_2d(a_dst, &tmA, fb, kb * kBK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBM / 64; ++i) tma_load_2d(a_dst + i * 8192, &tmA, fb, m0 + 64 * i, kb * kBK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(a_dst + L::kABytes, &tmB, fb, kb * kBK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(a_dst + L::kABytes + i * 8192, &tmB, fb, n0 + 64 * i, kb * kBK);
          }
        }
      }
#ifdef BOFI_GEMM_PROF
      if (ntiles > (int)blockIdx.x) PROF_ADD(prof_bucket(N, K, RESID), 3, w_slot);
#endif
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, BN, A_MN, B_MN);
      int it = 0, t = 0;
      PROF_DECL(w_acc = 0, w_full = 0, t_begin = PROF_T());
      for (int item = blockIdx.x; item < ntiles; item += gridDim.x, ++t) {
        const int kb0 = (item / ntiles_mn) * nk_per, kb1 = min(nk_all, kb0 + nk_per);
        const int as = t & 1;
        const uint32_t aph = (t >> 1) & 1;
        PROF_WAIT(w_acc, mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1));     // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; kb += KPS, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          PROF_WAIT(w_full, mbar_wait(smem_u32(&full_bar[s]), ph));
          tcgen05_fence_after();
          const uint32_t s_addr = smem_u32(smem + s * L::kStageBytes);
#pragma unroll
          for (int kk = 0; kk < KPS; ++kk) {
            const uint32_t a_addr = s_addr + kk * L::kABytes, b_addr = s_addr + KPS * L::kABytes + kk * L::kBBytes;
            const uint64_t adesc = A_MN ? make_sw128_desc_mn(a_addr) : make_sw128_desc(a_addr);
            const uint64_t bdesc = B_MN ? make_sw128_desc_mn(b_addr) : make_sw128_desc(b_addr);
            // per K step of 16: K-major +32 bytes inside the 128-byte swizzle row; MN-major +16 rows of 128 bytes (encoded >> 4)
            constexpr uint64_t a_step = A_MN ? 128 : 2, b_step = B_MN ? 128 : 2;
#pragma unroll
            for (int k = 0; k < kBK / kUK; ++k) {
              umma_bf16(tmem_d, adesc + a_step * k, bdesc + b_step * k, idesc, ((kb - kb0) | kk | k) != 0);
            }
          }
          umma_commit(smem_u32(&empty_bar[s]));      // frees the smem stage when these MMAs retire
        }
        umma_commit(smem_u32(&tmem_full_bar[as]));   // accumulator complete
      }
#ifdef BOFI_GEMM_PROF
      if (t > 0) {
        const int b = prof_bucket(N, K, RESID);
        PROF_ADD(b, 0, PROF_T() - t_begin); PROF_ADD(b, 1, w_acc); PROF_ADD(b, 2, w_full); PROF_ADD(b, 7, 1);
      }
#endif
    }
  } else {
    // Epilogue: TMEM -> registers (one accumulator row per thread) -> bias / ReLU / residual -> swizzled smem
    // staging tile (32 rows x 128 B) -> TMA store.  Full-line coalesced writes, M / N tails clipped by the
    // tensor map.  Eight warps: warp `quad + 4*half + 2` owns TMEM lanes [32*quad, +32) and the column chunks
    // c = half, half + 2, ...
    constexpr int CC = 128 / (int)sizeof(TOut);      // output columns per 128-byte staging row
    const int quad = warp & 3;                       // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - 2) >> 2;
    const uint32_t sbuf = smem_u32(smem + L::kOutOffset + (warp - 2) * 4096);
    const uint32_t srow = sbuf + (uint32_t)lane * 128u;
    int t = 0;
    PROF_DECL(w_tfull = 0, w_stage = 0, w_tld = 0, t_math = 0, t_store = 0, t_arrive = 0, t_begin = PROF_T());
    for (int item = blockIdx.x; item < ntiles; item += gridDim.x, ++t) {
      const int tile = item % ntiles_mn;
      const int m0 = (tile / tiles_n) * kBM, n0 = (tile % tiles_n) * BN;
      const int as = t & 1;
      const uint32_t aph = (t >> 1) & 1;
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < M;
      constexpr int NC = BN / CC;
      constexpr int NI = (NC + 1) / 2;                          // chunks per warp
      // The residual does not depend on the accumulator: fetch this warp's first chunk before waiting for the MMAs and
      // keep the next one in flight while a chunk drains.  Loads are COALESCED (lane l reads 16 bytes of row
      // j*4 + l/8 at column chunk l%8: every warp instruction covers four whole 128-byte lines) and are transposed to
      // the row-per-thread layout through the warp's swizzled staging tile.
      constexpr int RV = RESID ? 8 : 1;                         // residual only exists on the fp32-output path
      float4 res[2][RV];
      auto fetch_res = [&](float4 (&dst)[RV], int c) {
        if constexpr (RESID) {
          const int n = n0 + c * CC;
          if (c < NC && n + CC <= N) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int rr = m0 + quad * 32 + j * 4 + (lane >> 3);
              dst[j] = (rr < M) ? *reinterpret_cast<const float4*>(residual + (size_t)rr * ldr + n + 4 * (lane & 7))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
      };
      // Bias of this warp's NI chunks (NI * CC = 128 values): ONE float4 per lane, fetched before the accumulator wait and
      // handed out by shuffles.  The L2 is saturated by the operand stream (tools/gemm_stalls.py), so a load issued
      // inside the drain would put microseconds of loaded L2 latency on every chunk.
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      {
        const int ci = (lane * 4) / CC, cc = half + 2 * ci;
        const int nb = n0 + cc * CC + (lane * 4) % CC;
        if (ci < NI && cc < NC && nb + 4 <= N) bv = *reinterpret_cast<const float4*>(bias + nb);
      }
      fetch_res(res[0], half);
      PROF_WAIT(w_tfull, mbar_wait(smem_u32(&tmem_full_bar[as]), aph));
      tcgen05_fence_after();
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = half + 2 * i;
        const int n = n0 + c * CC;
        if (c >= NC || n >= N) break;                 // warp-uniform
        fetch_res(res[(i + 1) & 1], c + 2);
        const bool full = (n + CC <= N);              // warp-uniform
        uint32_t r[CC];
        PROF_DECL(t_ld0 = PROF_T());
        {
          uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
          tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + c * CC), r0);
          if constexpr (CC == 64) {
            uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + c * CC + 32), r1);
          }
        }
#ifdef BOFI_GEMM_PROF
        w_tld += PROF_T() - t_ld0;
        const long long t_m0 = PROF_T();
#endif
        if constexpr (RESID) {
          if (full) {
            // transpose the coalesced residual registers into this thread's row through the staging tile
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int rl = j * 4 + (lane >> 3);
              const float4 v = res[i & 1][j];
              st_shared_v4(sbuf + (uint32_t)rl * 128u + (uint32_t)(((lane & 7) ^ (rl & 7)) * 16), __float_as_uint(v.x), __float_as_uint(v.y),
                           __float_as_uint(v.z), __float_as_uint(v.w));
            }
            __syncwarp();
          }
        }
        if (full) {
#pragma unroll
          for (int j = 0; j < CC; j += 4) {
            const int src = i * (CC / 4) + j / 4;          // the lane holding the bias of these four columns
            float4 x = make_float4(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bv.x, src),
                                   __uint_as_float(r[j + 1]) + __shfl_sync(0xffffffffu, bv.y, src),
                                   __uint_as_float(r[j + 2]) + __shfl_sync(0xffffffffu, bv.z, src),
                                   __uint_as_float(r[j + 3]) + __shfl_sync(0xffffffffu, bv.w, src));
            if constexpr (RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            if constexpr (RESID) {
              const float4 rr = ld_shared_v4(srow + (uint32_t)(((j / 4) ^ (lane & 7)) * 16));
              x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
            }
            r[j] = __float_as_uint(x.x); r[j + 1] = __float_as_uint(x.y);
            r[j + 2] = __float_as_uint(x.z); r[j + 3] = __float_as_uint(x.w);
          }
        } else {
          // N tail (only the last column tile of the vocab projection): guarded scalar path
#pragma unroll
          for (int j = 0; j < CC; ++j) {
            float x = 0.f;
            if (n + j < N) {
              x = __uint_as_float(r[j]) + bias[n + j];
              if constexpr (RELU) x = fmaxf(x, 0.f);
              if constexpr (RESID) { if (row_ok) x += residual[(size_t)row * ldr + n + j]; }
            }
            r[j] = __float_as_uint(x);
          }
        }
        // this warp's staging tile must have been read out by the TMA engine (its previous store); on the residual path
        // every lane must also be done reading its residual row before the tile is overwritten
        PROF_WAIT(w_stage, if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); __syncwarp());
        if constexpr (CC == 32) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(srow + (uint32_t)((j ^ (lane & 7)) * 16), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(srow + (uint32_t)((j ^ (lane & 7)) * 16),
                         pack_bf16(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1])),
                         pack_bf16(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])),
                         pack_bf16(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])),
                         pack_bf16(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])));
        }
#ifdef BOFI_GEMM_PROF
        t_math += PROF_T() - t_m0;
        const long long t_s0 = PROF_T();
#endif
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if constexpr (REDUCE) tma_reduce_add_2d(&tmC, sbuf, n, m0 + quad * 32);
          else tma_store_2d(&tmC, sbuf, n, m0 + quad * 32);
        }
#ifdef BOFI_GEMM_PROF
        t_store += PROF_T() - t_s0;
#endif
      }
      PROF_DECL(t_a0 = PROF_T());
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
#ifdef BOFI_GEMM_PROF
      t_arrive += PROF_T() - t_a0;
#endif
    }
#ifdef BOFI_GEMM_PROF
    if (warp == 2 && lane == 0 && t > 0) {
      const int b = prof_bucket(N, K, RESID);
      PROF_ADD(b, 4, PROF_T() - t_begin); PROF_ADD(b, 5, w_tfull); PROF_ADD(b, 6, w_stage);
      PROF_ADD(b, 8, w_tld); PROF_ADD(b, 9, t_math); PROF_ADD(b, 10, t_store); PROF_ADD(b, 11, t_arrive);
    }
#endif
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D tensor [rows, cols] of 2-byte (bf16) or 4-byte (fp32) elements with row pitch ld (elements);
// box = 128 bytes of columns x box_rows, 128 B swizzle.
inline bool make_tmap(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, int elt) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * (uint64_t)elt};
  cuuint32_t box[2] = {(cuuint32_t)(128 / elt), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, elt == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

inline int num_sms() {
  static PerDevice<int> count;
  int& n = count.get();
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int STAGES, typename TOut, bool RELU, bool RESID, bool A_MN = false, bool B_MN = false, bool REDUCE = false, int KPS = 1>
inline cudaError_t launch(cudaStream_t s, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                          const float* bias, const float* residual, int ldr, int M, int N, int K, int relu,
                          const int* live_rows, int k_splits = 1, const int* rows_dev = nullptr) {
  using L = SmemLayout<BN, STAGES, KPS>;
  static_assert(L::kTotal <= 232448, "shared memory budget");
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, TOut, RELU, RESID, A_MN, B_MN, REDUCE, KPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int ntiles = ((M + kBM - 1) / kBM) * ((N + BN - 1) / BN) * k_splits;
  const int grid = ntiles < num_sms() ? ntiles : num_sms();
  launch_k(gemm_tc_kernel<BN, STAGES, TOut, RELU, RESID, A_MN, B_MN, REDUCE, KPS>, grid, kThreads, L::kTotal, s, tmA, tmB, tmC, bias, residual, ldr, M, N, K, relu, live_rows, k_splits, rows_dev);
  return cudaGetLastError();
}

// Tensor maps are pure functions of (pointer, shape, pitch, box): cache them (weights and workspace
// buffers are stable across steps), so steady-state launches make no driver call.
struct TmapKey {
  const void* p; uint64_t rows, cols, ld; uint32_t box; int elt;
  bool operator==(const TmapKey& o) const {
    return p == o.p && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box && elt == o.elt;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.p);
    h ^= k.rows * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.cols * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (k.ld * 31 + k.box * 7 + (uint64_t)k.elt) + (h << 6) + (h >> 2);
    return h;
  }
};
inline const CUtensorMap* cached_tmap(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, int elt = 2) {
  // Callers hold the returned pointers across the next two lookups: on overflow the full map is parked in `retired`
  // (node addresses survive the move) and only freed one overflow later, never inside a lookup sequence.
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapHash> cache, retired;
  TmapKey key{ptr, rows, cols, ld, box_rows, elt};
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  if (cache.size() > 8192) { retired = std::move(cache); cache.clear(); }
  CUtensorMap tm;
  if (!make_tmap(&tm, ptr, rows, cols, ld, box_rows, elt)) return nullptr;
  return &cache.emplace(key, tm).first->second;
}

// BOFI_RESID_RED=0: in-place residual adds go back through the register epilogue (A/B measurements)
inline bool inplace_reduce() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BOFI_RESID_RED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// K-major bf16 operand [rows, cols] (pitch ld) seen as the 3-D tensor (64 k, rows, cols/64 k-blocks): one box
// (64, box_rows, box_kb) lands as box_kb consecutive 128B-swizzled k-block tiles of box_rows x 128 bytes -- the smem
// image of box_kb separate 2-D loads, in one instruction (see Smem2T in gemm_tc2.cuh for why that matters).
inline const CUtensorMap* cached_tmap_kblocks(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_kb) {
  static thread_local std::unordered_map<TmapKey, CUtensorMap, TmapHash> cache, retired;   // see cached_tmap
  TmapKey key{ptr, rows, cols, ld, box_rows, (int)box_kb};
  auto it = cache.find(key);
  if (it != cache.end()) return &it->second;
  if (cache.size() > 8192) { retired = std::move(cache); cache.clear(); }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn || cols % 64 != 0) return nullptr;
  CUtensorMap tm;
  cuuint64_t gdim[3] = {64, rows, cols / 64};
  cuuint64_t gstride[2] = {ld * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_kb};
  cuuint32_t estr[3] = {1, 1, 1};
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return nullptr;
  return &cache.emplace(key, tm).first->second;
}

inline bool small_kps4() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BOFI_KPS_SMALL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// A [M,K] bf16 (pitch lda), W [N,K] bf16 (pitch ldw).  Requires K % 64 == 0, 16-byte aligned pitches.
template <typename TOut>
inline cudaError_t gemm_tc(cudaStream_t s, const bf16* A, int lda, const bf16* W, int ldw, const float* bias,
                           const float* residual, int ldr, TOut* C, int ldc, int M, int N, int K, int relu,
                           const int* live_rows, const int* rows_dev = nullptr) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (lda % 8 != 0 || ldw % 8 != 0 || (ldc * sizeof(TOut)) % 16 != 0 || (residual && ldr % 4 != 0))
    return cudaErrorInvalidValue;
  const int tiles256 = ((M + kBM - 1) / kBM) * ((N + 255) / 256);
  const bool wide = tiles256 >= num_sms() / 2;      // small problems: narrower tiles spread over more SMs
  // narrow tiles, whole multiples of four k-blocks: four k-blocks per ring stage (see the kernel's KPS note; BOFI_KPS_SMALL=0: one)
  const bool kps4 = !wide && small_kps4() && K % (4 * kBK) == 0;
  const CUtensorMap* tmA = kps4 ? cached_tmap_kblocks(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, 4)
                                : cached_tmap(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM);
  const CUtensorMap* tmB = kps4 ? cached_tmap_kblocks(W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, 64, 4)
                                : cached_tmap(W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, wide ? 256 : 64);
  const CUtensorMap* tmC = cached_tmap(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, (int)sizeof(TOut));
  if (!tmA || !tmB || !tmC) return cudaErrorInvalidValue;
  if (!bias) return cudaErrorInvalidValue;          // every nn.Linear of this model has a bias
#define BOFI_TC_LAUNCH(BN_, ST_, RELU_, RESID_) \
  (kps4 && BN_ == 64 ? launch<64, 2, TOut, RELU_, RESID_, false, false, false, 4>(s, *tmA, *tmB, *tmC, bias, residual, ldr, M, N, K, relu, live_rows, 1, rows_dev) \
                     : launch<BN_, ST_, TOut, RELU_, RESID_>(s, *tmA, *tmB, *tmC, bias, residual, ldr, M, N, K, relu, live_rows, 1, rows_dev))
  if constexpr (sizeof(TOut) == 4) {
    // fp32 outputs: plain (logits), +ReLU (att_embed), +residual (O-proj / FFN2 into the residual stream)
    if (residual && relu) return cudaErrorInvalidValue;
    // In-place residual (x += a . W^T + b, the O-projections and FFN2 of the decode path): the add is done by the L2 with
    // TMA reduce stores, so the SMs never read x -- no residual loads behind the saturated operand stream, no
    // transposes through the staging tile.  (acc + b) + x and x + (acc + b) are the same fp32 sum.
    if (residual && inplace_reduce() && (const void*)residual == (const void*)C && ldr == ldc)
      return wide ? launch<256, 4, TOut, false, false, false, false, true>(s, *tmA, *tmB, *tmC, bias, nullptr, 0, M, N, K, 0, live_rows, 1, rows_dev)
           : kps4 ? launch<64, 2, TOut, false, false, false, false, true, 4>(s, *tmA, *tmB, *tmC, bias, nullptr, 0, M, N, K, 0, live_rows, 1, rows_dev)
                  : launch<64, 6, TOut, false, false, false, false, true>(s, *tmA, *tmB, *tmC, bias, nullptr, 0, M, N, K, 0, live_rows, 1, rows_dev);
    if (residual) return wide ? BOFI_TC_LAUNCH(256, 4, false, true) : BOFI_TC_LAUNCH(64, 6, false, true);
    if (relu) return wide ? BOFI_TC_LAUNCH(256, 4, true, false) : BOFI_TC_LAUNCH(64, 6, true, false);
    return wide ? BOFI_TC_LAUNCH(256, 4, false, false) : BOFI_TC_LAUNCH(64, 6, false, false);
  } else {
    if (residual) return cudaErrorInvalidValue;
    if (relu) return wide ? BOFI_TC_LAUNCH(256, 4, true, false) : BOFI_TC_LAUNCH(64, 6, true, false);
    return wide ? BOFI_TC_LAUNCH(256, 4, false, false) : BOFI_TC_LAUNCH(64, 6, false, false);
  }
#undef BOFI_TC_LAUNCH
}

// Backward-pass contractions without transposed copies (bias must point at >= N zeros):
//   dgrad:  C[M,N] (bf16)  = A[M,K] . Bt[K,N]                 A K-major (pitch lda), Bt = the weight as stored [K,N] (pitch ldb)
//   wgrad:  C[M,N] (fp32) += At[K,M]^T . Bt[K,N]              At = dY as stored [rows, M], Bt = X as stored [rows, N]; residual = C
inline cudaError_t gemm_tc_dgrad(cudaStream_t s, const bf16* A, int lda, const bf16* Bt, int ldb, const float* zero_bias, bf16* C, int ldc,
                                 int M, int N, int K) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (lda % 8 != 0 || ldb % 8 != 0 || ldc % 8 != 0) return cudaErrorInvalidValue;
  const int tiles256 = ((M + kBM - 1) / kBM) * ((N + 255) / 256);
  const bool wide = tiles256 >= num_sms() / 2;
  const CUtensorMap* tmA = cached_tmap(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM);
  const CUtensorMap* tmB = cached_tmap(Bt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64);
  const CUtensorMap* tmC = cached_tmap(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, 2);
  if (!tmA || !tmB || !tmC) return cudaErrorInvalidValue;
  return wide ? launch<256, 4, bf16, false, false, false, true>(s, *tmA, *tmB, *tmC, zero_bias, nullptr, 0, M, N, K, 0, nullptr)
              : launch<64, 6, bf16, false, false, false, true>(s, *tmA, *tmB, *tmC, zero_bias, nullptr, 0, M, N, K, 0, nullptr);
}
inline cudaError_t gemm_tc_wgrad(cudaStream_t s, const bf16* At, int lda, const bf16* Bt, int ldb, const float* zero_bias, float* C, int ldc,
                                 int M, int N, int K) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (lda % 8 != 0 || ldb % 8 != 0 || ldc % 4 != 0) return cudaErrorInvalidValue;
  // few output tiles, long contraction: 256-wide tiles and enough K slices to give every SM about two work items
  const int tiles256 = ((M + kBM - 1) / kBM) * ((N + 255) / 256);
  const int nk = (K + kBK - 1) / kBK;
  int splits = (2 * num_sms()) / tiles256;
  splits = splits < 1 ? 1 : splits;
  if (splits > nk / 4) splits = nk / 4 < 1 ? 1 : nk / 4;
  const int per = (nk + splits - 1) / splits;
  splits = (nk + per - 1) / per;                      // no empty slice
  const CUtensorMap* tmA = cached_tmap(At, (uint64_t)K, (uint64_t)M, (uint64_t)lda, 64);
  const CUtensorMap* tmB = cached_tmap(Bt, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, 64);
  const CUtensorMap* tmC = cached_tmap(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, 4);
  if (!tmA || !tmB || !tmC) return cudaErrorInvalidValue;
  return launch<256, 4, float, false, false, true, true, true>(s, *tmA, *tmB, *tmC, zero_bias, nullptr, 0, M, N, K, 0, nullptr, splits);
}

}  // namespace tc
}  // namespace bofi
