// Residual GEMM with the NEXT LayerNorm in its epilogue (SublayerConnection, TransformerModel.py:1351-1363 followed by the
// LayerNorm of :1338-1349 that opens the next sublayer / closes the stack):
//
//     x[M,512] += A[M,K] . W[512,K]^T + bias          (fp32 residual stream, in place)
//     y[M,512]  = a_2 * (x - mean) / (std_unbiased + 1e-6) + b_2      (bf16 operand of the next GEMM)
//
// One launch instead of the residual GEMM (TMA reduce-add into x) + layernorm_kernel (which re-reads all of x): per row
// 2 KB of fp32 reads and a launch are saved.  Same CTA-pair machinery as gemm_tc2.cuh (cluster 2, tcgen05.mma.cta_group::2,
// 2-SM TMA loads, 3 x 64 KB ring), but a pair owns a whole 256-row x 512-column panel: the two 256-column halves of a row
// panel go to the two TMEM accumulator buffers (2 x 256 = all 512 columns), so every epilogue thread finds its complete row
// in its own SM's tensor memory and the LayerNorm statistics need no exchange between SMs -- only between the two warps
// that share a TMEM lane quadrant (two named-barrier rounds over 1 KB of shared memory).
//
// Epilogue per row panel (8 warps; warp `quad + 4*eh + 2` owns lanes [32*quad, +32) and the 64-column chunks eh, eh+2, ...):
//   pass A  per half, as soon as its accumulator is complete (the other half's MMAs are still running): TMEM -> registers,
//           + bias + residual (coalesced loads, transposed through the warp's staging tile) -> x; x goes back INTO tensor
//           memory (tcgen05.st) and, through the staging tile, out with a TMA store; row sum accumulated.
//   pass B  mean known: sum (x - mean)^2 from tensor memory (two-pass variance, the arithmetic of layernorm_kernel).
//   pass C  normalise from tensor memory, bf16, staging tile, TMA store into y; each accumulator buffer is handed back to
//           the MMA thread as soon as its half is drained, so the next panel's first half overlaps the second half's drain.
// Nothing 512-wide ever sits in registers; x is read from HBM once (the residual) and written once, y written once.
#pragma once
#include "gemm_tc2.cuh"

namespace bofi {
namespace tc {

struct SmemLN {
  using B = Smem2T<false, 2>;
  static constexpr int kStages = B::kStages;                       // 3 stages of 64 KB (two k-blocks of A rows + W half each)
  static constexpr int kStageBytes = B::kStageBytes;
  static constexpr int kOutOffset = kStages * kStageBytes;
  static constexpr int kOutBytes = 8 * 4096;                       // one staging tile per epilogue warp
  static constexpr int kXchOffset = kOutOffset + kOutBytes;        // [2][128] floats: row statistics of the two warps of a quadrant
  static constexpr int kBarOffset = kXchOffset + 1024;
  static constexpr int kTotal = kBarOffset + 512 + 1024;
};
static_assert(SmemLN::kTotal <= 232448, "shared memory budget");

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
gemm_tc2_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const float* __restrict__ bias,
                   const float* residual, int ldr, const float* __restrict__ ln_a, const float* __restrict__ ln_b, int M, int K,
                   const int* live_rows, const int* rows_dev) {
  pdl_launch();
  using L = SmemLN;
  constexpr int BN = k2BN, STAGES = L::kStages, KPS = 2, NH = 2;      // NH column halves = the two accumulator buffers
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* full_bar = bars;                      // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;   // [2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* xch = reinterpret_cast<float*>(smem + L::kXchOffset);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk_all = (K + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(smem_u32(&tmem_full_bar[a]), 1);
        mbar_init(smem_u32(&tmem_empty_bar[a]), 16);  // one arrival per epilogue warp of BOTH CTAs (only the leader's is used)
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int m_eff = rows_dev ? min(M, *rows_dev) : M;
  const int npanels = step_is_dead(live_rows) ? 0 : (m_eff + 2 * kBM - 1) / (2 * kBM);     // 256-row panels, one per CTA pair at a time
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  uint8_t* ring = smem;                            // the operand ring starts the carve-up

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int p = cid; p < npanels; p += ncl) {
        const int m0 = p * 2 * kBM + (int)rank * kBM;
        for (int h = 0; h < NH; ++h) {
          const int n0 = h * BN;
          for (int kb = 0; kb < nk_all; kb += KPS, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            if (rank == 0) mbar_expect_tx(fb, 2 * L::kStageBytes);
            const uint32_t a_dst = smem_u32(ring + s * L::kStageBytes);
            tma_load_3d_2sm(a_dst, &tmA, fb, 0, m0, kb);                                               // this CTA's 128 rows of A, two k-blocks
            tma_load_3d_2sm(a_dst + KPS * L::B::kABytes, &tmB, fb, 0, n0 + (int)rank * (BN / 2), kb);   // this CTA's half of the W half
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBM, BN, false, false);
      int it = 0, t = 0;
      for (int p = cid; p < npanels; p += ncl) {
        for (int h = 0; h < NH; ++h, ++t) {
          const int as = t & 1;                    // == h
          const uint32_t aph = (t >> 1) & 1;
          mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
          for (int kb = 0; kb < nk_all; kb += KPS, ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            mbar_wait(smem_u32(&full_bar[s]), ph);
            tcgen05_fence_after();
            const uint32_t s_addr = smem_u32(ring + s * L::kStageBytes);
#pragma unroll
            for (int kk = 0; kk < KPS; ++kk) {
              const uint64_t adesc = make_sw128_desc(s_addr + kk * L::B::kABytes);
              const uint64_t bdesc = make_sw128_desc(s_addr + KPS * L::B::kABytes + kk * L::B::kBBytes);
#pragma unroll
              for (int k = 0; k < kBK / kUK; ++k) umma_bf16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | kk | k) != 0);
            }
            umma_commit_2sm(smem_u32(&empty_bar[s]));
          }
          umma_commit_2sm(smem_u32(&tmem_full_bar[as]));
        }
      }
    }
  } else {
    const int quad = warp & 3;                       // TMEM lanes [32*quad, 32*quad+32)
    const int eh = (warp - 2) >> 2;                  // which of the quadrant's two warps
    const uint32_t sbuf = smem_u32(smem + L::kOutOffset + (warp - 2) * 4096);
    const uint32_t srow = sbuf + (uint32_t)lane * 128u;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* my_x = xch + eh * 128 + quad * 32 + lane;
    const float* peer_x = xch + (eh ^ 1) * 128 + quad * 32 + lane;
    auto staging_free = [&]() {                      // the TMA engine has read this warp's previous store out of the staging tile
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    };
    int j = 0;                                       // panels this pair has done: accumulator phase
    for (int p = cid; p < npanels; p += ncl, ++j) {
      const int m0 = p * 2 * kBM + (int)rank * kBM;
      const uint32_t aph = (uint32_t)(j & 1);
      // this warp's eight 32-column chunks of the row: g = 4*h + q  ->  column  h*256 + (4*(q>>1) + 2*eh + (q&1)) * 32
      auto chunk_col = [&](int g) { return (g >> 2) * BN + (4 * ((g & 3) >> 1) + 2 * eh + (g & 1)) * 32; };
      float4 res[2][8];
      auto fetch_res = [&](float4 (&dst)[8], int g) {
        const int n = chunk_col(g);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int rr = m0 + quad * 32 + jj * 4 + (lane >> 3);
          dst[jj] = (rr < M) ? *reinterpret_cast<const float4*>(residual + (size_t)rr * ldr + n + 4 * (lane & 7))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      fetch_res(res[0], 0);
      float rsum = 0.f;
      // ---- pass A ---------------------------------------------------------------------------------------------------
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        // bias of this warp's four chunks of the half: one float4 per lane (chunk q = lane >> 3), handed out by shuffles
        const float4 bv = *reinterpret_cast<const float4*>(bias + chunk_col(4 * h + (lane >> 3)) + 4 * (lane & 7));
        mbar_wait(smem_u32(&tmem_full_bar[h]), aph);
        tcgen05_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int g = 4 * h + q;
          const int n = chunk_col(g);
          if (g + 1 < 8) fetch_res(res[(g + 1) & 1], g + 1);
          uint32_t r[32];
          tmem_ld32(lane_base + (uint32_t)n, r);
          staging_free();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {            // coalesced residual registers -> row-per-thread through the staging tile
            const int rl = jj * 4 + (lane >> 3);
            const float4 v = res[g & 1][jj];
            st_shared_v4(sbuf + (uint32_t)rl * 128u + (uint32_t)(((lane & 7) ^ (rl & 7)) * 16), __float_as_uint(v.x), __float_as_uint(v.y),
                         __float_as_uint(v.z), __float_as_uint(v.w));
          }
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const int src = q * 8 + c / 4;
            const float4 rr = ld_shared_v4(srow + (uint32_t)(((c / 4) ^ (lane & 7)) * 16));
            float4 x = make_float4(__uint_as_float(r[c]) + __shfl_sync(0xffffffffu, bv.x, src),
                                   __uint_as_float(r[c + 1]) + __shfl_sync(0xffffffffu, bv.y, src),
                                   __uint_as_float(r[c + 2]) + __shfl_sync(0xffffffffu, bv.z, src),
                                   __uint_as_float(r[c + 3]) + __shfl_sync(0xffffffffu, bv.w, src));
            x.x += rr.x; x.y += rr.y; x.z += rr.z; x.w += rr.w;
            rsum += (x.x + x.y) + (x.z + x.w);
            r[c] = __float_as_uint(x.x); r[c + 1] = __float_as_uint(x.y);
            r[c + 2] = __float_as_uint(x.z); r[c + 3] = __float_as_uint(x.w);
          }
          tmem_st32(lane_base + (uint32_t)n, r);      // x replaces the accumulator: passes B and C read it from tensor memory
          __syncwarp();                               // every lane has read its residual row out of the staging tile
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            st_shared_v4(srow + (uint32_t)((jj ^ (lane & 7)) * 16), r[4 * jj], r[4 * jj + 1], r[4 * jj + 2], r[4 * jj + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmX, sbuf, n, m0 + quad * 32);
        }
      }
      // ---- row mean: this warp's 256 columns + the other warp of the quadrant's ---------------------------------------------
      *my_x = rsum;
      named_bar_sync(1 + quad, 64);
      const float mean = (rsum + *peer_x) * (1.0f / 512.0f);
      named_bar_sync(1 + quad, 64);                   // the peer has read before the slot is reused
      // ---- pass B: sum (x - mean)^2 -----------------------------------------------------------------------------------------
      float ss = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t r[32];
        tmem_ld32(lane_base + (uint32_t)chunk_col(g), r);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float d0 = __uint_as_float(r[c]) - mean, d1 = __uint_as_float(r[c + 1]) - mean;
          s0 = fmaf(d0, d0, s0);
          s1 = fmaf(d1, d1, s1);
        }
        ss += s0 + s1;
      }
      *my_x = ss;
      named_bar_sync(1 + quad, 64);
      const float denom = sqrtf((ss + *peer_x) * (1.0f / 511.0f)) + 1e-6f;
      named_bar_sync(1 + quad, 64);
      const float inv = 1.0f / denom;
      // ---- pass C: normalise, bf16, TMA store; hand each accumulator buffer back as soon as its half is drained -------------
#pragma unroll
      for (int h = 0; h < NH; ++h) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int n = h * BN + (eh + 2 * i) * 64;                 // == chunk_col(4*h + 2*i), 64 columns = two of this warp's chunks
          uint32_t r[64];
          {
            uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
            uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
            tmem_ld32(lane_base + (uint32_t)n, r0);
            tmem_ld32(lane_base + (uint32_t)n + 32u, r1);
          }
#pragma unroll
          for (int c = 0; c < 64; c += 4) {
            const float4 ga = __ldg(reinterpret_cast<const float4*>(ln_a + n + c));
            const float4 gb = __ldg(reinterpret_cast<const float4*>(ln_b + n + c));
            r[c] = __float_as_uint(ln_div(ga.x * (__uint_as_float(r[c]) - mean), denom, inv) + gb.x);
            r[c + 1] = __float_as_uint(ln_div(ga.y * (__uint_as_float(r[c + 1]) - mean), denom, inv) + gb.y);
            r[c + 2] = __float_as_uint(ln_div(ga.z * (__uint_as_float(r[c + 2]) - mean), denom, inv) + gb.z);
            r[c + 3] = __float_as_uint(ln_div(ga.w * (__uint_as_float(r[c + 3]) - mean), denom, inv) + gb.w);
          }
          staging_free();
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            st_shared_v4(srow + (uint32_t)((jj ^ (lane & 7)) * 16),
                         pack_bf16(__uint_as_float(r[8 * jj]), __uint_as_float(r[8 * jj + 1])),
                         pack_bf16(__uint_as_float(r[8 * jj + 2]), __uint_as_float(r[8 * jj + 3])),
                         pack_bf16(__uint_as_float(r[8 * jj + 4]), __uint_as_float(r[8 * jj + 5])),
                         pack_bf16(__uint_as_float(r[8 * jj + 6]), __uint_as_float(r[8 * jj + 7])));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmY, sbuf, n, m0 + quad * 32);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[h]), 0);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}

// x [M, 512] fp32 (pitch ldx) += A [M, K] bf16 . W [512, K]^T + bias;  y [M, 512] bf16 (pitch ldy) = LayerNorm(x; ln_a, ln_b).
// K % 128 == 0.  rows_dev / live_rows as in gemm_tc2.
inline cudaError_t gemm_tc2_ln(cudaStream_t s, const bf16* A, int lda, const bf16* W, int ldw, const float* bias, float* x, int ldx,
                               const float* ln_a, const float* ln_b, bf16* y, int ldy, int M, int K, const int* live_rows, const int* rows_dev) {
  if (M <= 0) return cudaSuccess;
  if (lda % 8 != 0 || ldw % 8 != 0 || ldx % 4 != 0 || ldy % 8 != 0 || !bias || K % (2 * kBK) != 0) return cudaErrorInvalidValue;
  const int N = 2 * k2BN;
  // four lookups: the maps are copied (the cache only guarantees a pointer across the next two lookups)
  const CUtensorMap* p = cached_tmap_kblocks(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, 2);
  if (!p) return cudaErrorInvalidValue;
  const CUtensorMap tmA = *p;
  p = cached_tmap_kblocks(W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, k2BN / 2, 2);
  if (!p) return cudaErrorInvalidValue;
  const CUtensorMap tmB = *p;
  p = cached_tmap(x, (uint64_t)M, (uint64_t)N, (uint64_t)ldx, 32, 4);
  if (!p) return cudaErrorInvalidValue;
  const CUtensorMap tmX = *p;
  p = cached_tmap(y, (uint64_t)M, (uint64_t)N, (uint64_t)ldy, 32, 2);
  if (!p) return cudaErrorInvalidValue;
  const CUtensorMap tmY = *p;
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLN::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int panels = (M + 2 * kBM - 1) / (2 * kBM);
  int clusters = num_sms() / 2;
  if (panels < clusters) clusters = panels;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = SmemLN::kTotal;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, gemm_tc2_ln_kernel, tmA, tmB, tmX, tmY, bias, (const float*)x, ldx, ln_a, ln_b, M, K, live_rows, rows_dev);
}

}  // namespace tc
}  // namespace bofi
