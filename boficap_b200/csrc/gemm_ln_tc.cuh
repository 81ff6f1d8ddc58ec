// LayerNorm-fused tcgen05 GEMM:  C[M,N] = act(LN(x)[M,512] . W[N,512]^T + bias)          (bf16 operands, fp32 accumulate)
//
// Every LayerNorm of the model feeds a projection of the SAME 512 inputs (Q|K|V, cross-attention Q, FFN1, the vocab
// projection).  Instead of a LayerNorm kernel writing a bf16 copy that the GEMM's TMA re-reads once per column tile,
// the CTA that owns a 128-row panel normalises it ONCE -- fp32 rows from the residual stream, the reference's custom
// LayerNorm (TransformerModel.py:1338-1349: unbiased std, eps added to std), two-pass statistics, warp per row --
// and writes the bf16 result straight into shared memory in the 128-byte-swizzled K-major layout the UMMA
// descriptors expect: the whole K = 512 extent of the panel (8 k-blocks x 16 KB) stays resident while the CTA walks
// the column tiles of the panel, so only W is streamed (TMA, 4-stage ring).  Shared-memory fill per MMA flop drops by
// a third compared with re-loading A per tile (these K = 512 GEMMs are L2->SM bandwidth bound), and the LayerNorm
// kernel with its bf16 round trip through HBM disappears.
//
// MEASURED RESULT (B200, encoder of B=1024 x R=36): slower than the separate LayerNorm kernel + streaming GEMM -- 2.29 ms vs
// 1.58 + 0.36 ms for the 12 LayerNorm-fed GEMMs of the encoder.  With 128 KB of shared memory pinned by the resident panel
// only 64 KB of W can be in flight (4 x 16 KB stages; the streaming kernel keeps 192 KB), which starves the MMA pipe
// (the W stream needs ~115 GB/s per SM), and the eight warps that normalise a panel are latency-bound (17 us per panel).
// The path is therefore OFF by default (BOFI_LNFUSE=1 enables it; parity-tested) and kept as the record of the experiment.
//
// Work item = (row panel, chunk of column tiles); small-M calls (the M = batch bounding rows) split a panel's column
// tiles over several CTAs, each normalising the panel for itself (the rows come from L2).
// Warp roles as in gemm_tc.cuh: warp 0 = TMA producer (W tiles), warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = LayerNorm producers, then epilogue (TMEM -> bias / ReLU -> swizzled staging -> TMA store).
#pragma once
#include <cstdlib>
#include "gemm_tc.cuh"

namespace bofi {
namespace tc {

constexpr int kLnBN = 128;            // column tile
constexpr int kLnStages = 4;          // W ring
constexpr int kLnK = 512;             // d_model: the whole contraction is resident
constexpr int kLnKB = kLnK / kBK;     // 8 k-blocks

struct LnSmem {
  static constexpr int kABytes = kLnKB * kBM * 128;                 // 128 KB: [kb][row][64 bf16], SW128
  static constexpr int kBStage = kLnBN * 128;                       // 16 KB
  static constexpr int kBOffset = kABytes;
  static constexpr int kOutOffset = kBOffset + kLnStages * kBStage; // 8 warps x 4 KB staging
  static constexpr int kBarOffset = kOutOffset + 8 * 4096;
  static constexpr int kTotal = kBarOffset + 256 + 1024;
};

template <typename TOut, bool RELU>
__global__ void __launch_bounds__(kThreads, 1)
gemm_ln_tc_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC, const float* __restrict__ x, size_t ldx,
                  const float* __restrict__ ln_a, const float* __restrict__ ln_b, const float* __restrict__ bias, int M, int N,
                  int n_chunks, const int* live_rows, const int* rows_dev, int debug_skip_ln) {
  pdl_launch();
  using L = LnSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* full_bar = bars;                            // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + kLnStages;               // [STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * kLnStages;       // [2]       MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * kLnStages + 2;  // [2]       epilogue -> MMA
  uint64_t* a_full_bar = bars + 2 * kLnStages + 4;      // [1]       LayerNorm warps -> MMA (panel written)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kLnStages + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = (N + kLnBN - 1) / kLnBN;
  const int tpc = (tiles_n + n_chunks - 1) / n_chunks;  // column tiles per chunk

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kLnStages; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(smem_u32(&tmem_full_bar[a]), 1);
        mbar_init(smem_u32(&tmem_empty_bar[a]), 8);
      }
      mbar_init(smem_u32(a_full_bar), 8);               // one arrival per LayerNorm warp
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * kLnBN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const int m_eff = rows_dev ? min(M, *rows_dev) : M;
  const int panels = (m_eff + kBM - 1) / kBM;
  const int nitems = step_is_dead(live_rows) ? 0 : panels * n_chunks;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        const int chunk = item % n_chunks;
        const int t0 = chunk * tpc, t1 = min(tiles_n, t0 + tpc);
        for (int tn = t0; tn < t1; ++tn) {
          for (int kb = 0; kb < kLnKB; ++kb, ++it) {
            const int s = it % kLnStages;
            const uint32_t ph = (it / kLnStages) & 1;
            mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
            const uint32_t fb = smem_u32(&full_bar[s]);
            mbar_expect_tx(fb, L::kBStage);
            tma_load_2d(smem_u32(smem + L::kBOffset + s * L::kBStage), &tmB, fb, kb * kBK, tn * kLnBN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, kLnBN);
      int it = 0, t = 0, ni = 0;
      for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++ni) {
        const int chunk = item % n_chunks;
        const int t0 = chunk * tpc, t1 = min(tiles_n, t0 + tpc);
        mbar_wait(smem_u32(a_full_bar), ni & 1);               // the panel of this item is in shared memory
        tcgen05_fence_after();
        for (int tn = t0; tn < t1; ++tn, ++t) {
          const int as = t & 1;
          const uint32_t aph = (t >> 1) & 1;
          mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(as * kLnBN);
          for (int kb = 0; kb < kLnKB; ++kb, ++it) {
            const int s = it % kLnStages;
            const uint32_t ph = (it / kLnStages) & 1;
            mbar_wait(smem_u32(&full_bar[s]), ph);
            tcgen05_fence_after();
            const uint64_t adesc = make_sw128_desc(smem_u32(smem + kb * (kBM * 128)));
            const uint64_t bdesc = make_sw128_desc(smem_u32(smem + L::kBOffset + s * L::kBStage));
#pragma unroll
            for (int k = 0; k < kBK / kUK; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
            umma_commit(smem_u32(&empty_bar[s]));
          }
          umma_commit(smem_u32(&tmem_full_bar[as]));
        }
      }
    }
  } else {
    constexpr int CC = 128 / (int)sizeof(TOut);
    constexpr int NC = kLnBN / CC;
    constexpr int NI = (NC + 1) / 2;
    const int ew = warp - 2;                           // 0..7
    const int quad = warp & 3, half = ew >> 2;
    const uint32_t sbuf = smem_u32(smem + L::kOutOffset + ew * 4096);
    const uint32_t srow = sbuf + (uint32_t)lane * 128u;
    const uint32_t a_base = smem_u32(smem);
    int t = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int panel = item / n_chunks, chunk = item % n_chunks;
      const int m0 = panel * kBM;
      const int t0 = chunk * tpc, t1 = min(tiles_n, t0 + tpc);
      // ---- LayerNorm of rows [m0 + ew*16, +16) into the resident A panel.  All MMAs that read the previous panel have
      // retired: this warp has waited on tmem_full of the previous item's last tile.  Four rows are in flight at a time
      // (16 independent 16-byte loads per lane, four interleaved reduction chains; 8 warps x 4 rows x 2 KB in flight per SM).
      constexpr int G = 4;
      float4 v[G][4];
      auto fetch = [&](float4 (&dst)[G][4], int g) {
#pragma unroll
        for (int q = 0; q < G; ++q) {
          const int row = m0 + ew * 16 + g * G + q;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[q][i] = (row < m_eff) ? *reinterpret_cast<const float4*>(x + (size_t)row * ldx + (i * 32 + lane) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
#pragma unroll 1
      for (int g = 0; g < (debug_skip_ln ? 0 : 16 / G); ++g) {
        fetch(v, g);
        float mean[G], denom[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
          float sm = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) sm += (v[q][i].x + v[q][i].y) + (v[q][i].z + v[q][i].w);
          mean[q] = sm;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < G; ++q) mean[q] += __shfl_xor_sync(0xffffffffu, mean[q], o);
        }
#pragma unroll
        for (int q = 0; q < G; ++q) {
          mean[q] *= (1.0f / kLnK);
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[q][i].x -= mean[q]; v[q][i].y -= mean[q]; v[q][i].z -= mean[q]; v[q][i].w -= mean[q];
            ss += (v[q][i].x * v[q][i].x + v[q][i].y * v[q][i].y) + (v[q][i].z * v[q][i].z + v[q][i].w * v[q][i].w);
          }
          denom[q] = ss;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int q = 0; q < G; ++q) denom[q] += __shfl_xor_sync(0xffffffffu, denom[q], o);
        }
#pragma unroll
        for (int q = 0; q < G; ++q) {
          const int r = ew * 16 + g * G + q, row = m0 + r;
          const float dn = sqrtf(denom[q] * (1.0f / (kLnK - 1))) + 1e-6f;
          const float inv = 1.0f / dn;            // same arithmetic as layernorm_kernel<bf16>: one reciprocal per row
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = (i * 32 + lane) * 4;
            const float4 ga = *reinterpret_cast<const float4*>(ln_a + c), gb = *reinterpret_cast<const float4*>(ln_b + c);
            float4 o;
            o.x = ln_div(ga.x * v[q][i].x, dn, inv) + gb.x;
            o.y = ln_div(ga.y * v[q][i].y, dn, inv) + gb.y;
            o.z = ln_div(ga.z * v[q][i].z, dn, inv) + gb.z;
            o.w = ln_div(ga.w * v[q][i].w, dn, inv) + gb.w;
            if (row >= m_eff) o = make_float4(0.f, 0.f, 0.f, 0.f);
            // column c: k-block c/64, 16-byte chunk (c%64)/8 swizzled with the row, 8-byte half (c%8)/4
            const uint32_t addr = a_base + (uint32_t)(c >> 6) * (kBM * 128) + (uint32_t)r * 128u +
                                  (uint32_t)((((c & 63) >> 3) ^ (r & 7)) * 16) + (uint32_t)((c & 7) * 2);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(pack_bf16(o.x, o.y)), "r"(pack_bf16(o.z, o.w)) : "memory");
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(a_full_bar));
      // ---- epilogues of the item's column tiles
      for (int tn = t0; tn < t1; ++tn, ++t) {
        const int n0 = tn * kLnBN;
        const int as = t & 1;
        const uint32_t aph = (t >> 1) & 1;
        mbar_wait(smem_u32(&tmem_full_bar[as]), aph);
        tcgen05_fence_after();
#pragma unroll
        for (int i = 0; i < NI; ++i) {
          const int c = half + 2 * i;
          const int n = n0 + c * CC;
          if (c >= NC || n >= N) break;                 // warp-uniform
          const bool full = (n + CC <= N);
          uint32_t r[CC];
          {
            uint32_t(&r0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[0]);
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kLnBN + c * CC), r0);
            if constexpr (CC == 64) {
              uint32_t(&r1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&r[32]);
              tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * kLnBN + c * CC + 32), r1);
            }
          }
          if (full) {
#pragma unroll
            for (int j = 0; j < CC; j += 4) {
              const float4 bb = *reinterpret_cast<const float4*>(bias + n + j);
              float4 xv = make_float4(__uint_as_float(r[j]) + bb.x, __uint_as_float(r[j + 1]) + bb.y, __uint_as_float(r[j + 2]) + bb.z,
                                      __uint_as_float(r[j + 3]) + bb.w);
              if constexpr (RELU) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
              r[j] = __float_as_uint(xv.x); r[j + 1] = __float_as_uint(xv.y); r[j + 2] = __float_as_uint(xv.z); r[j + 3] = __float_as_uint(xv.w);
            }
          } else {
#pragma unroll
            for (int j = 0; j < CC; ++j) {
              float xv = 0.f;
              if (n + j < N) {
                xv = __uint_as_float(r[j]) + bias[n + j];
                if constexpr (RELU) xv = fmaxf(xv, 0.f);
              }
              r[j] = __float_as_uint(xv);
            }
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
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
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tma_store_2d(&tmC, sbuf, n, m0 + quad * 32);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[as]));
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kLnBN) : "memory");
  }
}

template <typename TOut, bool RELU>
inline cudaError_t launch_ln(cudaStream_t s, const CUtensorMap& tmB, const CUtensorMap& tmC, const float* x, size_t ldx, const float* ln_a,
                             const float* ln_b, const float* bias, int M, int N, int n_chunks, int grid, const int* live_rows,
                             const int* rows_dev) {
  static const int debug_skip_ln = getenv("BOFI_DEBUG_SKIP_LN") ? 1 : 0;      // timing experiments only (results are then wrong)
  static PerDevice<bool> configured_dev;   // one flag per (TOut, RELU) instantiation and device
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_tc_kernel<TOut, RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, LnSmem::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  launch_k(gemm_ln_tc_kernel<TOut, RELU>, grid, kThreads, LnSmem::kTotal, s, tmB, tmC, x, ldx, ln_a, ln_b, bias, M, N, n_chunks, live_rows, rows_dev, debug_skip_ln);
  return cudaGetLastError();
}

// x fp32 [M, 512] (pitch ldx), W bf16 [N, 512], C [M, N] (pitch ldc).
template <typename TOut>
inline cudaError_t gemm_ln_tc(cudaStream_t s, const float* x, size_t ldx, const float* ln_a, const float* ln_b, const bf16* W, const float* bias,
                              TOut* C, int ldc, int M, int N, int relu, const int* live_rows, const int* rows_dev) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (ldx % 4 != 0 || (ldc * sizeof(TOut)) % 16 != 0 || !bias) return cudaErrorInvalidValue;
  const CUtensorMap* tmB = cached_tmap(W, (uint64_t)N, (uint64_t)kLnK, (uint64_t)kLnK, kLnBN);
  const CUtensorMap* tmC = cached_tmap(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, (int)sizeof(TOut));
  if (!tmB || !tmC) return cudaErrorInvalidValue;
  const int panels = (M + kBM - 1) / kBM, tiles_n = (N + kLnBN - 1) / kLnBN;
  int n_chunks = num_sms() / panels;                    // small M: spread a panel's column tiles over several CTAs
  n_chunks = n_chunks < 1 ? 1 : (n_chunks > tiles_n ? tiles_n : n_chunks);
  const int tpc = (tiles_n + n_chunks - 1) / n_chunks;
  n_chunks = (tiles_n + tpc - 1) / tpc;                 // no empty chunk
  const int items = panels * n_chunks;
  const int grid = items < num_sms() ? items : num_sms();
  return relu ? launch_ln<TOut, true>(s, *tmB, *tmC, x, ldx, ln_a, ln_b, bias, M, N, n_chunks, grid, live_rows, rows_dev)
              : launch_ln<TOut, false>(s, *tmB, *tmC, x, ldx, ln_a, ln_b, bias, M, N, n_chunks, grid, live_rows, rows_dev);
}

}  // namespace tc
}  // namespace bofi
