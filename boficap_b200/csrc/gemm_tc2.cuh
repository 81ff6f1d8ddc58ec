// 2-CTA tcgen05 GEMM (cta_group::2):  C[M,N] = act(A[M,K] . W[N,K]^T + bias) (+ residual) on 256 x 256 tile pairs.
//
// profiles/r01c_gemm_ncu_summary.txt: the K = 512 projections are bound by shared-memory FILL (13.5 TB/s of L2->SM traffic,
// tensor pipe 58 %): a 128 x 256 tile re-loads 16 KB of A and 32 KB of W per 64-deep k-block.  A CTA pair (two SMs of a
// TPC, cluster 2x1x1) shares the W tile: each CTA loads its own 128 rows of A and only HALF of W (16 + 16 KB), the
// leader's single thread issues tcgen05.mma.cta_group::2 (M = 256) which reads W from both CTAs' shared memory, and
// each SM's tensor core accumulates its 128 rows x 256 columns in its own TMEM.  Fill per MMA flop drops by a third and
// the ring gets two more stages.  Barrier wiring:
//   full[s]        leader only: expect_tx of BOTH CTAs' bytes; both CTAs' TMA loads (.cta_group::2) complete on it
//   empty[s]       per CTA, released by the leader's commit with cluster multicast (frees the stage in both CTAs)
//   tmem_full[a]   per CTA, multicast commit; tmem_empty[a]: leader's barrier, 16 arrivals (8 epilogue warps x 2 CTAs)
// Epilogue as in gemm_tc.cuh (8 warps per CTA drain their own 128 TMEM lanes).  K-major operands only.
//
// MEASURED (B200, B=1024 decode, tools/gemm_stalls.py + bench.py A/B, profiles/r01d_gemm_stalls.txt): numerically identical
// to gemm_tc.cuh.  As first built (5 stages, accumulator handed back with mbarrier.arrive.release.cluster) it was no faster
// than the 1-CTA tile: the epilogue warps spent 32-42 % of their time in that one remote arrive and the MMA thread
// waited for them.  With the default-semantics remote arrive and a 6-stage ring (one staging tile per warp) it is the
// default for M >= 2048, N >= 512: 166.7 k -> 174.6 k captions/s, roofline.frac 0.463 -> 0.486 (BOFI_GEMM2=0: 1-CTA tiles).
// The A-resident mode (ARES) halves the fill again (32 B/clk/SM) but leaves only a 4 x 16 KB ring for W: the MMA thread
// waits for operands just as long (40 %) with a quarter of the bytes in flight -- the slot turnaround (TMA issue ->
// landed -> MMAs retired -> commit -> TMA thread) is ~3 k cycles however few bytes move, so what counts is bytes IN
// FLIGHT, and a resident panel takes them away.  2 % slower end to end; opt-in BOFI_ARES=1.
#pragma once
#include "gemm_tc.cuh"

namespace bofi {
namespace tc {

constexpr int k2BN = 256;
#ifndef BOFI_TC2_STAGES
#define BOFI_TC2_STAGES 6
#endif
constexpr int k2Stages = BOFI_TC2_STAGES;
constexpr int k2PanelKB = 8;                         // A-resident mode: K <= 8 * 64
// ARES (A-resident, K <= 512): the CTA keeps its 128 rows of A for the WHOLE contraction in shared memory (8 k-block
// slices, 128 KB) while it walks the column tiles of that row panel, and streams only its half of W: 16 KB per CTA per
// k-block, 32 B/clk/SM at the tensor pipe's full rate -- below the ~46 B/clk/SM the L2 delivers chip-wide, where the
// streaming tile (32 KB per k-block) is capped at ~72 % and the 1-CTA tile (48 KB) at ~48 % (profiles/r01d_gemm_stalls.txt).
// KPS = k-blocks per ring stage.  tools/tma_feed_bench.cu: one SM's TMA path completes about one STAGE (barrier
// round) per ~1000 clk however many bytes it carries -- 16 KB boxes: 17 B/clk/SM, two of them per stage: 30, one
// 64 KB box: 65-77 -- so a stage must carry as many bytes as possible per instruction.  With KPS = 2 the operands
// are described as 3-D tensors (64 k, rows, K/64 k-blocks) and ONE box (64, 128, 2) brings two consecutive
// 128B-swizzled k-block tiles (32 KB); a stage is 64 KB per CTA and the ring has 3 of them.
template <bool ARES, int KPS = 1>
struct Smem2T {
  static constexpr int kABytes = kBM * kBK * 2;                    // 16 KB: this CTA's rows of A, one k-block
  static constexpr int kBBytes = (k2BN / 2) * kBK * 2;             // 16 KB: this CTA's half of W, one k-block
  static constexpr int kStages = (ARES ? 4 : k2Stages) / KPS;
  static constexpr int kStaging = (ARES || kStages * KPS >= 6) ? 1 : 2;  // staging tiles per epilogue warp (what is left of the 227 KB)
  static constexpr int kPanelBytes = ARES ? k2PanelKB * kABytes : 0;
  static constexpr int kStageBytes = KPS * (ARES ? kBBytes : kABytes + kBBytes);
  static constexpr int kRingOffset = kPanelBytes;
  static constexpr int kOutOffset = kRingOffset + kStages * kStageBytes;
  static constexpr int kOutBytes = 8 * kStaging * 4096;            // staging tiles of the eight epilogue warps
  static constexpr int kBarOffset = kOutOffset + kOutBytes;
  static constexpr int kTotal = kBarOffset + 512 + 1024;
};
static_assert(Smem2T<true>::kTotal <= 232448 && Smem2T<false>::kTotal <= 232448 && Smem2T<false, 2>::kTotal <= 232448, "shared memory budget");

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of the pair; the completion bytes go to the barrier of the pair's CTA 0 (peer bit clear)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
// Default semantics (release at CTA scope), as the accumulator pipelines of the vendor's 2-SM kernels arrive remotely: what
// this arrival publishes is "my tcgen05.ld of the accumulator are done" (tcgen05.wait::ld + fence::before_thread_sync), no
// generic-proxy memory.  With .release.cluster the arrival alone took 32-42 % of an epilogue warp's time (tools/gemm_stalls.py).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}

// Vocabulary-projection epilogues (TransformerModel.logit :1668-1669 + F.log_softmax AttModel.py:206-209 + sample_next_word
// CaptionModel.py:383-431), VMODE != 0.  The [rows, L, V] fp32 logits never exist in HBM:
//   VMODE 1 (statistics)  every epilogue thread owns one row of the tile and folds its 128 columns into a running
//            (max, sum exp(z - max), first index of the max [, sum exp(z - max) z, first NaN, Gumbel-max]) state; one record
//            per (column tile, epilogue half, row) goes to the SoA arrays below, vocab_merge_kernel (kernels.cuh) folds the
//            2 * ceil(N / 256) records of a row into token, max and log-sum-exp.  Nothing else is written.
//   VMODE 2 (log-probs)   only when the caller asked for the [rows, L, V] tensor: the projection is recomputed (K = 512: 0.2
//            ms, cheaper than writing and re-reading 1.5 GB of logits) and  (z - mx[row]) - lse[row]  goes straight to the
//            caller's rows, whose pitch (V = 9491 floats) rules out TMA: a warp transposes its 32 x 32 block through the
//            swizzled staging tile and writes each row segment with one 128-byte coalesced store instruction.
struct VocabEpi {
  float* pm = nullptr;       // [nparts][M] running max
  float* ps = nullptr;       // [nparts][M] sum exp(z - max)
  float* pt = nullptr;       // [nparts][M] sum exp(z - max) * z  (entropy of eval_split; nullptr = not needed)
  int* pi = nullptr;         // [nparts][M] first column holding the max (torch.max: lowest index on ties)
  int* pn = nullptr;         // [nparts][M] first NaN column, 0x7fffffff = none (torch.max: a NaN wins)
  float* pgv = nullptr;      // [nparts][M] sampling: best Gumbel-perturbed score ...
  int* pgi = nullptr;        // [nparts][M] ... its column ...
  float* pgz = nullptr;      // [nparts][M] ... and the unperturbed logit there (log-prob of the sampled token)
  float* pz0 = nullptr;      // [M] logit of column 0 (the token tail padding writes; statistics only, may be nullptr)
  Sampler sp;
  int fast_exp = 0;          // ex2.approx for the sum-exp (bf16 engine without entropy statistics)
  int need_lse = 1;          // 0: greedy tokens only (no log-probs, no statistics): the epilogue skips every exponential
  float* out = nullptr;      // VMODE 2: caller's [M, ldo] tensor
  long long ldo = 0;
  const float* mx = nullptr; // VMODE 2: per-row max and log-sum-exp from vocab_merge_kernel
  const float* lse = nullptr;
  int do_lsm = 0;            // 0: raw logits (output_logsoftmax = 0)
  int tma_out = 0;           // VMODE 2: the rows of `out` are 16-byte aligned (padded pitch): plain TMA stores through tmC
};

// DYN: dynamic tile scheduling through the hardware's cluster launch control (clusterlaunchcontrol.try_cancel).  The grid has
// one cluster per tile; a running pair, instead of walking a static stride of tiles, CANCELS a not-yet-launched cluster and takes
// over its tile (the 16-byte answer is multicast into both CTAs' shared memory, a 2-deep ring of answers run by one extra warp).
// Why: with several batches in flight a GEMM rarely finds all 148 SMs free -- a 64-CTA bounding-loop launch of another batch
// holds some of them for a few microseconds -- and with the static schedule the pairs that start late finish late by exactly that
// much while the others idle; here a late pair simply takes fewer tiles.
constexpr int kThreadsDyn = kThreads + 32;      // + the scheduler warp (warp 10)
template <typename TOut, bool RELU, bool RESID, bool REDUCE = false, bool ARES = false, int KPS = 1, int VMODE = 0, bool DYN = false>
__global__ void __launch_bounds__(DYN ? kThreadsDyn : kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const float* __restrict__ bias, const float* residual, int ldr,
               int M, int N, int K, int relu, const int* live_rows, const int* rows_dev, const VocabEpi ve) {
  pdl_launch();
  using L = Smem2T<ARES, KPS>;
  constexpr int BN = k2BN, STAGES = L::kStages, k2Staging = L::kStaging;
  constexpr bool A_MN = false, B_MN = false;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* full_bar = bars;                      // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]  MMA -> TMA
  uint64_t* tmem_full_bar = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;   // [2]   epilogue -> MMA
  uint64_t* a_full_bar = bars + 2 * STAGES + 4;   // [8]  ARES: A slice kb landed (both CTAs' bytes, leader's barrier)
  uint64_t* a_empty_bar = a_full_bar + k2PanelKB;  // [8]  ARES: the panel's last tile has consumed slice kb
  uint64_t* clc_full_bar = a_empty_bar + k2PanelKB;   // [2]  DYN: the answer of query q has landed (per CTA, 16 tx bytes)
  uint64_t* clc_empty_bar = clc_full_bar + 2;         // [2]  DYN: every consumer of both CTAs has read it (the leader's is used)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(clc_empty_bar + 2);
  uint8_t* clc_resp = smem + L::kBarOffset + 384;     // [2] x 16 B answers
  static_assert(!(ARES && DYN), "the A-resident mode keeps its contiguous static ranges");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk_all = (K + kBK - 1) / kBK;      // a K tail is zero-filled by TMA (out-of-bounds box elements)
  const int nk_per = nk_all;
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
        mbar_init(smem_u32(&tmem_empty_bar[a]), 16);  // one arrival per epilogue warp of BOTH CTAs (only the leader's is used)
      }
      for (int a = 0; a < k2PanelKB; ++a) {
        mbar_init(smem_u32(&a_full_bar[a]), 1);
        mbar_init(smem_u32(&a_empty_bar[a]), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(smem_u32(&clc_full_bar[a]), 1);
        // consumers of an answer: TMA thread + 8 epilogue warps per CTA, the leader's MMA thread and scheduler thread
        mbar_init(smem_u32(&clc_empty_bar[a]), 2 * 9 + 2);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();                              // barriers of both CTAs initialised, TMEM allocated on both SMs
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barriers, TMEM, descriptor prefetch) overlaps the tail of the previous kernel; global
  // memory is only touched after the programmatic-dependency wait.  A dead bounding step walks zero tiles.
  pdl_wait();
  // rows_dev: the number of valid A rows lives on the device (compacted SAIC step); tiles beyond it are skipped
  const int m_eff = rows_dev ? min(M, *rows_dev) : M;
  const int ntiles_mn = ((m_eff + 2 * kBM - 1) / (2 * kBM)) * tiles_n;     // 256-row tile pairs
  const int ntiles = step_is_dead(live_rows) ? 0 : ntiles_mn;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;                  // cluster id / number of clusters
  // Work items of this pair: strided over the pairs, or (ARES) one contiguous range so that the column tiles of a row
  // panel follow each other and the resident A slices are loaded once per panel.
  const int it_begin = ARES ? (int)(((long long)cid * ntiles) / ncl) : cid;
  const int it_end = ARES ? (int)(((long long)(cid + 1) * ntiles) / ncl) : ntiles;
  const int it_step = ARES ? 1 : ncl;
  uint8_t* ring = smem + L::kRingOffset;
  // ---- DYN: the tile sequence of this pair = its own cluster id, then the ids of the clusters it cancels ----------------------
  // Reads answer q (ring slot q & 1): >= 0 the cancelled cluster's id (= its tile), -1 no pending cluster is left.  `warp_wide`:
  // all 32 lanes call it (epilogue warps), lane 0 releases the slot.
  auto clc_read = [&](int q) -> int {
    const int slot = q & 1;
    mbar_wait(smem_u32(&clc_full_bar[slot]), (uint32_t)(q >> 1) & 1u);
    uint32_t valid = 0, cx = 0;
    asm volatile(
        "{\n\t.reg .pred p1;\n\t.reg .b128 resp;\n\t.reg .b32 d1, d2;\n\t"
        "ld.shared.b128 resp, [%2];\n\t"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, resp;\n\t"
        "selp.u32 %0, 1, 0, p1;\n\t"
        "mov.u32 %1, 0;\n\t"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%1, d1, d2, _}, resp;\n\t}"
        : "=r"(valid), "=r"(cx)
        : "r"(smem_u32(clc_resp + slot * 16))
        : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // our read precedes the next answer's (async proxy) write
    return valid ? (int)(cx >> 1) : -1;
  };
  auto clc_release = [&](int q) {
    const uint32_t eb = smem_u32(&clc_empty_bar[q & 1]);
    if (rank == 0) mbar_arrive(eb); else mbar_arrive_cluster(eb, 0);
  };
  auto next_item_thread = [&](int item, int& q) -> int {          // single-thread roles
    if constexpr (!DYN) return item + it_step;
    const int nx = clc_read(q);
    clc_release(q);
    ++q;
    return nx;
  };
  auto next_item_warp = [&](int item, int& q) -> int {            // epilogue warps: every lane reads, lane 0 releases
    if constexpr (!DYN) return item + it_step;
    const int nx = clc_read(q);
    __syncwarp();
    if (lane == 0) clc_release(q);
    ++q;
    return nx;
  };
  const int item0 = DYN ? cid : it_begin;
  auto item_live = [&](int item) { return DYN ? item >= 0 : item < it_end; };
  if constexpr (DYN) {
    if (warp == 10 && lane == 0 && rank == 0) {
      // scheduler: one query in flight, at most two answers ahead of the slowest consumer
      for (int q = 0;; ++q) {
        const int slot = q & 1;
        mbar_wait(smem_u32(&clc_empty_bar[slot]), ((uint32_t)(q >> 1) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&clc_full_bar[slot]);
        mbar_expect_tx(fb, 16);                                    // arm both CTAs' barriers, then ask
        asm volatile(
            "{\n\t.reg .b32 ra;\n\t"
            "mapa.shared::cluster.u32 ra, %0, 1;\n\t"
            "mbarrier.arrive.expect_tx.shared::cluster.b64 _, [ra], 16;\n\t}"
            ::"r"(fb)
            : "memory");
        asm volatile(
            "clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
            ::"r"(smem_u32(clc_resp + slot * 16)), "r"(fb)
            : "memory");
        const int nx = clc_read(q);
        clc_release(q);
        if (nx < 0) break;                                         // (a query after a failed one is undefined)
      }
    }
  }

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, pc = 0;                          // ring position; ARES: row panels this pair has finished
      int q = 0;
      PROF_DECL(w_slot = 0);
      for (int item = item0; item_live(item); item = next_item_thread(item, q)) {
        if (DYN && item >= ntiles) continue;       // (a tile beyond the device-side row count)
        const int tile = item % ntiles_mn, kb0 = (item / ntiles_mn) * nk_per, kb1 = min(nk_all, kb0 + nk_per);
        const int m0 = (tile / tiles_n) * 2 * kBM + (int)rank * kBM, n0 = (tile % tiles_n) * BN;
        const bool first = ARES && (item == it_begin || tile % tiles_n == 0);
        const bool last = ARES && (item == it_end - 1 || tile % tiles_n == tiles_n - 1);
        for (int kb = kb0; kb < kb1; kb += KPS, ++it) {
          if constexpr (ARES) {
            static_assert(!ARES || KPS == 1, "the A-resident mode keeps one k-block per stage");
            if (first) {   // (re)load slice kb of the panel once the previous panel's last tile is done with it
              PROF_WAIT(w_slot, mbar_wait(smem_u32(&a_empty_bar[kb]), (uint32_t)(pc & 1) ^ 1u));
              const uint32_t ab = smem_u32(&a_full_bar[kb]);
              if (rank == 0) mbar_expect_tx(ab, 2 * L::kABytes);
              tma_load_2d_2sm(smem_u32(smem + kb * L::kABytes), &tmA, ab, kb * kBK, m0);
            }
          }
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          PROF_WAIT(w_slot, mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1));
          // both CTAs' loads complete on the LEADER's full barrier (cta_group::2 TMA: barrier address with the peer bit
          // clear); the leader announces the bytes of the pair
          const uint32_t fb = smem_u32(&full_bar[s]);
          if (rank == 0) mbar_expect_tx(fb, 2 * L::kStageBytes);
          const uint32_t a_dst = smem_u32(ring + s * L::kStageBytes);
          if constexpr (KPS == 1) {
            if constexpr (!ARES) tma_load_2d_2sm(a_dst, &tmA, fb, kb * kBK, m0);                 // this CTA's 128 rows of A
            tma_load_2d_2sm(a_dst + (ARES ? 0 : L::kABytes), &tmB, fb, kb * kBK, n0 + (int)rank * (BN / 2));   // this CTA's half of W
          } else {
            // one 3-D box per operand: KPS consecutive k-block tiles  [A kb | A kb+1 | W kb | W kb+1]
            tma_load_3d_2sm(a_dst, &tmA, fb, 0, m0, kb);
            tma_load_3d_2sm(a_dst + KPS * L::kABytes, &tmB, fb, 0, n0 + (int)rank * (BN / 2), kb);
          }
        }
        if (last) ++pc;
      }
#ifdef BOFI_GEMM_PROF
      if (it_end > it_begin && rank == 0) PROF_ADD(prof_bucket(N, K, RESID), 3, w_slot);
#endif
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kBM, BN, A_MN, B_MN);
      int it = 0, t = 0, pc = 0, q = 0;
      PROF_DECL(w_acc = 0, w_full = 0, t_begin = PROF_T());
      for (int item = item0; item_live(item); item = next_item_thread(item, q)) {
        if (DYN && item >= ntiles) continue;
        const int kb0 = (item / ntiles_mn) * nk_per, kb1 = min(nk_all, kb0 + nk_per);
        const int tile_n = (item % ntiles_mn) % tiles_n;
        const bool first = ARES && (item == it_begin || tile_n == 0);
        const bool last = ARES && (item == it_end - 1 || tile_n == tiles_n - 1);
        const int as = t & 1;
        const uint32_t aph = (t >> 1) & 1;
        PROF_WAIT(w_acc, mbar_wait(smem_u32(&tmem_empty_bar[as]), aph ^ 1));     // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb0; kb < kb1; kb += KPS, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          if constexpr (ARES) {
            if (first) PROF_WAIT(w_full, mbar_wait(smem_u32(&a_full_bar[kb]), (uint32_t)(pc & 1)));
          }
          PROF_WAIT(w_full, mbar_wait(smem_u32(&full_bar[s]), ph));
          tcgen05_fence_after();
          const uint32_t s_addr = smem_u32(ring + s * L::kStageBytes);
#pragma unroll
          for (int kk = 0; kk < KPS; ++kk) {
            const uint32_t a_addr = ARES ? smem_u32(smem + kb * L::kABytes) : s_addr + kk * L::kABytes;
            const uint32_t b_addr = ARES ? s_addr : s_addr + KPS * L::kABytes + kk * L::kBBytes;
            const uint64_t adesc = A_MN ? make_sw128_desc_mn(a_addr) : make_sw128_desc(a_addr);
            const uint64_t bdesc = B_MN ? make_sw128_desc_mn(b_addr) : make_sw128_desc(b_addr);
            // per K step of 16: K-major +32 bytes inside the 128-byte swizzle row; MN-major +16 rows of 128 bytes (encoded >> 4)
            constexpr uint64_t a_step = A_MN ? 128 : 2, b_step = B_MN ? 128 : 2;
#pragma unroll
            for (int k = 0; k < kBK / kUK; ++k) {
              umma_bf16_2sm(tmem_d, adesc + a_step * k, bdesc + b_step * k, idesc, ((kb - kb0) | kk | k) != 0);
            }
          }
          umma_commit_2sm(smem_u32(&empty_bar[s]));      // frees the stage in BOTH CTAs when these MMAs retire
          if constexpr (ARES) {
            if (last) umma_commit_2sm(smem_u32(&a_empty_bar[kb]));   // the panel's slice may be overwritten in both CTAs
          }
        }
        umma_commit_2sm(smem_u32(&tmem_full_bar[as]));   // accumulator complete: both CTAs' epilogues
        if (last) ++pc;
        ++t;
      }
#ifdef BOFI_GEMM_PROF
      if (t > 0) {
        const int b = prof_bucket(N, K, RESID);
        PROF_ADD(b, 0, PROF_T() - t_begin); PROF_ADD(b, 1, w_acc); PROF_ADD(b, 2, w_full); PROF_ADD(b, 7, 1);
      }
#endif
    }
  } else if (warp < 10) {
    // Epilogue: TMEM -> registers (one accumulator row per thread) -> bias / ReLU / residual -> swizzled smem
    // staging tile (32 rows x 128 B) -> TMA store.  Full-line coalesced writes, M / N tails clipped by the
    // tensor map.  Eight warps: warp `quad + 4*half + 2` owns TMEM lanes [32*quad, +32) and the column chunks
    // c = half, half + 2, ...
    constexpr int CC = 128 / (int)sizeof(TOut);      // output columns per 128-byte staging row
    const int quad = warp & 3;                       // TMEM lanes [32*quad, 32*quad+32)
    const int half = (warp - 2) >> 2;
    // two staging tiles per warp: the TMA engine is shared with the operand loads, so a store can sit in its queue for a
    // while; waiting for the PREVIOUS store (not the one before it) stalled the epilogue (36 % of the kernel's stall samples)
    const uint32_t sbuf0 = smem_u32(smem + L::kOutOffset + (warp - 2) * (k2Staging * 4096));
    int nstore = 0;
    int t = 0, q = 0;
    PROF_DECL(w_tfull = 0, w_stage = 0, w_tld = 0, t_math = 0, t_store = 0, t_arrive = 0, t_begin = PROF_T());
    for (int item = item0; item_live(item); item = next_item_warp(item, q)) {
      if (DYN && item >= ntiles) continue;
      const int tile = item % ntiles_mn;
      const int m0 = (tile / tiles_n) * 2 * kBM + (int)rank * kBM, n0 = (tile % tiles_n) * BN;
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
      // vocabulary epilogues: this thread's running statistics over the columns it sees of this tile / its row's max and lse
      float vm = -INFINITY, vs = 0.f, vt = 0.f, vgv = -INFINITY, vgz = 0.f, mx_row = 0.f, lse_row = 0.f;
      int vi = 0x7fffffff, vn = 0x7fffffff, vgi = 0x7fffffff;
      if constexpr (VMODE == 2) {
        if (ve.do_lsm && row_ok) { mx_row = ve.mx[row]; lse_row = ve.lse[row]; }
      }
      if constexpr (VMODE == 1) {           // this warp's 128 bias values -> its (otherwise unused) staging tile
        __syncwarp();
        st_shared_v4(sbuf0 + (uint32_t)lane * 16u, __float_as_uint(bv.x), __float_as_uint(bv.y), __float_as_uint(bv.z), __float_as_uint(bv.w));
        __syncwarp();
      }
      PROF_WAIT(w_tfull, mbar_wait(smem_u32(&tmem_full_bar[as]), aph));
      tcgen05_fence_after();
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int c = half + 2 * i;
        const int n = n0 + c * CC;
        if (c >= NC || n >= N) break;                 // warp-uniform
        fetch_res(res[(i + 1) & 1], c + 2);
        const bool full = (n + CC <= N);              // warp-uniform
        const uint32_t sbuf = sbuf0 + (uint32_t)(k2Staging == 2 ? (nstore & 1) : 0) * 4096u;
        const uint32_t srow = sbuf + (uint32_t)lane * 128u;
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
        if constexpr (VMODE == 1) {
          static_assert(VMODE != 1 || (CC == 32 && !RELU && !RESID && !REDUCE), "statistics epilogue: fp32 logits");
          // z = accumulator + bias.  The bias of this warp's chunks sits in its staging tile (written once per tile below the
          // accumulator wait): 8 broadcast ld.shared.v4 per chunk instead of 32 shuffles.
          float z[CC];
          if (full) {
#pragma unroll
            for (int j = 0; j < CC; j += 4) {
              const float4 bq = ld_shared_v4(sbuf0 + (uint32_t)(i * CC + j) * 4u);
              z[j] = __uint_as_float(r[j]) + bq.x;
              z[j + 1] = __uint_as_float(r[j + 1]) + bq.y;
              z[j + 2] = __uint_as_float(r[j + 2]) + bq.z;
              z[j + 3] = __uint_as_float(r[j + 3]) + bq.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < CC; ++j) z[j] = (n + j < N) ? __uint_as_float(r[j]) + bias[n + j] : -INFINITY;
          }
          if (n == 0 && ve.pz0 != nullptr && row_ok) ve.pz0[row] = z[0];
          // chunk maximum (fmaxf skips NaNs; FMNMX3 takes two values per instruction)
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < CC; j += 2) cm = fmaxf(cm, fmaxf(z[j], z[j + 1]));
          if (cm > vm) {                    // a strictly larger value: the row's first maximum so far lies in this chunk
#pragma unroll
            for (int j = CC - 1; j >= 0; --j) if (z[j] == cm) vi = n + j;
            if (ve.need_lse) {
              const float sc = ve.fast_exp ? exp2f((vm - cm) * 1.4426950408889634f) : expf(vm - cm);       // exp(-inf) = 0 on the first chunk
              vs *= sc;
              vt *= sc;
            }
            vm = cm;
          }
          // NaNs are found through a sum they poison (one FADD / nothing extra per element) instead of a compare per element
          float probe;
          if (!ve.need_lse) {               // greedy pick without log-probs: no exponentials at all
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int j = 0; j < CC; j += 2) { a0 += z[j]; a1 += z[j + 1]; }
            probe = full ? a0 + a1 : 0.f;   // (the masked tail holds -inf; its columns cannot be NaN-checked this way ...)
            if (!full) {
#pragma unroll
              for (int j = 0; j < CC; ++j) if (z[j] != z[j]) probe = z[j];
            }
          } else if (ve.fast_exp) {
            const float nvml = -vm * 1.4426950408889634f;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int j = 0; j < CC; j += 2) {
              a0 += exp2f(fmaf(z[j], 1.4426950408889634f, nvml));
              a1 += exp2f(fmaf(z[j + 1], 1.4426950408889634f, nvml));
            }
            probe = a0 + a1;
            vs += probe;
          } else {
            float a0 = 0.f;
#pragma unroll
            for (int j = 0; j < CC; ++j) {
              const float ex = expf(z[j] - vm);
              a0 += ex;
              if (ve.pt != nullptr && (full || n + j < N)) vt = fmaf(ex, z[j], vt);
            }
            probe = a0;
            vs += a0;
          }
          if (probe != probe && vn == 0x7fffffff) {       // rare: locate the first NaN of the row (torch.max: a NaN wins)
#pragma unroll
            for (int j = CC - 1; j >= 0; --j) if (z[j] != z[j]) vn = n + j;
          }
          if (ve.sp.enabled) {
#pragma unroll
            for (int j = 0; j < CC; ++j) {       // fully unrolled: a partial unroll would index z[] dynamically and push it to local memory
              if (full || n + j < N) {
                const float g = gumbel_score(ve.sp, z[j], row, n + j);
                if (g > vgv) { vgv = g; vgi = n + j; vgz = z[j]; }
              }
            }
          }
          continue;
        }
        if constexpr (RESID) {
          if (full) {
            // transpose the coalesced residual registers into this thread's row through the staging tile
            if (lane == 0) { if constexpr (k2Staging == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
        if constexpr (VMODE == 2) {
          static_assert(VMODE != 2 || (CC == 32 && !RELU && !RESID && !REDUCE), "log-prob epilogue: fp32 output");
          if (ve.do_lsm) {
#pragma unroll
            for (int j = 0; j < CC; ++j) r[j] = __float_as_uint((__uint_as_float(r[j]) - mx_row) - lse_row);
          }
          if (!ve.tma_out) {
          __syncwarp();                     // the previous chunk's row stores have read the staging tile
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(srow + (uint32_t)((j ^ (lane & 7)) * 16), r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
          __syncwarp();
          // (16-byte stores on the output's alignment grid -- eight lanes per row, the four floats of a group fetched one by one
          // from the swizzled tile -- were measured SLOWER than one 128-byte row segment per instruction: 486 vs 361 us.)
          const bool col_ok = n + lane < N;
          float* orow = ve.out + (size_t)(m0 + quad * 32) * (size_t)ve.ldo + (size_t)(n + lane);
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            float v;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sbuf + (uint32_t)rr * 128u + (uint32_t)((((lane >> 2) ^ (rr & 7)) * 16) + (lane & 3) * 4)));
            if (col_ok && m0 + quad * 32 + rr < M) __stcs(orow + (size_t)rr * (size_t)ve.ldo, v);
          }
          continue;
          }
          // padded pitch: fall through to the staging tile + TMA store of the plain epilogue (columns >= N are clipped by tmC)
        }
        // this warp's staging tile must have been read out by the TMA engine (its previous store); on the residual path
        // every lane must also be done reading its residual row before the tile is overwritten
        PROF_WAIT(w_stage, if (lane == 0) { if constexpr (k2Staging == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); } __syncwarp());
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
          if constexpr (REDUCE) tma_reduce_add_2d(&tmC, sbuf, n, m0 + quad * 32);     // in-place residual: the L2 adds
          else tma_store_2d(&tmC, sbuf, n, m0 + quad * 32);
        }
        ++nstore;
#ifdef BOFI_GEMM_PROF
        t_store += PROF_T() - t_s0;
#endif
      }
      if constexpr (VMODE == 1) {
        if (row_ok) {                       // SoA records: consecutive lanes = consecutive rows, coalesced
          const size_t at = (size_t)((tile % tiles_n) * 2 + half) * (size_t)M + (size_t)row;
          ve.pm[at] = vm;
          ve.ps[at] = vs;
          ve.pi[at] = vi;
          ve.pn[at] = vn;
          if (ve.pt) ve.pt[at] = vt;
          if (ve.sp.enabled) { ve.pgv[at] = vgv; ve.pgi[at] = vgi; ve.pgz[at] = vgz; }
        }
      }
      PROF_DECL(t_a0 = PROF_T());
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[as]), 0);     // the leader's MMA thread waits for both CTAs
#ifdef BOFI_GEMM_PROF
      t_arrive += PROF_T() - t_a0;
#endif
      ++t;
    }
#ifdef BOFI_GEMM_PROF
    if (warp == 2 && lane == 0 && t > 0 && rank == 0) {
      const int b = prof_bucket(N, K, RESID);
      PROF_ADD(b, 4, PROF_T() - t_begin); PROF_ADD(b, 5, w_tfull); PROF_ADD(b, 6, w_stage);
      PROF_ADD(b, 8, w_tld); PROF_ADD(b, 9, t_math); PROF_ADD(b, 10, t_store); PROF_ADD(b, 11, t_arrive);
    }
#endif
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tcgen05_fence_before();
  cluster_sync_all();                              // the peer may still be reading / the leader's MMAs still reading peer smem
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
  }
}


// BOFI_GEMM_DYN=1: dynamic tile scheduling (cluster launch control), see the kernel's DYN note
inline bool gemm_dyn() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BOFI_GEMM_DYN");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

template <typename TOut, bool RELU, bool RESID, bool REDUCE, bool ARES, int KPS, int VMODE, bool DYN>
inline cudaError_t launch2_impl(cudaStream_t s, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const float* bias,
                                const float* residual, int ldr, int M, int N, int K, int relu, const int* live_rows, const int* rows_dev,
                                const VocabEpi& ve) {
  static PerDevice<bool> configured_dev;
  bool& configured = configured_dev.get();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<TOut, RELU, RESID, REDUCE, ARES, KPS, VMODE, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem2T<ARES, KPS>::kTotal);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int pairs = ((M + 2 * kBM - 1) / (2 * kBM)) * ((N + k2BN - 1) / k2BN);
  int clusters = num_sms() / 2;
  if (pairs < clusters || DYN) clusters = pairs;          // DYN: one cluster per tile, the running ones cancel the pending ones
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(DYN ? kThreadsDyn : kThreads);
  cfg.dynamicSmemBytes = Smem2T<ARES, KPS>::kTotal;
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
  return cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<TOut, RELU, RESID, REDUCE, ARES, KPS, VMODE, DYN>, tmA, tmB, tmC, bias, residual, ldr, M, N, K, relu, live_rows, rows_dev, ve);
}

template <typename TOut, bool RELU, bool RESID, bool REDUCE = false, bool ARES = false, int KPS = 1, int VMODE = 0>
inline cudaError_t launch2(cudaStream_t s, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const float* bias,
                           const float* residual, int ldr, int M, int N, int K, int relu, const int* live_rows, const int* rows_dev,
                           const VocabEpi& ve = VocabEpi()) {
  if constexpr (!ARES) {
    if (gemm_dyn()) return launch2_impl<TOut, RELU, RESID, REDUCE, ARES, KPS, VMODE, true>(s, tmA, tmB, tmC, bias, residual, ldr, M, N, K, relu, live_rows, rows_dev, ve);
  }
  return launch2_impl<TOut, RELU, RESID, REDUCE, ARES, KPS, VMODE, false>(s, tmA, tmB, tmC, bias, residual, ldr, M, N, K, relu, live_rows, rows_dev, ve);
}

// Same contract as gemm_tc (K-major A [M,K], W [N,K]).  `want_ares`: A-resident tiles where the shape allows (opt-in,
// BOFI_ARES=1 -- measured 2 % slower end to end than the streaming tile, see the header).
template <typename TOut>
inline cudaError_t gemm_tc2(cudaStream_t s, const bf16* A, int lda, const bf16* W, int ldw, const float* bias, const float* residual, int ldr,
                            TOut* C, int ldc, int M, int N, int K, int relu, const int* live_rows, const int* rows_dev, bool want_ares = false, int kps = 2) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (lda % 8 != 0 || ldw % 8 != 0 || (ldc * sizeof(TOut)) % 16 != 0 || (residual && ldr % 4 != 0) || !bias) return cudaErrorInvalidValue;
  // A-resident tiles: the whole contraction of a row panel fits the 128 KB panel, and the panel is reused by >= 4 column tiles
  const bool ares = want_ares && !residual && K <= k2PanelKB * kBK && N >= 4 * k2BN;
  const bool kps2 = kps >= 2 && !ares && K % (2 * kBK) == 0;      // two k-blocks per stage, one 3-D box per operand
  const CUtensorMap* tmA = kps2 ? cached_tmap_kblocks(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, 2)
                                : cached_tmap(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM);
  const CUtensorMap* tmB = kps2 ? cached_tmap_kblocks(W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, k2BN / 2, 2)
                                : cached_tmap(W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, k2BN / 2);
  const CUtensorMap* tmC = cached_tmap(C, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 32, (int)sizeof(TOut));
  if (!tmA || !tmB || !tmC) return cudaErrorInvalidValue;
#define BOFI_TC2(RELU_, RESID_) \
  (kps2 ? launch2<TOut, RELU_, RESID_, false, false, 2>(s, *tmA, *tmB, *tmC, bias, residual, ldr, M, N, K, relu, live_rows, rows_dev) \
        : launch2<TOut, RELU_, RESID_>(s, *tmA, *tmB, *tmC, bias, residual, ldr, M, N, K, relu, live_rows, rows_dev))
#define BOFI_TC2A(RELU_) launch2<TOut, RELU_, false, false, true>(s, *tmA, *tmB, *tmC, bias, nullptr, 0, M, N, K, relu, live_rows, rows_dev)
  if constexpr (sizeof(TOut) == 4) {
    if (residual && relu) return cudaErrorInvalidValue;
    if (residual && inplace_reduce() && (const void*)residual == (const void*)C && ldr == ldc)
      return kps2 ? launch2<TOut, false, false, true, false, 2>(s, *tmA, *tmB, *tmC, bias, nullptr, 0, M, N, K, 0, live_rows, rows_dev)
                  : launch2<TOut, false, false, true>(s, *tmA, *tmB, *tmC, bias, nullptr, 0, M, N, K, 0, live_rows, rows_dev);
    if (residual) return BOFI_TC2(false, true);
    if (ares) return relu ? BOFI_TC2A(true) : BOFI_TC2A(false);
    if (relu) return BOFI_TC2(true, false);
    return BOFI_TC2(false, false);
  } else {
    if (residual) return cudaErrorInvalidValue;
    if (ares) return relu ? BOFI_TC2A(true) : BOFI_TC2A(false);
    if (relu) return BOFI_TC2(true, false);
    return BOFI_TC2(false, false);
  }
#undef BOFI_TC2
#undef BOFI_TC2A
}

// Vocabulary projection without a logits tensor (see VocabEpi): vmode 1 = statistics records, vmode 2 = log-probs / raw logits
// into the caller's rows.  A [M, K] bf16 (K % 128 == 0), W [N, K] bf16, bias fp32 [N].
inline cudaError_t gemm_tc2_vocab(cudaStream_t s, const bf16* A, int lda, const bf16* W, int ldw, const float* bias, int M, int N, int K,
                                  int vmode, const VocabEpi& ve) {
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (lda % 8 != 0 || ldw % 8 != 0 || !bias || K % (2 * kBK) != 0) return cudaErrorInvalidValue;
  const CUtensorMap* tmA = cached_tmap_kblocks(A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, kBM, 2);
  const CUtensorMap* tmB = cached_tmap_kblocks(W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, k2BN / 2, 2);
  if (!tmA || !tmB) return cudaErrorInvalidValue;
  // no TMA store in the statistics pass, nor in the log-prob pass on a dense [M, V] tensor (rows only 4-byte aligned): the C
  // tensor map is a placeholder there (never dereferenced)
  if (vmode == 1) return launch2<float, false, false, false, false, 2, 1>(s, *tmA, *tmB, *tmA, bias, nullptr, 0, M, N, K, 0, nullptr, nullptr, ve);
  if (vmode == 2) {
    if (!ve.tma_out) return launch2<float, false, false, false, false, 2, 2>(s, *tmA, *tmB, *tmA, bias, nullptr, 0, M, N, K, 0, nullptr, nullptr, ve);
    const CUtensorMap a = *tmA, b = *tmB;           // (a third lookup follows: copies, see cached_tmap)
    const CUtensorMap* tmC = cached_tmap(ve.out, (uint64_t)M, (uint64_t)N, (uint64_t)ve.ldo, 32, 4);
    if (!tmC) return cudaErrorInvalidValue;
    return launch2<float, false, false, false, false, 2, 2>(s, a, b, *tmC, bias, nullptr, 0, M, N, K, 0, nullptr, nullptr, ve);
  }
  return cudaErrorInvalidValue;
}

}  // namespace tc
}  // namespace bofi
