// Microbenchmark: how fast can ONE SM's TMA engine fill shared memory from an L2-resident operand, and what is the
// loaded latency?  Answers the question tools/gemm_stalls.py raises (the GEMM's MMA thread waits for operands 40-50 % of
// the time although neither HBM nor the L2 is saturated).
//
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools_tmp/tma_feed_bench tools/tma_feed_bench.cu
//   ./tma_feed_bench
//
// Every CTA (one per SM, `grid` of them) runs a producer thread that keeps a ring of `stages` x `stage_bytes` full and
// a consumer thread that frees a stage as soon as it has landed (no MMA).  Modes:
//   0  tensor 2D, 128B swizzle, boxes of 64 x 128 bf16 (16 KB) -- what gemm_tc*.cuh issue
//   1  tensor 2D, 128B swizzle, boxes of 64 x 256 bf16 (32 KB)
//   2  1D bulk copies (cp.async.bulk.shared::cluster.global) of 16 KB from a pre-tiled buffer
//   4  tensor 3D boxes (64 k, 128 rows, 2 k-blocks) = 32 KB per instruction: two consecutive 16 KB k-block tiles
//   5  tensor 3D boxes (64 k, 128 rows, 4 k-blocks) = 64 KB per instruction
//   6  tensor 3D boxes (64 k, 256 rows, 2 k-blocks) = 64 KB per instruction
//   9  as 0, but the boxes of a stage alternate between TWO tensor maps (two separate buffers), as A and W do in a GEMM
//   7  one 16 KB tensor box (TMA) + 16 KB by cp.async (LDGSTS, 16 B per thread, 128 threads of four extra warps) per stage
//   8  32 KB per stage by cp.async only (128 threads)
//   +100 every CTA reads one of only 8 tile sequences (hot L2 lines, like the W tile of a GEMM wave)
//   +10  (e.g. 11) the same, while eight other warps each keep issuing 4 KB TMA STORES (32 x 128 B, the GEMM epilogue's)
//   +20  (e.g. 21) the same, while two other warps each keep issuing 16 KB TMA stores (128 x 128 B)
// Output: bytes / clk / SM over the whole run (clock64 of the slowest CTA) per (mode, stages, stage bytes, grid).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../boficap_b200/csrc/gemm_tc.cuh"

using namespace bofi::tc;

// SPIN=1: poll with the non-suspending mbarrier.test_wait instead of try_wait (separates the memory path from the
// wake-up behaviour of a suspended try_wait)
#ifdef SPIN
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
#else
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
#endif
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// [rows, K] bf16 row-major seen as (64 k, rows, K/64 k-blocks): box (64, box_rows, box_kb) lands as box_kb consecutive
// 128B-swizzled k-block tiles of box_rows x 128 B
static bool make_tmap_3d(CUtensorMap* tm, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_kb) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[3] = {64, rows, cols / 64};
  cuuint64_t gstride[2] = {ld * 2, 128};
  cuuint32_t box[3] = {64, box_rows, box_kb};
  cuuint32_t estr[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

__global__ void __launch_bounds__(320, 1)
feed_kernel(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm256, const __grid_constant__ CUtensorMap tm3a,
            const __grid_constant__ CUtensorMap tm3b, const __grid_constant__ CUtensorMap tm3c, const uint8_t* flat,
            const __grid_constant__ CUtensorMap tmS4, const __grid_constant__ CUtensorMap tmS16, const __grid_constant__ CUtensorMap tm128b,
            int mode, int stages, int stage_bytes, int iters, int rows_total, long long* cycles, unsigned long long* stores_done) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[16], empty[16];
  __shared__ volatile int done_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool same_tiles = mode >= 100;      // +100: every CTA walks the SAME tile sequence (what the CTAs of a GEMM wave do with W)
  mode %= 100;
  const int st_mode = mode / 10;
  mode %= 10;
  if (threadIdx.x == 0) done_flag = 0;
  if (threadIdx.x == 0) {
    const int extra = (mode % 100 % 10 == 7 || mode % 100 % 10 == 8) ? 128 : 0;      // cp.async.mbarrier.arrive.noinc of the LDGSTS threads
    for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full[s]), (mode % 100 % 10 == 8 ? 0 : 1) + extra); mbar_init(smem_u32(&empty[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      const uint32_t ph = (it / stages) & 1;
      wait_bar(smem_u32(&empty[s]), ph ^ 1);
      const uint32_t fb = smem_u32(&full[s]);
      if (mode == 8) continue;                       // nothing for the TMA thread to do
      mbar_expect_tx(fb, mode == 7 ? 16384 : stage_bytes);
      const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes);
      // walk an L2-resident window: k-block (it % 8), row block depends on CTA and iteration
      const int kb = it & 7;
      const int r0 = (int)(((long long)(same_tiles ? (blockIdx.x & 7) : blockIdx.x) * 131 + (long long)(it >> 3) * 7) % (rows_total / 256)) * 256;
      if (mode == 7) { tma_load_2d(dst, &tm128, fb, kb * 64, r0); continue; }
      if (mode == 4) {
        for (int i = 0; i < stage_bytes / 32768; ++i) tma_load_3d(dst + i * 32768, &tm3a, fb, 0, (r0 + 128 * i) % rows_total, (kb & 3) * 2);
      } else if (mode == 5) {
        for (int i = 0; i < stage_bytes / 65536; ++i) tma_load_3d(dst + i * 65536, &tm3b, fb, 0, (r0 + 128 * i) % rows_total, (kb & 1) * 4);
      } else if (mode == 6) {
        for (int i = 0; i < stage_bytes / 65536; ++i) tma_load_3d(dst + i * 65536, &tm3c, fb, 0, r0, (kb & 3) * 2);
      } else if (mode == 9) {
        for (int i = 0; i < stage_bytes / 16384; ++i) tma_load_2d(dst + i * 16384, (i & 1) ? &tm128b : &tm128, fb, kb * 64, (r0 + 128 * i) % rows_total);
      } else if (mode == 0) {
        for (int i = 0; i < stage_bytes / 16384; ++i) tma_load_2d(dst + i * 16384, &tm128, fb, kb * 64, (r0 + 128 * i) % rows_total);
      } else if (mode == 1) {
        for (int i = 0; i < stage_bytes / 32768; ++i) tma_load_2d(dst + i * 32768, &tm256, fb, kb * 64, r0);
      } else {
        for (int i = 0; i < stage_bytes / 16384; ++i)
          bulk_load_1d(dst + i * 16384, flat + ((size_t)(r0 / 128 + i) * 8 + kb) % ((size_t)rows_total / 128 * 8) * 16384, 16384, fb);
      }
    }
  } else if (warp == 1 && lane == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      const uint32_t ph = (it / stages) & 1;
      wait_bar(smem_u32(&full[s]), ph);
      mbar_arrive(smem_u32(&empty[s]));
    }
    cycles[blockIdx.x] = clock64() - t0;
    done_flag = 1;
  } else if ((mode == 7 || mode == 8) && warp >= 6) {
    // LDGSTS producers: 128 threads copy 16 KB (mode 7: the second half of the stage) or 32 KB (mode 8) per stage in
    // 16-byte pieces, written in the 128B-swizzle pattern, and signal the stage's full barrier when their copies land
    const int tid = threadIdx.x - 192;
    const int bytes = mode == 7 ? 16384 : 32768;
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      const uint32_t ph = (it / stages) & 1;
      wait_bar(smem_u32(&empty[s]), ph ^ 1);
      const int kb = it & 7;
      const int r0 = (int)(((long long)blockIdx.x * 131 + (long long)(it >> 3) * 7) % (rows_total / 256)) * 256 + 128;
      const uint32_t dst = smem_u32(smem + (size_t)s * stage_bytes + (mode == 7 ? 16384 : 0));
      for (int c = tid; c < bytes / 16; c += 128) {
        const int row = c >> 3, ch = c & 7;            // rows of 128 bytes, 8 chunks each
        const uint8_t* src = flat + ((size_t)((r0 + row) % rows_total) * 512 + kb * 64) * 2 + ch * 16;
        const uint32_t d = dst + row * 128 + ((ch ^ (row & 7)) * 16);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
    }
  } else if (st_mode && warp >= 2 && lane == 0 && (st_mode == 1 || warp < 4)) {
    // concurrent TMA stores from a scratch staging area behind the ring (contents irrelevant)
    const uint32_t src = smem_u32(smem + (size_t)stages * stage_bytes + (st_mode == 1 ? (warp - 2) * 4096 : (warp - 2) * 16384));
    unsigned long long n = 0;
    while (!done_flag) {
      if (st_mode == 1) tma_store_2d(&tmS4, src, 0, (int)((blockIdx.x * 8 + (warp - 2)) * 32));
      else tma_store_2d(&tmS16, src, 0, (int)((blockIdx.x * 2 + (warp - 2)) * 128));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      ++n;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    atomicAdd(stores_done, n);
  }
}

int main() {
  const int rows = 32768, K = 512;                 // 32 MB of bf16: L2 resident after the warm-up pass
  __nv_bfloat16* buf;
  cudaMalloc(&buf, (size_t)rows * K * 2);
  cudaMemset(buf, 0, (size_t)rows * K * 2);
  long long* cyc;
  cudaMalloc(&cyc, 148 * sizeof(long long));
  CUtensorMap tm128, tm256, tm3a, tm3b, tm3c;
  if (!make_tmap(&tm128, buf, rows, K, K, 128, 2) || !make_tmap(&tm256, buf, rows, K, K, 256, 2) || !make_tmap_3d(&tm3a, buf, rows, K, K, 128, 2) ||
      !make_tmap_3d(&tm3b, buf, rows, K, K, 128, 4) || !make_tmap_3d(&tm3c, buf, rows, K, K, 256, 2)) { printf("tmap failed\n"); return 1; }
  // store target: [148 * 256 rows, 64 bf16] scratch
  __nv_bfloat16* sbuf;
  cudaMalloc(&sbuf, (size_t)148 * 256 * 64 * 2);
  CUtensorMap tmS4, tmS16;
  if (!make_tmap(&tmS4, sbuf, 148 * 256, 64, 64, 32, 2) || !make_tmap(&tmS16, sbuf, 148 * 256, 64, 64, 128, 2)) { printf("tmap failed\n"); return 1; }
  __nv_bfloat16* buf2;
  cudaMalloc(&buf2, (size_t)rows * K * 2);
  cudaMemset(buf2, 0, (size_t)rows * K * 2);
  CUtensorMap tm128b;
  if (!make_tmap(&tm128b, buf2, rows, K, K, 128, 2)) { printf("tmap failed\n"); return 1; }
  unsigned long long* stores_done;
  cudaMalloc(&stores_done, 8);
  cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("mode stages stage_KB grid | B/clk/SM  (in flight KB)  implied latency clk\n");
  const int iters = 4000;
  struct Cfg { int mode, stages, stage_bytes, grid; };
  std::vector<Cfg> cfgs;
  for (int grid : {1}) {
    if (grid == 1) break;                       // 16 CTAs: the per-SM limit without chip-wide contention                       // 16 CTAs: the per-SM limit without chip-wide contention
    for (int mode : {0, 1, 2, 4})
      for (int st : {1, 2, 4, 6}) cfgs.push_back({mode, st, 32768, grid});
    for (int mode : {5, 6})
      for (int st : {1, 2, 3}) cfgs.push_back({mode, st, 65536, grid});
    for (int st : {4, 8, 12}) cfgs.push_back({0, st, 16384, grid});
  }
  for (int mode : {0}) cfgs.push_back({mode, mode % 10 == 5 ? 2 : 4, mode % 10 == 5 ? 65536 : 32768, 148});
  // hot lines: 148 CTAs walking only 8 distinct tile sequences (+100) against 148 distinct ones
  for (int mode : {0, 100, 4, 104}) { cfgs.push_back({mode, 3, 65536, 148}); cfgs.push_back({mode, 6, 32768, 148}); }
  for (int n : {2, 4}) { cfgs.push_back({0, 2, n * 16384, 148}); cfgs.push_back({9, 2, n * 16384, 148}); cfgs.push_back({9, 4, n * 16384, 148}); }
  // how a stage's cost grows with the number of boxes on its barrier
  for (int n : {1, 2, 3, 4, 6}) cfgs.push_back({0, 2, n * 16384, 148});
  for (int n : {1, 2, 3}) cfgs.push_back({1, 2, n * 32768, 148});
  for (int n : {1, 2, 3}) cfgs.push_back({4, 2, n * 32768, 148});
  for (int mode : {0, 7, 8})
    for (int st : {2, 4, 6}) cfgs.push_back({mode, st, 32768, 148});
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(stores_done, 0, 8);
      feed_kernel<<<c.grid, 320, c.stages * c.stage_bytes + 32768 + 1024>>>(tm128, tm256, tm3a, tm3b, tm3c, (const uint8_t*)buf, tmS4, tmS16, tm128b, c.mode, c.stages, c.stage_bytes, iters, rows, cyc, stores_done);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    }
    std::vector<long long> h(c.grid);
    cudaMemcpy(h.data(), cyc, c.grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long v : h) mx = v > mx ? v : mx;
    const double bpc = (double)iters * c.stage_bytes / (double)mx;
    const double inflight = (double)c.stages * c.stage_bytes;
    unsigned long long nst = 0;
    cudaMemcpy(&nst, stores_done, 8, cudaMemcpyDeviceToHost);
    printf("%4d %6d %8d %4d | %8.1f  (%6.0f)  %8.0f   stores/SM %.0f (one per %.0f clk)\n", c.mode, c.stages, c.stage_bytes / 1024, c.grid, bpc, inflight / 1024,
           inflight / bpc, (double)nst / c.grid, nst ? (double)mx * c.grid / (double)nst : 0.0);
  }
  return 0;
}
