// The whole bounding loop of core_NAIC (TransformerModel.py:1833-1870 driving LengthPredictor_UIC :357-383, N_len == 1) as ONE
// kernel launch (bf16 engine).
//
// Round 1 enqueued 19 steps x 10 dependent launches of small kernels on M = rows (one graph replay): at B = 1024 that chain is
// ~2 ms of mostly idle GPU per batch, at B = 1 it is the whole latency (84 us per step).  Rows are independent in the bounding
// loop, and a step only ever touches the [LEN] row of an image, so a group of 64 rows can run all its steps without talking to
// any other group: here a thread-block CLUSTER of 8 CTAs owns 64 rows and walks the steps with cluster barriers only --
// no grid-wide synchronisation, no cooperative launch, clusters are scheduled like ordinary blocks.
//
// Work split inside a cluster (16 warps per CTA):
//   row phases   (self-attention over the tabulated slot K/V, cross-attention over the image regions, heads + box rule):
//                CTA c owns rows [8c, 8c+8) of the cluster's 64
//   GEMM phases  (O-projections, Q-projection, FFN): CTA c owns 1/8 of the OUTPUT COLUMNS for all 64 rows; the 64 x K
//                activation block and the CTA's 64-column weight slice are staged in shared memory with cp.async and
//                multiplied with mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  Every weight byte is read once per cluster
//                and step (8.4 MB), i.e. 1.05 MB per CTA and step.
// LayerNorm is fused into the staging of the A operand (each CTA normalises the 64 rows it is about to multiply), residual
// adds are in the GEMM epilogues, activations move between phases through L2 (x fp32, ao / q / ffh bf16) under
// barrier.cluster release / acquire.  tcgen05 would need M = 128 rows per CTA; the point here is latency, not tensor peak.
//
// MEASURED (B200, round 2, profiles/r02d_*): parity-green (tests/test_gpu_decode.py::test_cluster_bounding_loop_matches_launch_chain)
// and one launch instead of ~190 -- but SLOWER than the chain it replaces, so it is opt-in (BOFI_BOUND_CLUSTER=1):
//   B = 1024: 3.84 ms for the loop (200 us per step) -> 128.6 k captions/s against 184.5 k with the launch chain (7.96 vs 5.55 ms/step)
//   B = 1:    2.08 ms per decode against 1.92 ms; B = 32: 2.80 vs 2.24 ms; B = 512: 4.65 vs 3.98 ms
// A step is a chain of ~25 "issue loads, wait for the last byte, compute" stages (11 GEMM units, two attention phases with
// four sequential (row, head) tasks per warp, two LayerNorm stagings, a 16-round-trip head) with nothing prefetched across a
// stage: ~95 us per step even for one cluster, i.e. the latency per step of the launch chain, and under the L2 contention of
// 128 CTAs twice that.  The launch chain spreads each of those stages over all 148 SMs.  What would make this form win is
// prefetching the (activation-independent) weight slices one unit ahead and batching the per-task loads of the row phases;
// ncu: issue active 26 %, tensor pipe 5 %, DRAM 0.3 %.
#pragma once
#include "attention_mma.cuh"
#include "gemm_tc2.cuh"
#include "kernels.cuh"

namespace bofi {

constexpr int kBlRows = 64;                       // rows per cluster
constexpr int kBlCtas = 8;                        // CTAs per cluster (portable maximum)
constexpr int kBlOwn = kBlRows / kBlCtas;         // rows a CTA owns in the row phases
constexpr int kBlThreads = 512;
constexpr int kBlPitch = kD + 8;                  // bf16 per staged row: 1040 bytes, conflict-free ldmatrix
constexpr int kBlTileBytes = kBlRows * kBlPitch * 2;
constexpr int kBlSmem = 2 * kBlTileBytes + kD * 4 + (kBlThreads / 32) * kMaxKeys * 4 + 64;

struct BoundLoopParams {
  const bf16 *w_so, *w_q, *w_co, *w_1, *w_2;      // [512,512] x3, [d_ff,512], [512,d_ff]  (K-major, nn.Linear layout)
  const float *b_so, *b_q, *b_co, *b_1, *b_2;
  const float *ln1_a, *ln1_b, *ln2_a, *ln2_b;     // sublayer norms 1 (cross-attention) and 2 (FFN)
  const float *lnh_a, *lnh_b, *w1t, *b1h, *w_len, *b_len, *w_syn, *b_syn;   // length_predictor.norm + the two heads (fp32)
  const bf16* tab_qkv;                            // [10 * Lb, 1536] LN + QKV of every (syn id, position) input row
  const float* x0;                                // the constant [LEN] input row (bound_in[q_row])
  int q_row;
  const bf16* kv;                                 // memory K | V of the bounding layer [B*R or compact, 1024]
  const int* mem_len;                             // valid regions per image or nullptr
  const int* mem_off;                             // compact (varlen) row offset per image or nullptr
  int R, sn;
  float* x;                                       // [rows, 512] fp32 residual stream
  bf16 *ao, *q, *ffh;                             // [rows, 512], [rows, 512], [rows, d_ff]
  int* cl_live;                                   // [clusters * 8] live rows per CTA after the last head phase
  DecodeState st;
  int rows, Lb, L, nsteps, d_ff, hh, n_len, n_syn, syn_lo, syn_hi;
};

__device__ __forceinline__ void bl_cluster_sync() { tc::cluster_sync_all(); }

// 64 rows x 512 bf16 columns of a row-major global matrix (pitch ld, column offset already applied) -> staged tile
__device__ __forceinline__ void bl_stage_rows(bf16* dst, const bf16* __restrict__ src, size_t ld, int valid_rows, int tid) {
#pragma unroll
  for (int i = 0; i < (kBlRows * (kD / 8)) / kBlThreads; ++i) {
    const int idx = tid + i * kBlThreads;
    const int row = idx >> 6, ch = idx & 63;
    const bool ok = row < valid_rows;
    cp_async_16(dst + row * kBlPitch + ch * 8, src + (size_t)(ok ? row : 0) * ld + ch * 8, ok);
  }
}

// A operand = LayerNorm(x rows) in bf16 (the arithmetic of layernorm_kernel's bf16 branch); warp w normalises rows w, w+16, ...
__device__ __forceinline__ void bl_stage_layernorm(bf16* dst, const float* __restrict__ x, int row0, int valid_rows, const float* __restrict__ ga,
                                                   const float* __restrict__ gb, int warp, int lane) {
  for (int r = warp; r < kBlRows; r += kBlThreads / 32) {
    bf16* d = dst + r * kBlPitch;
    if (r >= valid_rows) {
#pragma unroll
      for (int i = 0; i < 4; ++i) store4(d + (i * 32 + lane) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    const float* xr = x + (size_t)(row0 + r) * kD;
    float4 v[4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = __ldcg(reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4));
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / kD);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float denom = sqrtf(warp_sum(ss) * (1.0f / (kD - 1))) + 1e-6f;
    const float inv = 1.0f / denom;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 a = load4(ga + c), b = load4(gb + c);
      float4 o;
      o.x = ln_div(a.x * v[i].x, denom, inv) + b.x;
      o.y = ln_div(a.y * v[i].y, denom, inv) + b.y;
      o.z = ln_div(a.z * v[i].z, denom, inv) + b.z;
      o.w = ln_div(a.w * v[i].w, denom, inv) + b.w;
      store4(d + c, o);
    }
  }
}

// acc (16 rows x 16 columns per warp) += As[64 x 512] . Ws[64 x 512]^T ; warp (rt, cg): rows 16 rt.., columns 16 cg..
__device__ __forceinline__ void bl_mma_512(float (&acc)[2][4], uint32_t as_base, uint32_t ws_base, int rt, int cg, int lane) {
  const int arow = rt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, acol = (lane >> 4) * 8;
  const int brow = cg * 16 + (lane & 7), bcol = (lane >> 3) * 8;
#pragma unroll 4
  for (int k0 = 0; k0 < kD; k0 += 32) {
    uint32_t a0[4], a1[4];
    ldmatrix_x4(a0, as_base + (uint32_t)(arow * kBlPitch + k0 + acol) * 2u);
    ldmatrix_x4(a1, as_base + (uint32_t)(arow * kBlPitch + k0 + 16 + acol) * 2u);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      uint32_t kb[4];
      ldmatrix_x4(kb, ws_base + (uint32_t)((brow + nt * 8) * kBlPitch + k0 + bcol) * 2u);
      mma_bf16_16816(acc[nt], a0, kb[0], kb[1]);
      mma_bf16_16816(acc[nt], a1, kb[2], kb[3]);
    }
  }
}

// Epilogues.  Thread (g = lane >> 2, t4 = lane & 3) of warp (rt, cg) holds rows 16 rt + g (+8), columns 16 cg + 8 nt + 2 t4 (+1).
enum { BL_EPI_X0 = 0, BL_EPI_XADD = 1, BL_EPI_BF16 = 2, BL_EPI_RELU_BF16 = 3 };
template <int EPI>
__device__ __forceinline__ void bl_epilogue(const float (&acc)[2][4], const float* __restrict__ bias, int col0, int row0, int valid_rows, int rt, int cg,
                                            int lane, float* __restrict__ x, const float* __restrict__ x0, bf16* __restrict__ out, size_t ldo) {
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int c = col0 + cg * 16 + nt * 8 + 2 * t4;
    const float b0 = bias[c], b1 = bias[c + 1];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = rt * 16 + g + h * 8;
      if (r >= valid_rows) continue;
      const float v0 = acc[nt][2 * h] + b0, v1 = acc[nt][2 * h + 1] + b1;
      if constexpr (EPI == BL_EPI_X0) {
        *reinterpret_cast<float2*>(x + (size_t)(row0 + r) * kD + c) = make_float2(x0[c] + v0, x0[c + 1] + v1);
      } else if constexpr (EPI == BL_EPI_XADD) {
        float2* p = reinterpret_cast<float2*>(x + (size_t)(row0 + r) * kD + c);
        const float2 o = *p;                      // written by this very thread in an earlier phase
        *p = make_float2(o.x + v0, o.y + v1);
      } else if constexpr (EPI == BL_EPI_BF16) {
        *reinterpret_cast<uint32_t*>(out + (size_t)(row0 + r) * ldo + c) = pack2_bf16(v0, v1);
      } else {
        *reinterpret_cast<uint32_t*>(out + (size_t)(row0 + r) * ldo + c) = pack2_bf16(fmaxf(v0, 0.f), fmaxf(v1, 0.f));
      }
    }
  }
}

// softmax(q . K^T / 8 over the first nvis keys) . V for ONE (row, head), one warp.  K / V rows: 64 bf16 of this head, row j at
// kbase + idx(j) * ld.  16-byte loads, four lanes per key for the scores, eight lanes per key for P.V (as attention_row_bf16_kernel).
template <typename IdxFn>
__device__ __forceinline__ void bl_attend(const float (&qf)[16], const bf16* __restrict__ kbase, const bf16* __restrict__ vbase, size_t ld, int nvis,
                                          IdxFn idx, float* __restrict__ ps, bf16* __restrict__ orow, int lane) {
  const int sub = lane & 3, kslot = lane >> 2;
  float sc[kMaxKeys / 8];
  float mx = -INFINITY;
  const int npass = (nvis + 7) >> 3;
#pragma unroll
  for (int p = 0; p < kMaxKeys / 8; ++p) sc[p] = -INFINITY;
#pragma unroll
  for (int p0 = 0; p0 < kMaxKeys / 8; p0 += 4) {
    if (p0 >= npass) break;
    uint4 k0[4], k1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = (p0 + u) * 8 + kslot;
      k0[u] = k1[u] = make_uint4(0u, 0u, 0u, 0u);
      if (j < nvis) {
        const bf16* kr = kbase + (size_t)idx(j) * ld + sub * 16;
        k0[u] = __ldcg(reinterpret_cast<const uint4*>(kr));
        k1[u] = __ldcg(reinterpret_cast<const uint4*>(kr + 8));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = (p0 + u) * 8 + kslot;
      const uint32_t w[8] = {k0[u].x, k0[u].y, k0[u].z, k0[u].w, k1[u].x, k1[u].y, k1[u].z, k1[u].w};
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        d = fmaf(qf[2 * i], f.x, d);
        d = fmaf(qf[2 * i + 1], f.y, d);
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      if (j < nvis) sc[p0 + u] = d * 0.125f;
      mx = fmaxf(mx, sc[p0 + u]);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int p = 0; p < kMaxKeys / 8; ++p) {
    if (p < npass) {
      const int j = p * 8 + kslot;
      const float e = (j < nvis) ? __expf(sc[p] - mx) : 0.f;
      sc[p] = e;
      if (sub == 0) sum += e;
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  __syncwarp();
#pragma unroll
  for (int p = 0; p < kMaxKeys / 8; ++p) {
    if (p < npass && sub == 0) ps[p * 8 + kslot] = sc[p] * inv;
  }
  __syncwarp();
  const int g = lane & 7, ksub = lane >> 3;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int j0 = 0; j0 < nvis; j0 += 16) {
    uint4 vv[4];
    float pj[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 4 + ksub;
      pj[u] = 0.f;
      vv[u] = make_uint4(0u, 0u, 0u, 0u);
      if (j < nvis) {
        vv[u] = __ldcg(reinterpret_cast<const uint4*>(vbase + (size_t)idx(j) * ld + g * 8));
        pj[u] = ps[j];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t w[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        acc[2 * i] = fmaf(pj[u], f.x, acc[2 * i]);
        acc[2 * i + 1] = fmaf(pj[u], f.y, acc[2 * i + 1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
  }
  if (ksub == 0) {
    if (nvis <= 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = __int_as_float(0x7fc00000);
    }
    uint4 o;
    o.x = pack2_bf16(acc[0], acc[1]);
    o.y = pack2_bf16(acc[2], acc[3]);
    o.z = pack2_bf16(acc[4], acc[5]);
    o.w = pack2_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(orow + g * 8) = o;
  }
  __syncwarp();
}

__global__ void __cluster_dims__(kBlCtas, 1, 1) __launch_bounds__(kBlThreads, 1) bound_loop_kernel(const BoundLoopParams p) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t bl_smem[];
  bf16* As = reinterpret_cast<bf16*>(bl_smem);
  bf16* Ws = reinterpret_cast<bf16*>(bl_smem + kBlTileBytes);
  float* q0s = reinterpret_cast<float*>(bl_smem + 2 * kBlTileBytes);                  // [512] Q of the [LEN] table row
  float* ps_all = q0s + kD;                                                             // [16][128] attention probabilities
  // head phase scratch aliases the staging tiles (no GEMM is in flight then)
  float* hs = reinterpret_cast<float*>(bl_smem);                                        // [8][512]
  float* hid = hs + kBlOwn * kD;                                                        // [8][2*hh]
  float* lg = hid + kBlOwn * 2 * p.hh;                                                  // [8][32]
  float* part = lg + kBlOwn * 32;                                                       // [2][8][2*hh]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t crank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int c = (int)crank, cl = blockIdx.x / kBlCtas;
  const int row0 = cl * kBlRows;
  const int valid = min(kBlRows, p.rows - row0);
  const int rt = warp & 3, cg = warp >> 2;
  const uint32_t as_base = (uint32_t)__cvta_generic_to_shared(As), ws_base = (uint32_t)__cvta_generic_to_shared(Ws);
  const int Lb = p.Lb;
  for (int i = tid; i < kD; i += kBlThreads) q0s[i] = __bfloat162float(p.tab_qkv[(size_t)p.q_row * 3 * kD + i]);
  __syncthreads();

  for (int step = 0; step < p.nsteps; ++step) {
    // ---- P1: self-attention of the [LEN] row over the assigned slots (tabulated K / V), rows 8c .. 8c+7 -------------------
    for (int t = warp; t < kBlOwn * 8; t += kBlThreads / 32) {
      const int lr = c * kBlOwn + (t >> 3), head = t & 7, row = row0 + lr;
      if (lr >= valid || p.st.finished[row]) continue;
      const int nvis = min(p.st.last[row], Lb);
      float qf[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) qf[i] = q0s[head * kHeadDim + (lane & 3) * 16 + i];
      const int* ext = p.st.ext + (size_t)row * Lb;
      auto idx = [&](int j) { return ext[j] * Lb + j; };
      bl_attend(qf, p.tab_qkv + kD + head * kHeadDim, p.tab_qkv + 2 * kD + head * kHeadDim, (size_t)3 * kD, nvis, idx, ps_all + warp * kMaxKeys,
                p.ao + (size_t)row * kD + head * kHeadDim, lane);
    }
    bl_cluster_sync();
    // ---- P2: x = x0 + ao . Wso^T + b  (columns 64c .. 64c+63) ----------------------------------------------------------------
    {
      bl_stage_rows(As, p.ao + (size_t)row0 * kD, kD, valid, tid);
      bl_stage_rows(Ws, p.w_so + (size_t)(c * 64) * kD, kD, 64, tid);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      float acc[2][4] = {};
      bl_mma_512(acc, as_base, ws_base, rt, cg, lane);
      bl_epilogue<BL_EPI_X0>(acc, p.b_so, c * 64, row0, valid, rt, cg, lane, p.x, p.x0, nullptr, 0);
      __syncthreads();
    }
    bl_cluster_sync();
    // ---- P3: q = LN1(x) . Wq^T + b ---------------------------------------------------------------------------------------------
    {
      bl_stage_rows(Ws, p.w_q + (size_t)(c * 64) * kD, kD, 64, tid);
      bl_stage_layernorm(As, p.x, row0, valid, p.ln1_a, p.ln1_b, warp, lane);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      float acc[2][4] = {};
      bl_mma_512(acc, as_base, ws_base, rt, cg, lane);
      bl_epilogue<BL_EPI_BF16>(acc, p.b_q, c * 64, row0, valid, rt, cg, lane, nullptr, nullptr, p.q, kD);
      __syncthreads();
    }
    bl_cluster_sync();
    // ---- P4: cross-attention over the image's regions, rows 8c .. 8c+7 -----------------------------------------------------------
    for (int t = warp; t < kBlOwn * 8; t += kBlThreads / 32) {
      const int lr = c * kBlOwn + (t >> 3), head = t & 7, row = row0 + lr;
      if (lr >= valid || p.st.finished[row]) continue;
      const int img = row / p.sn;
      int tk = p.R;
      size_t kv0 = (size_t)img * p.R;
      if (p.mem_off) {
        const int o0 = p.mem_off[img];
        tk = min(p.R, p.mem_off[img + 1] - o0);
        kv0 = (size_t)o0;
      }
      const int nvis = min(p.mem_len ? p.mem_len[img] : tk, tk);
      float qf[16];
      {
        const bf16* qg = p.q + (size_t)row * kD + head * kHeadDim + (lane & 3) * 16;
        const uint4 q0 = __ldcg(reinterpret_cast<const uint4*>(qg)), q1 = __ldcg(reinterpret_cast<const uint4*>(qg + 8));
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
          qf[2 * i] = f.x;
          qf[2 * i + 1] = f.y;
        }
      }
      auto idx = [&](int j) { return j; };
      const bf16* kb = p.kv + kv0 * (size_t)(2 * kD) + head * kHeadDim;
      bl_attend(qf, kb, kb + kD, (size_t)2 * kD, nvis, idx, ps_all + warp * kMaxKeys, p.ao + (size_t)row * kD + head * kHeadDim, lane);
    }
    bl_cluster_sync();
    // ---- P5: x += ao . Wco^T + b ---------------------------------------------------------------------------------------------------
    {
      bl_stage_rows(As, p.ao + (size_t)row0 * kD, kD, valid, tid);
      bl_stage_rows(Ws, p.w_co + (size_t)(c * 64) * kD, kD, 64, tid);
      asm volatile("cp.async.wait_all;" ::: "memory");
      __syncthreads();
      float acc[2][4] = {};
      bl_mma_512(acc, as_base, ws_base, rt, cg, lane);
      bl_epilogue<BL_EPI_XADD>(acc, p.b_co, c * 64, row0, valid, rt, cg, lane, p.x, nullptr, nullptr, 0);
      __syncthreads();
    }
    bl_cluster_sync();
    // ---- P6: ffh = relu(LN2(x) . W1^T + b), this CTA's d_ff / 8 columns in slices of 64 ----------------------------------------------
    {
      const int per = p.d_ff / kBlCtas;
      bl_stage_layernorm(As, p.x, row0, valid, p.ln2_a, p.ln2_b, warp, lane);
      for (int s0 = 0; s0 < per; s0 += 64) {
        const int col0 = c * per + s0;
        bl_stage_rows(Ws, p.w_1 + (size_t)col0 * kD, kD, 64, tid);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        float acc[2][4] = {};
        bl_mma_512(acc, as_base, ws_base, rt, cg, lane);
        bl_epilogue<BL_EPI_RELU_BF16>(acc, p.b_1, col0, row0, valid, rt, cg, lane, nullptr, nullptr, p.ffh, (size_t)p.d_ff);
        __syncthreads();
      }
    }
    bl_cluster_sync();
    // ---- P7: x += ffh . W2^T + b, contraction in chunks of 512 -----------------------------------------------------------------------
    {
      float acc[2][4] = {};
      for (int k0 = 0; k0 < p.d_ff; k0 += kD) {
        bl_stage_rows(As, p.ffh + (size_t)row0 * p.d_ff + k0, (size_t)p.d_ff, valid, tid);
        bl_stage_rows(Ws, p.w_2 + (size_t)(c * 64) * p.d_ff + k0, (size_t)p.d_ff, 64, tid);
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        bl_mma_512(acc, as_base, ws_base, rt, cg, lane);
        __syncthreads();
      }
      bl_epilogue<BL_EPI_XADD>(acc, p.b_2, c * 64, row0, valid, rt, cg, lane, p.x, nullptr, nullptr, 0);
    }
    bl_cluster_sync();
    // ---- P8: length_predictor.norm + both heads + box rule for rows 8c .. 8c+7 (fp32, the arithmetic of bound_head_kernel) ------
    {
      const int hh2 = 2 * p.hh;
      if (warp < kBlOwn) {
        const int lr = c * kBlOwn + warp;
        if (lr < valid) {
          const float* xr = p.x + (size_t)(row0 + lr) * kD;
          float4 v[4];
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[i] = __ldcg(reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4));
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
          }
          const float mean = warp_sum(s) * (1.0f / kD);
          float ss = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            ss += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
          }
          const float denom = sqrtf(warp_sum(ss) * (1.0f / (kD - 1))) + 1e-6f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int cc = (i * 32 + lane) * 4;
            const float4 a = load4(p.lnh_a + cc), bb = load4(p.lnh_b + cc);
            float4 o;
            o.x = a.x * v[i].x / denom + bb.x;
            o.y = a.y * v[i].y / denom + bb.y;
            o.z = a.z * v[i].z / denom + bb.z;
            o.w = a.w * v[i].w / denom + bb.w;
            *reinterpret_cast<float4*>(hs + warp * kD + cc) = o;
          }
        } else {
          for (int cc = lane; cc < kD; cc += 32) hs[warp * kD + cc] = 0.f;
        }
      }
      __syncthreads();
      {   // classifier1 of both heads: 512 threads = 2 K-halves x 256 output columns
        const int o = tid & 255, kq = tid >> 8;
        float acc[kBlOwn];
#pragma unroll
        for (int r = 0; r < kBlOwn; ++r) acc[r] = 0.f;
        if (o < hh2) {
          for (int k0 = kq * (kD / 2); k0 < (kq + 1) * (kD / 2); k0 += 16) {
            float w[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) w[u] = p.w1t[(size_t)(k0 + u) * hh2 + o];
#pragma unroll
            for (int u = 0; u < 16; u += 4) {
#pragma unroll
              for (int r = 0; r < kBlOwn; ++r) {
                const float4 hv = *reinterpret_cast<const float4*>(hs + r * kD + k0 + u);
                acc[r] = fmaf(hv.x, w[u], acc[r]);
                acc[r] = fmaf(hv.y, w[u + 1], acc[r]);
                acc[r] = fmaf(hv.z, w[u + 2], acc[r]);
                acc[r] = fmaf(hv.w, w[u + 3], acc[r]);
              }
            }
          }
#pragma unroll
          for (int r = 0; r < kBlOwn; ++r) part[(kq * kBlOwn + r) * hh2 + o] = acc[r];
        }
      }
      __syncthreads();
      for (int i = tid; i < kBlOwn * hh2; i += kBlThreads) {
        const int o = i % hh2;
        hid[i] = fmaxf((part[i] + part[kBlOwn * hh2 + i]) + p.b1h[o], 0.f);
      }
      __syncthreads();
      int my_live = 0;
      if (warp < kBlOwn) {
        const int lr = c * kBlOwn + warp, b = row0 + lr;
        if (lr < valid) {
          const float* hr = hid + warp * hh2;
          if (lane < p.n_len + p.n_syn) {
            const bool is_len = lane < p.n_len;
            const float* w = is_len ? p.w_len + lane * p.hh : p.w_syn + (lane - p.n_len) * p.hh;
            const float* hx = is_len ? hr : hr + p.hh;
            float a = 0.f;
            for (int k = 0; k < p.hh; ++k) a = fmaf(hx[k], w[k], a);
            lg[warp * 32 + lane] = a + (is_len ? p.b_len[lane] : p.b_syn[lane - p.n_len]);
          }
          __syncwarp();
          if (lane == 0) {
            DecodeState st = p.st;
            if (st.finished[b]) {
              st.step_len[b] = 0;
            } else {
              atomicMax(&st.counters[1], step + 1);   // bounding iterations that still had a live row
              auto argmax_logp = [&](const float* z, int n) {       // log_softmax then torch.max: first maximal index
                float m = z[0];
                for (int i = 1; i < n; ++i) m = fmaxf(m, z[i]);
                float s = 0.f;
                for (int i = 0; i < n; ++i) s += expf(z[i] - m);
                const float lse = logf(s);
                int best = 0;
                float bv = (z[0] - m) - lse;
                for (int i = 1; i < n; ++i) {
                  const float v = (z[i] - m) - lse;
                  if (v > bv) { bv = v; best = i; }
                }
                return best;
              };
              int len_n = argmax_logp(lg + warp * 32, p.n_len);
              const int syn_n = argmax_logp(lg + warp * 32 + p.n_len, p.n_syn);
              const int last = st.last[b];
              if (len_n == 0 || syn_n < p.syn_lo || syn_n > p.syn_hi) {      // EOS (:1846)
                st.finished[b] = 1;
                st.step_len[b] = 0;
                atomicSub(st.counters, 1);
              } else {
                bool fin = false;
                if (len_n + last >= p.L + 1) {                              // clip to the last slot and finish (:1850-1859)
                  len_n = p.L + 1 - last;
                  st.finished[b] = 1;
                  atomicSub(st.counters, 1);
                  fin = true;
                }
                st.phrase_length[b * Lb + step] = len_n;
                st.phrase_syn[b * Lb + step] = syn_n;
                st.phrase_num[b] += 1;
                st.step_len[b] = len_n;
                for (int r = last; r < last + len_n; ++r) st.ext[b * Lb + r] = syn_n;
                const int nl = last + len_n;
                for (int r = last; r < Lb; ++r) st.vis[b * Lb + r] = nl;    // tgt_mask[j, last:, :last+len] = True
                st.vis[b * Lb] = nl;                                        // tgt_mask[j, 0, :last] = True
                st.last[b] = nl;
                if (!fin) my_live = 1;
              }
            }
          }
        }
      }
      // live rows of this CTA -> the cluster's slot (every CTA of the cluster sums the eight slots after the barrier)
      const int cta_live = __syncthreads_count(my_live);
      if (tid == 0) p.cl_live[cl * kBlCtas + c] = cta_live;
    }
    bl_cluster_sync();
    int alive = 0;
#pragma unroll
    for (int i = 0; i < kBlCtas; ++i) alive += __ldcg(p.cl_live + cl * kBlCtas + i);
    if (alive == 0) break;                        // the reference's `break` when every row has finished (:1869-1870), per cluster
  }
}

}  // namespace bofi
