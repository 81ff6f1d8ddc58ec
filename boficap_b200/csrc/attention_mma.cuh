// bf16 tensor-core attention for the short sequences of this path (<= 128 queries / keys per image):
// softmax(Q K^T / sqrt(64) with keys >= nvis masked) V, one CTA per (head, batch row), one warp per
// 16-query tile.  Q, K, V tiles are staged once in shared memory (row pitch 72 halves: conflict-free
// ldmatrix); S = Q K^T and O = P V run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with the whole
// score row block kept in registers, so the softmax never leaves the register file (no materialised
// mask, score or probability tensors; TransformerModel.py:1421-1432).
// (The sequences are far too short for a tcgen05/TMEM pipeline to pay: a 36 x 36 x 64 problem is
// 0.3 us of tensor work; the legacy warp-level MMA keeps all 4 SM sub-partitions busy instead.)
#pragma once
#include "common.cuh"

namespace bofi {

constexpr int kAttPitch = 72;   // halves per staged row (64 + 8 pad)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// KT = number of 16-key tiles (Tk <= 16*KT).
// 16-byte asynchronous global -> shared copy; `valid == false` writes zeros (src-size 0, the source is not read)
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}

template <int KT>
__global__ void __launch_bounds__(128, KT <= 3 ? 8 : 1)     // short key ranges: cap registers at 64 so that eight CTAs share an SM
attention_mma_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, const bf16* __restrict__ V, int ldkv,
                     bf16* __restrict__ O, int ldo, int Tq, int Tk, const int* __restrict__ vis, int vis_bs, int vis_qs,
                     int vis_div, int kv_div, float scale, const int* live_rows, Drop drop, const int* __restrict__ seq_off, int q_varlen) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  extern __shared__ __align__(16) uint8_t att_smem[];
  constexpr int TKP = KT * 16;
  bf16* Ks = reinterpret_cast<bf16*>(att_smem);          // [TKP][72]
  bf16* Vs = Ks + TKP * kAttPitch;                       // [TKP][72]
  bf16* Qs = Vs + TKP * kAttPitch;                       // [TQP][72]
  const int TQP = (Tq + 15) & ~15;
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Varlen layout (seq_off != nullptr): the K/V rows of sequence j are the compact rows [seq_off[j], seq_off[j+1]);
  // q_varlen: the queries are those same rows (encoder self-attention over the valid regions only).
  size_t kvrow0 = (size_t)(b / kv_div) * Tk, qrow0 = (size_t)b * Tq;
  if (seq_off) {
    const int o0 = seq_off[b / kv_div];
    Tk = min(Tk, seq_off[b / kv_div + 1] - o0);
    kvrow0 = (size_t)o0;
    if (q_varlen) { qrow0 = kvrow0; Tq = Tk; }
  }
  // All tiles are fetched with cp.async (16 bytes, L1 bypass; rows past the end zero-filled through the src-size
  // operand): every request of the CTA is in flight at once and the global latency is paid once.  The register
  // round trip it replaces (load, store, load, store ...) had the kernel stalled on long_scoreboard for 8 of every
  // 15 issue slots (ncu, 57 us for the encoder self-attention).
  for (int idx = tid; idx < TKP * 8; idx += 128) {       // 8 x 16-byte chunks per row
    const int j = idx >> 3, c = (idx & 7) * 8;
    const bool ok = j < Tk;
    const size_t src = (kvrow0 + (ok ? j : 0)) * ldkv + head * kHeadDim + c;
    cp_async_16(Ks + j * kAttPitch + c, K + src, ok);
    cp_async_16(Vs + j * kAttPitch + c, V + src, ok);
  }
  for (int idx = tid; idx < TQP * 8; idx += 128) {
    const int t = idx >> 3, c = (idx & 7) * 8;
    const bool ok = t < Tq;
    cp_async_16(Qs + t * kAttPitch + c, Q + (qrow0 + (ok ? t : 0)) * ldq + head * kHeadDim + c, ok);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  const uint32_t ks_base = (uint32_t)__cvta_generic_to_shared(Ks);
  const uint32_t vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
  const uint32_t qs_base = (uint32_t)__cvta_generic_to_shared(Qs);
  const int g = lane >> 2, t4 = lane & 3;

  for (int m0 = warp * 16; m0 < Tq; m0 += 64) {
    // ---- Q fragments (16 x 64) ----
    uint32_t qa[4][4];
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8, col = kt * 16 + (lane >> 4) * 8;
      ldmatrix_x4(qa[kt], qs_base + (uint32_t)(row * kAttPitch + col) * 2u);
    }
    // ---- S = Q K^T ----
    float sacc[2 * KT][4];
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) {
      sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
#pragma unroll
      for (int kp = 0; kp < 2; ++kp) {      // two k-tiles (32 dims) per ldmatrix.x4
        uint32_t kb[4];
        const int row = nt * 8 + (lane & 7), col = kp * 32 + (lane >> 3) * 8;
        ldmatrix_x4(kb, ks_base + (uint32_t)(row * kAttPitch + col) * 2u);
        mma_bf16_16816(sacc[nt], qa[2 * kp], kb[0], kb[1]);
        mma_bf16_16816(sacc[nt], qa[2 * kp + 1], kb[2], kb[3]);
      }
    }
    // ---- masked softmax over the row block (rows g and g+8 of this tile) ----
    const int r0 = m0 + g, r1 = m0 + g + 8;
    int nv0 = Tk, nv1 = Tk;
    if (vis) {
      const size_t vb = (size_t)(b / vis_div) * vis_bs;
      nv0 = min(vis[vb + (size_t)min(r0, Tq - 1) * vis_qs], Tk);
      nv1 = min(vis[vb + (size_t)min(r1, Tq - 1) * vis_qs], Tk);
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) {
      const int c = nt * 8 + 2 * t4;
      sacc[nt][0] = (c < nv0) ? sacc[nt][0] * scale : -INFINITY;
      sacc[nt][1] = (c + 1 < nv0) ? sacc[nt][1] * scale : -INFINITY;
      sacc[nt][2] = (c < nv1) ? sacc[nt][2] * scale : -INFINITY;
      sacc[nt][3] = (c + 1 < nv1) ? sacc[nt][3] * scale : -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float sub0 = (mx0 == -INFINITY) ? 0.f : mx0, sub1 = (mx1 == -INFINITY) ? 0.f : mx1;
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[KT][4];
#pragma unroll
    for (int nt = 0; nt < 2 * KT; ++nt) {
      const float p0 = __expf(sacc[nt][0] - sub0), p1 = __expf(sacc[nt][1] - sub0);
      const float p2 = __expf(sacc[nt][2] - sub1), p3 = __expf(sacc[nt][3] - sub1);
      sum0 += p0 + p1;
      sum1 += p2 + p3;
      float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;           // dropout on the probabilities (training only)
      if (drop.thresh) {
        const int c = nt * 8 + 2 * t4;
        k0 = drop_mul(drop, att_idx((int)qrow0 + r0, head, c)); k1 = drop_mul(drop, att_idx((int)qrow0 + r0, head, c + 1));
        k2 = drop_mul(drop, att_idx((int)qrow0 + r1, head, c)); k3 = drop_mul(drop, att_idx((int)qrow0 + r1, head, c + 1));
      }
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack2_bf16(p0 * k0, p1 * k1);     // a0a1 / a4a5 : row g
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2_bf16(p2 * k2, p3 * k3);     // a2a3 / a6a7 : row g+8
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // ---- O = P V ----
    float oacc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {      // two 8-dim n-tiles per ldmatrix.x4.trans
        uint32_t vb4[4];
        const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = np * 16 + (lane >> 4) * 8;
        ldmatrix_x4_trans(vb4, vs_base + (uint32_t)(row * kAttPitch + col) * 2u);
        mma_bf16_16816(oacc[2 * np], pa[kt], vb4[0], vb4[1]);
        mma_bf16_16816(oacc[2 * np + 1], pa[kt], vb4[2], vb4[3]);
      }
    }
    // all-masked row -> NaN, like softmax over -inf in the reference
    const float nanv = __int_as_float(0x7fc00000);
    const float inv0 = (nv0 <= 0) ? nanv : 1.f / sum0, inv1 = (nv1 <= 0) ? nanv : 1.f / sum1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = head * kHeadDim + nt * 8 + 2 * t4;
      if (r0 < Tq) *reinterpret_cast<uint32_t*>(O + (qrow0 + r0) * ldo + c) = pack2_bf16(oacc[nt][0] * inv0, oacc[nt][1] * inv0);
      if (r1 < Tq) *reinterpret_cast<uint32_t*>(O + (qrow0 + r1) * ldo + c) = pack2_bf16(oacc[nt][2] * inv1, oacc[nt][3] * inv1);
    }
  }
}

inline size_t attention_mma_smem_bytes(int KT, int Tq) {
  return (size_t)(2 * KT * 16 + ((Tq + 15) & ~15)) * kAttPitch * 2;
}

// Single-query attention (the [LEN] row against the image regions in the bounding step): one warp per
// (row, head), K/V read straight from L2.  Arithmetic order mirrors attention_kernel (bit-identical to
// the generic kernel's row in fp32).
template <typename T>
__global__ void __launch_bounds__(256)
attention_row_kernel(const T* __restrict__ Q, int ldq, const T* __restrict__ K, const T* __restrict__ V, int ldkv,
                     T* __restrict__ O, int ldo, int Tk, const int* __restrict__ vis, int vis_div, int kv_div, float scale,
                     const int* live_rows, const int* __restrict__ finished, Drop drop, const int* __restrict__ rowmap, int rowmap_div,
                     const int* rows_dev, const int* __restrict__ seq_off) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  if (finished && finished[blockIdx.x]) return;
  if (rows_dev && (int)blockIdx.x >= *rows_dev) return;
  __shared__ float qs[8][kHeadDim];
  __shared__ float ps[8][kMaxKeys];
  const int b = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bs = rowmap ? rowmap[b] / rowmap_div : b;      // sequence the (compact) query row belongs to
  size_t kvrow0 = (size_t)(bs / kv_div) * Tk;
  if (seq_off) {                                           // varlen memory: compact K/V rows of the image
    const int o0 = seq_off[bs / kv_div];
    Tk = min(Tk, seq_off[bs / kv_div + 1] - o0);
    kvrow0 = (size_t)o0;
  }
  const T* qg = Q + (size_t)b * ldq + head * kHeadDim;
  qs[head][lane] = to_float<T>(qg[lane]);
  qs[head][lane + 32] = to_float<T>(qg[lane + 32]);
  __syncwarp();
  int nvis = vis ? vis[bs / vis_div] : Tk;
  nvis = min(nvis, Tk);
  float sc[kMaxKeys / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int jj = 0; jj < kMaxKeys / 32; ++jj) {
    const int j = jj * 32 + lane;
    float s = -INFINITY;
    if (j < nvis) {
      const T* kr = K + (kvrow0 + j) * ldkv + head * kHeadDim;
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < kHeadDim; c += 4) {
        const float4 kq = load4(kr + c);
        d = fmaf(qs[head][c], kq.x, d);
        d = fmaf(qs[head][c + 1], kq.y, d);
        d = fmaf(qs[head][c + 2], kq.z, d);
        d = fmaf(qs[head][c + 3], kq.w, d);
      }
      s = d * scale;
    }
    sc[jj] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int jj = 0; jj < kMaxKeys / 32; ++jj) {
    const int j = jj * 32 + lane;
    const float e = (j < nvis) ? expf(sc[jj] - mx) : 0.f;
    sc[jj] = e;
    sum += e;
  }
  sum = warp_sum(sum);
#pragma unroll
  for (int jj = 0; jj < kMaxKeys / 32; ++jj) {
    const int j = jj * 32 + lane;
    if (j < Tk) ps[head][j] = sc[jj] / sum * drop_mul(drop, att_idx(b, head, j));
  }
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  // V rows come from L2: issue 8 independent row loads per trip (the adds stay in key order)
  for (int j0 = 0; j0 < nvis; j0 += 8) {
    float v0[8], v1[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = min(j0 + u, nvis - 1);
      const T* vr = V + (kvrow0 + j) * ldkv + head * kHeadDim;
      v0[u] = to_float<T>(vr[lane]);
      v1[u] = to_float<T>(vr[lane + 32]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (j0 + u < nvis) {
        const float pj = ps[head][j0 + u];
        o0 = fmaf(pj, v0[u], o0);
        o1 = fmaf(pj, v1[u], o1);
      }
    }
  }
  if (nvis <= 0) o0 = o1 = __int_as_float(0x7fc00000);
  T* og = O + (size_t)b * ldo + head * kHeadDim;
  og[lane] = from_float<T>(o0);
  og[lane + 32] = from_float<T>(o1);
}

// bf16 single-query attention tuned for HBM streaming (the K/V of `memory`, 2 KB per region, are re-read at
// every bounding step: 75 MB per step at B=1024, R=36).  4 rows x 8 heads per CTA, one warp per (row, head);
// every global access is a 16-byte load and all loads of a phase are independent:
//   scores : 4 lanes per key (16 dims each), 8 keys per pass            -> 2 loads / lane / pass
//   P.V    : 8 lanes per key (8 dims each), 4 keys in flight per trip   -> 1 load / lane / key
constexpr int kRowsPerCta = 4;
__global__ void __launch_bounds__(kRowsPerCta * 256)
attention_row_bf16_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, const bf16* __restrict__ V, int ldkv,
                          bf16* __restrict__ O, int ldo, int nb, int Tk, const int* __restrict__ vis, int vis_div, int kv_div,
                          float scale, const int* live_rows, const int* __restrict__ finished, Drop drop, const int* __restrict__ rowmap,
                          int rowmap_div, const int* rows_dev, const int* __restrict__ seq_off) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  __shared__ float ps[kRowsPerCta * 8][kMaxKeys];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = warp & 7;
  if (rows_dev) nb = min(nb, *rows_dev);   // compacted step: the valid-row count lives on the device
  for (int b = blockIdx.x * kRowsPerCta + (warp >> 3); b < nb; b += gridDim.x * kRowsPerCta) {
  if (finished && finished[b]) continue;   // bounding step: finished rows are ignored by the head
  const int bs = rowmap ? rowmap[b] / rowmap_div : b;      // sequence the (compact) query row belongs to
  size_t kvrow0 = (size_t)(bs / kv_div) * Tk;
  int tk = Tk;
  if (seq_off) {                                           // varlen memory: compact K/V rows of the image
    const int o0 = seq_off[bs / kv_div];
    tk = min(Tk, seq_off[bs / kv_div + 1] - o0);
    kvrow0 = (size_t)o0;
  }
  int nvis = vis ? vis[bs / vis_div] : tk;
  nvis = min(nvis, tk);
  // ---- scores ----
  const int sub = lane & 3, kslot = lane >> 2;
  float qf[16];
  {
    const bf16* qg = Q + (size_t)b * ldq + head * kHeadDim + sub * 16;
    const uint4 q0 = *reinterpret_cast<const uint4*>(qg), q1 = *reinterpret_cast<const uint4*>(qg + 8);
    const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      qf[2 * i] = f.x;
      qf[2 * i + 1] = f.y;
    }
  }
  float sc[kMaxKeys / 8];
  float mx = -INFINITY;
  const int npass = (nvis + 7) >> 3;
#pragma unroll
  for (int p = 0; p < kMaxKeys / 8; ++p) sc[p] = -INFINITY;
  // four passes (32 keys) per trip: their eight 16-byte key loads are issued together, then reduced -- one global
  // latency per 32 keys instead of one per 8 (this kernel sits on the bounding loop's dependency chain)
#pragma unroll
  for (int p0 = 0; p0 < kMaxKeys / 8; p0 += 4) {
    if (p0 >= npass) break;                    // uniform over the warp (one (row, head) per warp)
    uint4 k0[4], k1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = (p0 + u) * 8 + kslot;
      k0[u] = k1[u] = make_uint4(0u, 0u, 0u, 0u);
      if (j < nvis) {
        const bf16* kr = K + (kvrow0 + j) * ldkv + head * kHeadDim + sub * 16;
        k0[u] = *reinterpret_cast<const uint4*>(kr);
        k1[u] = *reinterpret_cast<const uint4*>(kr + 8);
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
      if (j < nvis) sc[p0 + u] = d * scale;
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
#pragma unroll
  for (int p = 0; p < kMaxKeys / 8; ++p) {
    if (p < npass && sub == 0) ps[warp][p * 8 + kslot] = sc[p] * inv * drop_mul(drop, att_idx(b, head, p * 8 + kslot));
  }
  __syncwarp();
  // ---- P.V ----
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
        vv[u] = *reinterpret_cast<const uint4*>(V + (kvrow0 + j) * ldkv + head * kHeadDim + g * 8);
        pj[u] = ps[warp][j];
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
    *reinterpret_cast<uint4*>(O + (size_t)b * ldo + head * kHeadDim + g * 8) = o;
  }
  __syncwarp();                            // ps[warp] is rewritten by the next row of this warp
  }
}

}  // namespace bofi
