// bf16 tensor-core backward of the short-sequence attention (softmax(Q K^T / sqrt(64), keys >= nvis masked) V),
// same problem decomposition as the FFMA version in train_kernels.cuh: one CTA per (head, K/V block), `qpk * Tq`
// consecutive query rows attend to the block (one sequence for self-attention, the captions / bounding passes of one
// image for cross-attention).  128 threads; queries are processed in chunks of 64 (one 16-query tile per warp):
//
//   phase 1, warp = 16 queries     S = Q K^T (recomputed), P = softmax(S), dP = dO V^T, dS = P o (dP - rowsum(P o dP)),
//                                  dQ = dS K * scale  -> global;  P and dS (bf16) -> shared memory [query][key]
//   phase 2, warp = 16-key tiles   dK += dS^T Q * scale,  dV += P^T dO   (A operands read transposed with ldmatrix.trans;
//                                  accumulated in registers over all chunks: fixed summation order, no atomics)
//
// All five contractions run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate); the [Tq, Tk] score / probability
// matrices never leave the SM.  (Shapes are <= 128 x 128 x 64 per CTA: far below a tcgen05 tile.)
#pragma once
#include "attention_mma.cuh"

namespace bofi {

constexpr int kBwdQ = 64;       // queries per chunk

template <int KT>               // Tk <= 16 * KT
__global__ void __launch_bounds__(128)
attention_bwd_mma_kernel(const bf16* __restrict__ Q, int ldq, const bf16* __restrict__ K, const bf16* __restrict__ V, int ldkv,
                         const bf16* __restrict__ dO, int ldo, bf16* __restrict__ dQ, int lddq, bf16* __restrict__ dK, bf16* __restrict__ dV,
                         int lddkv, int Tq, int Tk, int qpk, const int* __restrict__ vis, int vis_bs, int vis_qs, int vis_div, float scale,
                         int accumulate_kv, Drop drop) {
  pdl_enter();
  extern __shared__ __align__(16) uint8_t bw_smem[];
  constexpr int TKP = KT * 16;
  constexpr int PP = TKP + 8;                              // pitch of the P / dS tiles (halves)
  bf16* Ks = reinterpret_cast<bf16*>(bw_smem);             // [TKP][72]
  bf16* Vs = Ks + TKP * kAttPitch;                         // [TKP][72]
  bf16* Qs = Vs + TKP * kAttPitch;                         // [64][72]
  bf16* Os = Qs + kBwdQ * kAttPitch;                       // [64][72]  dO
  bf16* Ps = Os + kBwdQ * kAttPitch;                       // [64][PP]
  bf16* Ds = Ps + kBwdQ * PP;                              // [64][PP]  dS
  const int head = blockIdx.x, kb = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  // tiles are fetched with cp.async as in attention_mma_kernel (all requests in flight at once, zero-filled tails);
  // the K/V copies are waited for together with the first Q/dO block
  for (int idx = tid; idx < TKP * 8; idx += 128) {
    const int j = idx >> 3, c = (idx & 7) * 8;
    const bool ok = j < Tk;
    const size_t src = ((size_t)kb * Tk + (ok ? j : 0)) * ldkv + head * kHeadDim + c;
    cp_async_16(Ks + j * kAttPitch + c, K + src, ok);
    cp_async_16(Vs + j * kAttPitch + c, V + src, ok);
  }
  const uint32_t ks_base = (uint32_t)__cvta_generic_to_shared(Ks), vs_base = (uint32_t)__cvta_generic_to_shared(Vs);
  const uint32_t qs_base = (uint32_t)__cvta_generic_to_shared(Qs), os_base = (uint32_t)__cvta_generic_to_shared(Os);
  const uint32_t ps_base = (uint32_t)__cvta_generic_to_shared(Ps), ds_base = (uint32_t)__cvta_generic_to_shared(Ds);

  // phase-2 accumulators: key tiles warp, warp+4 (KT <= 8)
  constexpr int MT = (KT + 3) / 4;
  float dkacc[MT][8][4], dvacc[MT][8][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < 8; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) dkacc[m][n][i] = dvacc[m][n][i] = 0.f;

  const int nq = qpk * Tq;
  for (int q0 = 0; q0 < nq; q0 += kBwdQ) {
    __syncthreads();                                       // K/V staged (first trip) / previous phase 2 finished
    for (int idx = tid; idx < kBwdQ * 8; idx += 128) {
      const int i = idx >> 3, c = (idx & 7) * 8;
      const bool ok = q0 + i < nq;
      const size_t r = (size_t)kb * nq + (ok ? q0 + i : 0);
      cp_async_16(Qs + i * kAttPitch + c, Q + r * ldq + head * kHeadDim + c, ok);
      cp_async_16(Os + i * kAttPitch + c, dO + r * ldo + head * kHeadDim + c, ok);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ------------------------------------------------------------------ phase 1
    {
      const int m0 = warp * 16;
      uint32_t qa[4][4], oa[4][4];
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const int row = m0 + (lane & 7) + ((lane >> 3) & 1) * 8, col = kt * 16 + (lane >> 4) * 8;
        ldmatrix_x4(qa[kt], qs_base + (uint32_t)(row * kAttPitch + col) * 2u);
        ldmatrix_x4(oa[kt], os_base + (uint32_t)(row * kAttPitch + col) * 2u);
      }
      float sacc[2 * KT][4], pacc[2 * KT][4];
#pragma unroll
      for (int nt = 0; nt < 2 * KT; ++nt) {
        sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f;
        pacc[nt][0] = pacc[nt][1] = pacc[nt][2] = pacc[nt][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t kb4[4], vb4[4];
          const int row = nt * 8 + (lane & 7), col = kp * 32 + (lane >> 3) * 8;
          ldmatrix_x4(kb4, ks_base + (uint32_t)(row * kAttPitch + col) * 2u);
          ldmatrix_x4(vb4, vs_base + (uint32_t)(row * kAttPitch + col) * 2u);
          mma_bf16_16816(sacc[nt], qa[2 * kp], kb4[0], kb4[1]);
          mma_bf16_16816(sacc[nt], qa[2 * kp + 1], kb4[2], kb4[3]);
          mma_bf16_16816(pacc[nt], oa[2 * kp], vb4[0], vb4[1]);
          mma_bf16_16816(pacc[nt], oa[2 * kp + 1], vb4[2], vb4[3]);
        }
      }
      const int r0 = q0 + m0 + g, r1 = r0 + 8;             // query indices inside the block
      int nv0 = 0, nv1 = 0;
      if (r0 < nq) {
        const int bq = kb * qpk + r0 / Tq, t = r0 % Tq;
        nv0 = vis ? min(vis[(size_t)(bq / vis_div) * vis_bs + (size_t)t * vis_qs], Tk) : Tk;
      }
      if (r1 < nq) {
        const int bq = kb * qpk + r1 / Tq, t = r1 % Tq;
        nv1 = vis ? min(vis[(size_t)(bq / vis_div) * vis_bs + (size_t)t * vis_qs], Tk) : Tk;
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
#pragma unroll
      for (int nt = 0; nt < 2 * KT; ++nt) {
        sacc[nt][0] = __expf(sacc[nt][0] - sub0); sacc[nt][1] = __expf(sacc[nt][1] - sub0);
        sacc[nt][2] = __expf(sacc[nt][2] - sub1); sacc[nt][3] = __expf(sacc[nt][3] - sub1);
        sum0 += sacc[nt][0] + sacc[nt][1];
        sum1 += sacc[nt][2] + sacc[nt][3];
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      const float inv0 = (nv0 > 0) ? 1.f / sum0 : 0.f, inv1 = (nv1 > 0) ? 1.f / sum1 : 0.f;
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 2 * KT; ++nt) {
        sacc[nt][0] *= inv0; sacc[nt][1] *= inv0; sacc[nt][2] *= inv1; sacc[nt][3] *= inv1;      // P
        if (drop.thresh) {                               // dP = d(P_dropped) o mask / (1 - p)
          const int c = nt * 8 + 2 * t4;
          const int g0 = kb * nq + r0, g1 = kb * nq + r1;
          pacc[nt][0] *= drop_mul(drop, att_idx(g0, head, c)); pacc[nt][1] *= drop_mul(drop, att_idx(g0, head, c + 1));
          pacc[nt][2] *= drop_mul(drop, att_idx(g1, head, c)); pacc[nt][3] *= drop_mul(drop, att_idx(g1, head, c + 1));
        }
        d0 += sacc[nt][0] * pacc[nt][0] + sacc[nt][1] * pacc[nt][1];
        d1 += sacc[nt][2] * pacc[nt][2] + sacc[nt][3] * pacc[nt][3];
      }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
      d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      uint32_t dsa[KT][4];
#pragma unroll
      for (int nt = 0; nt < 2 * KT; ++nt) {
        const float s0 = sacc[nt][0] * (pacc[nt][0] - d0), s1 = sacc[nt][1] * (pacc[nt][1] - d0);
        const float s2 = sacc[nt][2] * (pacc[nt][2] - d1), s3 = sacc[nt][3] * (pacc[nt][3] - d1);
        const uint32_t ds01 = pack2_bf16(s0, s1), ds23 = pack2_bf16(s2, s3);
        dsa[nt >> 1][(nt & 1) * 2 + 0] = ds01;
        dsa[nt >> 1][(nt & 1) * 2 + 1] = ds23;
        const int c = nt * 8 + 2 * t4;
        float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;     // dV uses the dropped probabilities
        if (drop.thresh) {
          const int g0 = kb * nq + r0, g1 = kb * nq + r1;
          k0 = drop_mul(drop, att_idx(g0, head, c)); k1 = drop_mul(drop, att_idx(g0, head, c + 1));
          k2 = drop_mul(drop, att_idx(g1, head, c)); k3 = drop_mul(drop, att_idx(g1, head, c + 1));
        }
        *reinterpret_cast<uint32_t*>(Ps + (m0 + g) * PP + c) = pack2_bf16(sacc[nt][0] * k0, sacc[nt][1] * k1);
        *reinterpret_cast<uint32_t*>(Ps + (m0 + g + 8) * PP + c) = pack2_bf16(sacc[nt][2] * k2, sacc[nt][3] * k3);
        *reinterpret_cast<uint32_t*>(Ds + (m0 + g) * PP + c) = ds01;
        *reinterpret_cast<uint32_t*>(Ds + (m0 + g + 8) * PP + c) = ds23;
      }
      // dQ = dS . K
      float qacc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) qacc[nt][0] = qacc[nt][1] = qacc[nt][2] = qacc[nt][3] = 0.f;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          uint32_t kb4[4];
          const int row = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = np * 16 + (lane >> 4) * 8;
          ldmatrix_x4_trans(kb4, ks_base + (uint32_t)(row * kAttPitch + col) * 2u);
          mma_bf16_16816(qacc[2 * np], dsa[kt], kb4[0], kb4[1]);
          mma_bf16_16816(qacc[2 * np + 1], dsa[kt], kb4[2], kb4[3]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = head * kHeadDim + nt * 8 + 2 * t4;
        if (r0 < nq) *reinterpret_cast<uint32_t*>(dQ + ((size_t)kb * nq + r0) * lddq + c) = pack2_bf16(qacc[nt][0] * scale, qacc[nt][1] * scale);
        if (r1 < nq) *reinterpret_cast<uint32_t*>(dQ + ((size_t)kb * nq + r1) * lddq + c) = pack2_bf16(qacc[nt][2] * scale, qacc[nt][3] * scale);
      }
    }
    __syncthreads();
    // ------------------------------------------------------------------ phase 2
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const int kt = warp + 4 * m;                         // key tile of this warp
      if (kt < KT) {
#pragma unroll
        for (int qk = 0; qk < kBwdQ / 16; ++qk) {          // contraction over the chunk's queries, 16 at a time
          // A = (dS^T | P^T) tile [16 keys][16 queries], stored [query][key]: transposed 8x8 loads
          uint32_t da[4], pa[4];
          const int mat = lane >> 3, r = lane & 7;
          const int qrow = qk * 16 + r + (mat >> 1) * 8, kcol = kt * 16 + (mat & 1) * 8;
          ldmatrix_x4_trans(da, ds_base + (uint32_t)(qrow * PP + kcol) * 2u);
          ldmatrix_x4_trans(pa, ps_base + (uint32_t)(qrow * PP + kcol) * 2u);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            uint32_t qb4[4], ob4[4];
            const int row = qk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = np * 16 + (lane >> 4) * 8;
            ldmatrix_x4_trans(qb4, qs_base + (uint32_t)(row * kAttPitch + col) * 2u);
            ldmatrix_x4_trans(ob4, os_base + (uint32_t)(row * kAttPitch + col) * 2u);
            mma_bf16_16816(dkacc[m][2 * np], da, qb4[0], qb4[1]);
            mma_bf16_16816(dkacc[m][2 * np + 1], da, qb4[2], qb4[3]);
            mma_bf16_16816(dvacc[m][2 * np], pa, ob4[0], ob4[1]);
            mma_bf16_16816(dvacc[m][2 * np + 1], pa, ob4[2], ob4[3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    const int kt = warp + 4 * m;
    if (kt < KT) {
      const int j0 = kt * 16 + g, j1 = j0 + 8;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int c = head * kHeadDim + nt * 8 + 2 * t4;
        if (j0 < Tk) {
          bf16* pk = dK + ((size_t)kb * Tk + j0) * lddkv + c;
          bf16* pv = dV + ((size_t)kb * Tk + j0) * lddkv + c;
          float k0 = dkacc[m][nt][0] * scale, k1 = dkacc[m][nt][1] * scale, v0 = dvacc[m][nt][0], v1 = dvacc[m][nt][1];
          if (accumulate_kv) {
            const float2 ok = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pk));
            const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pv));
            k0 += ok.x; k1 += ok.y; v0 += ov.x; v1 += ov.y;
          }
          *reinterpret_cast<uint32_t*>(pk) = pack2_bf16(k0, k1);
          *reinterpret_cast<uint32_t*>(pv) = pack2_bf16(v0, v1);
        }
        if (j1 < Tk) {
          bf16* pk = dK + ((size_t)kb * Tk + j1) * lddkv + c;
          bf16* pv = dV + ((size_t)kb * Tk + j1) * lddkv + c;
          float k0 = dkacc[m][nt][2] * scale, k1 = dkacc[m][nt][3] * scale, v0 = dvacc[m][nt][2], v1 = dvacc[m][nt][3];
          if (accumulate_kv) {
            const float2 ok = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pk));
            const float2 ov = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(pv));
            k0 += ok.x; k1 += ok.y; v0 += ov.x; v1 += ov.y;
          }
          *reinterpret_cast<uint32_t*>(pk) = pack2_bf16(k0, k1);
          *reinterpret_cast<uint32_t*>(pv) = pack2_bf16(v0, v1);
        }
      }
    }
  }
}

inline size_t attention_bwd_mma_smem_bytes(int KT) {
  const int TKP = KT * 16;
  return (size_t)2 * (2 * TKP * kAttPitch + 2 * kBwdQ * kAttPitch + 2 * kBwdQ * (TKP + 8));
}

}  // namespace bofi
