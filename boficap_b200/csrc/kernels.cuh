// Memory-bound and short-sequence kernels of the BoFi decode path (sm_100a).
// Reference semantics cited per kernel (paths relative to /root/reference/captioning/models/).
#pragma once
#include "common.cuh"

namespace bofi {

// ---------------------------------------------------------------------------------------------
// LayerNorm of TransformerModel.py:1338-1349:  a_2 * (x - mean) / (std_unbiased + 1e-6) + b_2.
// NOT nn.LayerNorm: N-1 divisor and eps added to the standard deviation.
// One warp per 512-wide row, 4 x 128-bit loads per lane, shuffle reductions, two-pass variance.
// `in_stride` lets the caller normalise one row out of every block (the [LEN] row, :375).
// ---------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, size_t in_stride, const float* __restrict__ a2,
                 const float* __restrict__ b2, TOut* __restrict__ out, size_t out_stride, int rows,
                 float* __restrict__ out_f32_copy, const int* live_rows, const int* rows_dev) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  const int lane = threadIdx.x & 31;
  if (rows_dev) rows = min(rows, *rows_dev);           // compacted step: the valid-row count lives on the device
  // gain / bias of this lane's 16 columns: loaded once, the grid is persistent (a warp walks several rows)
  float4 ga[4], gb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ga[i] = load4(a2 + (i * 32 + lane) * 4);
    gb[i] = load4(b2 + (i * 32 + lane) * 4);
  }
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += (gridDim.x * blockDim.x) >> 5) {
  const float* xr = x + (size_t)row * in_stride;
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = load4(xr + (i * 32 + lane) * 4);
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
    const int c = (i * 32 + lane) * 4;
    const float4 a = ga[i], b = gb[i];
    float4 o;
    if (sizeof(TOut) == 4 || out_f32_copy) {  // fp32 parity path (and fp32 copies): the reference's a * (x - mean) / (std + eps) + b, divisions kept
      o.x = a.x * v[i].x / denom + b.x;
      o.y = a.y * v[i].y / denom + b.y;
      o.z = a.z * v[i].z / denom + b.z;
      o.w = a.w * v[i].w / denom + b.w;
    } else {                                  // bf16 output: one reciprocal per row + a residual correction per element
      const float inv = 1.0f / denom;
      o.x = ln_div(a.x * v[i].x, denom, inv) + b.x;
      o.y = ln_div(a.y * v[i].y, denom, inv) + b.y;
      o.z = ln_div(a.z * v[i].z, denom, inv) + b.z;
      o.w = ln_div(a.w * v[i].w, denom, inv) + b.w;
    }
    store4(out + (size_t)row * out_stride + c, o);
    if (out_f32_copy) store4(out_f32_copy + (size_t)row * kD + c, o);
  }
  }
}

// ---------------------------------------------------------------------------------------------
// Scaled-dot-product attention of TransformerModel.py:1421-1432 for short sequences, one CTA per
// (head, batch row).  The mask tensors of the reference ([B,1,R] region padding, [B,T,T] phrase
// masks) are all PREFIX masks, so a query only needs its visible-key count:
//     nvis = vis[(b / vis_div) * vis_bs + t * vis_qs]      (vis == nullptr -> all Tk keys)
// nvis == 0 reproduces the reference's all-masked softmax row: NaN.
// K/V of batch row b live at row (b / kv_div) (sample_n repeats share the image's memory).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
attention_kernel(const T* __restrict__ Q, int ldq, const T* __restrict__ K, const T* __restrict__ V, int ldkv,
                 T* __restrict__ O, int ldo, int Tq, int Tk, const int* __restrict__ vis, int vis_bs, int vis_qs,
                 int vis_div, int kv_div, float scale, const int* live_rows, Drop drop, const int* __restrict__ seq_off, int q_varlen) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  extern __shared__ float smem[];
  float* Ks = smem;                         // [Tk][65]
  float* Vs = Ks + ((Tk * (kHeadDim + 1) + 3) & ~3);   // [Tk][64], 16-byte aligned
  float* qs = Vs + Tk * kHeadDim;           // [4][64]
  float* ps = qs + 4 * kHeadDim;            // [4][kMaxKeys]
  const int head = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // varlen layout: see attention_mma_kernel (the shared-memory carve-up above keeps the padded Tk)
  size_t kvrow0 = (size_t)(b / kv_div) * Tk, qrow0 = (size_t)b * Tq;
  if (seq_off) {
    const int o0 = seq_off[b / kv_div];
    Tk = min(Tk, seq_off[b / kv_div + 1] - o0);
    kvrow0 = (size_t)o0;
    if (q_varlen) { qrow0 = kvrow0; Tq = Tk; }
  }
  for (int idx = tid; idx < Tk * (kHeadDim / 4); idx += 128) {
    const int j = idx >> 4, c = (idx & 15) * 4;
    const float4 kq = load4(K + (kvrow0 + j) * ldkv + head * kHeadDim + c);
    const float4 vq = load4(V + (kvrow0 + j) * ldkv + head * kHeadDim + c);
    float* kd = Ks + j * (kHeadDim + 1) + c;
    kd[0] = kq.x; kd[1] = kq.y; kd[2] = kq.z; kd[3] = kq.w;
    *reinterpret_cast<float4*>(Vs + j * kHeadDim + c) = vq;
  }
  __syncthreads();
  float* q = qs + warp * kHeadDim;
  float* p = ps + warp * kMaxKeys;
  for (int t = warp; t < Tq; t += 4) {
    const T* qg = Q + (qrow0 + t) * ldq + head * kHeadDim;
    q[lane] = to_float<T>(qg[lane]);
    q[lane + 32] = to_float<T>(qg[lane + 32]);
    __syncwarp();
    int nvis = vis ? vis[(size_t)(b / vis_div) * vis_bs + (size_t)t * vis_qs] : Tk;
    nvis = min(nvis, Tk);
    float sc[kMaxKeys / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < kMaxKeys / 32; ++jj) {
      const int j = jj * 32 + lane;
      float s = -INFINITY;
      if (j < nvis) {
        const float* kr = Ks + j * (kHeadDim + 1);
        float d = 0.f;
#pragma unroll 16
        for (int c = 0; c < kHeadDim; ++c) d = fmaf(q[c], kr[c], d);
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
      if (j < Tk) p[j] = sc[jj] / sum * drop_mul(drop, att_idx((int)qrow0 + t, head, j));
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < nvis; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, Vs[j * kHeadDim + lane], o0);
      o1 = fmaf(pj, Vs[j * kHeadDim + lane + 32], o1);
    }
    if (nvis <= 0) o0 = o1 = __int_as_float(0x7fc00000);   // softmax over an all -inf row
    T* og = O + (qrow0 + t) * ldo + head * kHeadDim;
    og[lane] = from_float<T>(o0);
    og[lane + 32] = from_float<T>(o1);
    __syncwarp();
  }
}

// Self-attention of the [LEN] row in the NAIC bounding layer when N_len == 1 (SURVEY.md Appendix A,
// "F_useful"): the layer input is a pure function of (syn id, position), so LN + Q/K/V projections of
// all 10 x Lb possible rows are tabulated per checkpoint (qkv_tab [10*Lb, 1536]); row 0 only attends
// to the assigned slots k < last (tgt_mask[j, 0, :last], TransformerModel.py:1859,1867).
// One CTA per batch row, one warp per head; arithmetic order mirrors attention_kernel so the result
// is bit-identical to running the full 22-row layer.
template <typename T>
__global__ void __launch_bounds__(256)
bound_self_attn_kernel(const T* __restrict__ qkv_tab, int Lb, int q_row, const int* __restrict__ ext,
                       const int* __restrict__ last, T* __restrict__ O, float scale, const int* live_rows,
                       const int* __restrict__ finished) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  if (finished[blockIdx.x]) return;        // the head ignores finished rows; their activations may go stale
  __shared__ float qs[8][kHeadDim];
  __shared__ int rowid[32];
  const int b = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvis = min(last[b], Lb);
  if (threadIdx.x < 32) rowid[lane] = (lane < Lb) ? ext[b * Lb + lane] * Lb + lane : 0;
  const T* qg = qkv_tab + (size_t)q_row * 3 * kD + head * kHeadDim;
  qs[head][lane] = to_float<T>(qg[lane]);
  qs[head][lane + 32] = to_float<T>(qg[lane + 32]);
  __syncthreads();
  float s = -INFINITY;
  if (lane < nvis) {
    const T* kr = qkv_tab + (size_t)rowid[lane] * 3 * kD + kD + head * kHeadDim;
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
  const float mx = warp_max(s);
  const float e = (lane < nvis) ? expf(s - mx) : 0.f;
  const float sum = warp_sum(e);
  const float p = e / sum;
  float o0 = 0.f, o1 = 0.f;
  for (int j0 = 0; j0 < nvis; j0 += 8) {       // 8 independent table-row loads per trip, adds in key order
    float v0[8], v1[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = min(j0 + u, nvis - 1);
      const T* vr = qkv_tab + (size_t)rowid[j] * 3 * kD + 2 * kD + head * kHeadDim;
      v0[u] = to_float<T>(vr[lane]);
      v1[u] = to_float<T>(vr[lane + 32]);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float pj = __shfl_sync(0xffffffffu, p, min(j0 + u, 31));
      if (j0 + u < nvis) {
        o0 = fmaf(pj, v0[u], o0);
        o1 = fmaf(pj, v1[u], o1);
      }
    }
  }
  T* og = O + (size_t)b * kD + head * kHeadDim;
  og[lane] = from_float<T>(o0);
  og[lane + 32] = from_float<T>(o1);
}

inline size_t attention_smem_bytes(int Tk) {
  return sizeof(float) * ((((size_t)Tk * (kHeadDim + 1) + 3) & ~(size_t)3) + (size_t)Tk * kHeadDim + 4 * kHeadDim + 4 * kMaxKeys);
}

// fp32 -> T, 128-bit vectorised (n % 4 == 0)
template <typename T>
__global__ void __launch_bounds__(256) cast_kernel(const float* __restrict__ in, T* __restrict__ out, size_t n4) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) store4(out + i * 4, load4(in + i * 4));
}

// bf16 -> fp32 (unit entry points only)
__global__ void __launch_bounds__(256) widen_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n4) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) store4(out + i * 4, load4(in + i * 4));
}

// out[c][r] = in[r][c]  (in: [rows][cols])
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
  pdl_enter();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}

// pack_wrapper (AttModel.py:46-51): rows of padded regions come back as zeros.
__global__ void __launch_bounds__(256)
zero_padded_rows_kernel(float* __restrict__ x, const int* __restrict__ att_len, int B, int R) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * R) return;
  const int b = row / R, r = row - b * R;
  if (r < att_len[b]) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) store4(x + (size_t)row * kD + (i * 32 + lane) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
}

// ---------------------------------------------------------------------------------------------
// Varlen encoder (pack_wrapper / PackedSequence, AttModel.py:33-51): the reference applies att_embed to the
// valid rows only and masks the padded keys in every attention; here the valid rows are COMPACTED once after att_embed
// (row offsets = exclusive scan of the valid-region counts) and the whole encoder, the memory K/V projections and the
// cross-attention K/V reads work on sum(att_len) rows.  No sort, no pack.
// ---------------------------------------------------------------------------------------------
// seq_off[0..B] = exclusive scan of clamp(att_len, 0, R); *total = seq_off[B].  One CTA.
__global__ void __launch_bounds__(1024) varlen_scan_kernel(const int* __restrict__ att_len, int B, int R, int* __restrict__ seq_off,
                                                           int* __restrict__ total) {
  pdl_enter();
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + tid;
    const int v = (i < B) ? min(max(att_len[i], 0), R) : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += n;
      }
      s_warp[lane] = w;                      // inclusive over the warps of this chunk
    }
    __syncthreads();
    const int carry = s_carry;
    const int excl = carry + (warp ? s_warp[warp - 1] : 0) + incl - v;
    if (i < B) seq_off[i] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = carry + s_warp[31];
    __syncthreads();
  }
  if (tid == 0) { seq_off[B] = s_carry; *total = s_carry; }
}

// compact:  xc[seq_off[b] + r] = xp[b*R + r] for r < len_b      scatter: out[b*R + r] = r < len_b ? xc[seq_off[b] + r] : 0
// (one warp per padded row; len_b = seq_off[b+1] - seq_off[b])
template <bool SCATTER>
__global__ void __launch_bounds__(256)
varlen_rows_kernel(const float* __restrict__ src, const int* __restrict__ seq_off, int B, int R, float* __restrict__ dst) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= B * R) return;
  const int b = row / R, r = row - b * R;
  const int o0 = seq_off[b], len = seq_off[b + 1] - o0;
  const bool valid = r < len;
  if (!SCATTER && !valid) return;
  const size_t crow = (size_t)o0 + r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (SCATTER) store4(dst + (size_t)row * kD + c, valid ? load4(src + crow * kD + c) : make_float4(0.f, 0.f, 0.f, 0.f));
    else store4(dst + crow * kD + c, load4(src + (size_t)row * kD + c));
  }
}

// att_masks [B, R] (f32, 0 / non-zero) -> valid-region counts, and a device-side check that every mask is a PREFIX mask
// (what dataloader.py:333-338 produces; the kernels only take counts).  One warp per image; *bad is set to 1 otherwise.
__global__ void __launch_bounds__(256)
masks_to_len_kernel(const float* __restrict__ masks, int B, int R, int* __restrict__ att_len, int* __restrict__ bad) {
  pdl_enter();
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  int n = 0;
  for (int r = lane; r < R; r += 32) n += masks[(size_t)b * R + r] != 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  bool ok = true;
  for (int r = lane; r < R; r += 32) ok &= ((masks[(size_t)b * R + r] != 0.f) == (r < n));
  if (!__all_sync(0xffffffffu, ok) && lane == 0) atomicExch(bad, 1);
  if (lane == 0) att_len[b] = n;
}

// fp16 -> T and bf16 -> fp32 feature conversions of the *_ex entry points (128-bit loads, n % 8 == 0)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) convert8_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, size_t n8) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n8; i += stride) {
    const uint4 raw = *reinterpret_cast<const uint4*>(in + i * 8);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    float f[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 t;
      if constexpr (std::is_same<TIn, __half>::value) t = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
      else t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[q]));
      f[2 * q] = t.x;
      f[2 * q + 1] = t.y;
    }
    store4(out + i * 8, make_float4(f[0], f[1], f[2], f[3]));
    store4(out + i * 8 + 4, make_float4(f[4], f[5], f[6], f[7]));
  }
}

// (id, position) input tables.  Embeddings (x sqrt(d), TransformerModel.py:1480-1487) + sinusoid
// (:1489-1507) are pure functions of (id, pos):
//   bound_in[s][p] = syn_embed[s]*sqrt(d) + pe[p]                              (:568)
//   fill_in [s][p] = (tgt_embed[bos]*sqrt(d) + syn_embed[s]*sqrt(d)) + pe[p]   (:572-577)
// built with the reference's own operation order so the rows are bit-identical.
__global__ void build_tables_kernel(const float* __restrict__ syn_lut, const float* __restrict__ tgt_lut,
                                    const float* __restrict__ pe, int bos, int n_syn, int Lb, int L,
                                    float sqrt_d, float* __restrict__ bound_in, float* __restrict__ fill_in) {
  pdl_enter();
  const int s = blockIdx.x, p = blockIdx.y;
  for (int c = threadIdx.x; c < kD; c += blockDim.x) {
    const float se = syn_lut[s * kD + c] * sqrt_d;
    bound_in[((size_t)s * Lb + p) * kD + c] = se + pe[p * kD + c];
    if (p < L) fill_in[((size_t)s * L + p) * kD + c] = (tgt_lut[(size_t)bos * kD + c] * sqrt_d + se) + pe[p * kD + c];
  }
}

// x[(b*T + r), :] = table[ids[b*ids_stride + ids_off + r]][r]   (one warp per row)
__global__ void __launch_bounds__(256)
gather_table_kernel(const float* __restrict__ table, int P, const int* __restrict__ ids, int ids_stride, int ids_off,
                    float* __restrict__ x, int rows, int T, const int* live_rows) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / T, r = row - b * T;
  const int id = ids[(size_t)b * ids_stride + ids_off + r];
  const float* src = table + ((size_t)id * P + r) * kD;
#pragma unroll
  for (int i = 0; i < 4; ++i) store4(x + (size_t)row * kD + (i * 32 + lane) * 4, load4(src + (i * 32 + lane) * 4));
}

// x[(b*T + r), :] = word_lut[w]*sqrt(d) (+ syn_lut[s]*sqrt(d)) + pe[r]   -- SAIC inputs, where the
// word id is data dependent (TransformerModel.py:518, :522).  syn_ids == nullptr -> word only.
__global__ void __launch_bounds__(256)
embed_words_kernel(const float* __restrict__ word_lut, const float* __restrict__ syn_lut, const float* __restrict__ pe,
                   const int* __restrict__ word_ids, const int* __restrict__ syn_ids, int ids_stride, int ids_off,
                   float sqrt_d, float* __restrict__ x, int rows, int T, const int* live_rows) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / T, r = row - b * T;
  const int w = word_ids[(size_t)b * ids_stride + ids_off + r];
  const int s = syn_ids ? syn_ids[(size_t)b * ids_stride + ids_off + r] : -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 e = load4(word_lut + (size_t)w * kD + c);
    e.x *= sqrt_d; e.y *= sqrt_d; e.z *= sqrt_d; e.w *= sqrt_d;
    if (s >= 0) {
      const float4 g = load4(syn_lut + (size_t)s * kD + c);
      e.x += g.x * sqrt_d; e.y += g.y * sqrt_d; e.z += g.z * sqrt_d; e.w += g.w * sqrt_d;
    }
    const float4 pp = load4(pe + (size_t)r * kD + c);
    e.x += pp.x; e.y += pp.y; e.z += pp.z; e.w += pp.w;
    store4(x + (size_t)row * kD + c, e);
  }
}

// ---------------------------------------------------------------------------------------------
// Per-row decode state (core_NAIC, TransformerModel.py:1825-1838; core_SAIC :1879-1905).
// ---------------------------------------------------------------------------------------------
struct DecodeState {
  int* ext;            // [rows, Lb] NAIC: syn id per slot (slot 0 = len_idx) ; SAIC: words fed to the bounding head
  int* ext_word;       // [rows, Lb] SAIC: words fed to the decoder
  int* ext_syn;        // [rows, Lb] SAIC: syn ids fed to the decoder
  int* seq22;          // [rows, Lb] SAIC: generated words incl. bos slot
  int* vis;            // [rows, Lb] visible-key count of every bounding row (prefix form of tgt_mask / len_mask)
  int* vis_fill;       // [rows, L]  visible-key count of every fill / decoder row
  int* last;           // [rows]     end of assigned slots (NAIC `last`, SAIC `phrase_last`)
  int* seq_last;       // [rows]     SAIC
  int* step_len;       // [rows]     phrase length accepted at the current step (0 = none)
  int* finished;       // [rows]
  int* phrase_num;     // [rows]
  int* phrase_length;  // [rows, Lb]
  int* phrase_syn;     // [rows, Lb]
  int* counters;       // [0] live rows, [1] bounding steps executed, [2] fill width w, [3] nan flag
};

__global__ void init_state_kernel(DecodeState st, int rows, int Lb, int L, int len_idx, int bos_idx, int saic) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) { st.counters[0] = rows; st.counters[1] = 0; st.counters[2] = 0; st.counters[3] = 0; st.counters[4] = rows; st.counters[5] = 0; }
  if (b >= rows) return;
  for (int r = 0; r < Lb; ++r) {
    st.ext[b * Lb + r] = (r == 0) ? len_idx : 0;
    st.ext_word[b * Lb + r] = 0;
    st.ext_syn[b * Lb + r] = 0;
    st.seq22[b * Lb + r] = (saic && r == 0) ? bos_idx : 0;
    st.vis[b * Lb + r] = 1;                       // tgt_mask[:, :, 0] = True
    st.phrase_length[b * Lb + r] = (saic && r == 0) ? 1 : 0;
    st.phrase_syn[b * Lb + r] = 0;
  }
  for (int r = 0; r < L; ++r) st.vis_fill[b * L + r] = 0;
  st.last[b] = 1;
  st.seq_last[b] = 0;
  st.step_len[b] = 0;
  st.finished[b] = 0;
  st.phrase_num[b] = 0;
}

// Bounding heads + box rule for one step, fused: final LayerNorm of the [LEN] row (length_predictor.norm,
// TransformerModel.py:374-375), classifier1 + ReLU of both heads (:376,:378; w1t = [512][2*Hh] = the two
// classifier1 weights concatenated and transposed), classifier2, log_softmax, first-max argmax (:377-382),
// then the EOS / clip / write rule (:1843-1867 NAIC, :1910-1926 SAIC).  8 rows per CTA, one warp per row.
// Always fp32: the calibrated heads amplify their input.
// NAIC (`saic == 0`) also writes the syn id into ext[last:last+len] and advances `last`/`vis`;
// SAIC defers that to saic_commit_kernel (the slots are filled with generated words).
constexpr int kHeadRows = 8;
__global__ void __launch_bounds__(1024)
bound_head_kernel(const float* __restrict__ x, size_t x_stride, const float* __restrict__ ln_a, const float* __restrict__ ln_b,
                  const float* __restrict__ w1t, const float* __restrict__ b1, int Hh,
                  const float* __restrict__ w_len, const float* __restrict__ b_len,
                  const float* __restrict__ w_syn, const float* __restrict__ b_syn, int n_len, int n_syn,
                  DecodeState st, int rows, int Lb, int L, int step_col, int step_no, int syn_lo, int syn_hi, int saic) {
  pdl_enter();
  if (st.counters[0] == 0) return;
  extern __shared__ float hsm[];
  float* h = hsm;                               // [8][512]  normalised [LEN] rows
  float* hid = h + kHeadRows * kD;              // [8][2*Hh]
  float* lg = hid + kHeadRows * 2 * Hh;         // [8][32]
  float* part = lg + kHeadRows * 32;            // [4][8][2*Hh]  K-quarter partial sums
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b0 = blockIdx.x * kHeadRows;
  if (warp < kHeadRows) {   // LayerNorm, same arithmetic as layernorm_kernel
    const int b = b0 + warp;
    if (b < rows) {
      const float* xr = x + (size_t)b * x_stride;
      float4 v[4];
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[i] = load4(xr + (i * 32 + lane) * 4);
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
        const int c = (i * 32 + lane) * 4;
        const float4 a = load4(ln_a + c), bb = load4(ln_b + c);
        float4 o;
        o.x = a.x * v[i].x / denom + bb.x;
        o.y = a.y * v[i].y / denom + bb.y;
        o.z = a.z * v[i].z / denom + bb.z;
        o.w = a.w * v[i].w / denom + bb.w;
        *reinterpret_cast<float4*>(h + warp * kD + c) = o;
      }
    } else {
      for (int c = lane; c < kD; c += 32) h[warp * kD + c] = 0.f;
    }
  }
  __syncthreads();
  // classifier1 of both heads: 1024 threads = 4 K-quarters x 256 output columns.  A thread walks its 128-deep
  // K slice of column o in batches of 32 independent coalesced loads (the weight lives in L2), so only four
  // round trips sit on the critical path; the four partial sums are combined through shared memory.
  {
    const int o = threadIdx.x & 255, kq = threadIdx.x >> 8;
    float acc[kHeadRows];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.f;
    if (o < 2 * Hh) {
      for (int k0 = kq * (kD / 4); k0 < (kq + 1) * (kD / 4); k0 += 32) {
        float w[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) w[u] = w1t[(size_t)(k0 + u) * 2 * Hh + o];
#pragma unroll
        for (int u = 0; u < 32; u += 4) {
#pragma unroll
          for (int r = 0; r < kHeadRows; ++r) {
            const float4 hv = *reinterpret_cast<const float4*>(h + r * kD + k0 + u);
            acc[r] = fmaf(hv.x, w[u], acc[r]);
            acc[r] = fmaf(hv.y, w[u + 1], acc[r]);
            acc[r] = fmaf(hv.z, w[u + 2], acc[r]);
            acc[r] = fmaf(hv.w, w[u + 3], acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) part[(kq * kHeadRows + r) * 2 * Hh + o] = acc[r];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kHeadRows * 2 * Hh; i += blockDim.x) {
    const int o = i % (2 * Hh);
    const float v = ((part[i] + part[kHeadRows * 2 * Hh + i]) + part[2 * kHeadRows * 2 * Hh + i]) + part[3 * kHeadRows * 2 * Hh + i];
    hid[i] = fmaxf(v + b1[o], 0.f);
  }
  __syncthreads();
  if (warp >= kHeadRows) return;
  const int b = b0 + warp;
  if (b >= rows) return;
  const float* hr = hid + warp * 2 * Hh;
  if (lane < n_len + n_syn) {
    const bool is_len = lane < n_len;
    const float* w = is_len ? w_len + lane * Hh : w_syn + (lane - n_len) * Hh;
    const float* hh = is_len ? hr : hr + Hh;
    float acc = 0.f;
    for (int c = 0; c < Hh; ++c) acc = fmaf(hh[c], w[c], acc);
    lg[warp * 32 + lane] = acc + (is_len ? b_len[lane] : b_syn[lane - n_len]);
  }
  __syncwarp();
  if (lane != 0) return;
  if (st.finished[b]) { st.step_len[b] = 0; return; }
  atomicMax(&st.counters[1], step_no);   // bounding iterations that still had a live row
  // log_softmax then torch.max: first maximal index
  auto argmax_logp = [&](const float* z, int n) {
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
  int len_n = argmax_logp(lg + warp * 32, n_len);
  const int syn_n = argmax_logp(lg + warp * 32 + n_len, n_syn);
  const int last = st.last[b];
  if (len_n == 0 || syn_n < syn_lo || syn_n > syn_hi) {
    st.finished[b] = 1;
    st.step_len[b] = 0;
    atomicSub(st.counters, 1);
    return;
  }
  if (len_n + last >= L + 1) {
    len_n = L + 1 - last;
    st.finished[b] = 1;
    atomicSub(st.counters, 1);
  }
  st.phrase_length[b * Lb + step_col] = len_n;
  st.phrase_syn[b * Lb + step_col] = syn_n;
  st.phrase_num[b] += 1;
  st.step_len[b] = len_n;
  if (!saic) {
    for (int r = last; r < last + len_n; ++r) st.ext[b * Lb + r] = syn_n;
    const int nl = last + len_n;
    for (int r = last; r < Lb; ++r) st.vis[b * Lb + r] = nl;   // tgt_mask[j, last:, :last+len] = True
    st.vis[b * Lb] = nl;                                        // tgt_mask[j, 0, :last] = True
    st.last[b] = nl;
  }
}

// NAIC fill window: every row uses w = last[rows-1] - 1 (the reference's stale loop index,
// TransformerModel.py:1871-1873).  Also records w and the NaN-batch flag.
// shard_rows > 0: the rows are several batches of shard_rows rows decoded in one call (bofi_set_shard); every batch keeps ITS
// window, w = last[last row of the row's batch] - 1, i.e. exactly what a stand-alone decode of that batch uses.  counters[2]
// reports the last batch's window, counters[3] whether any batch's window is empty.
__global__ void fill_window_kernel(DecodeState st, int rows, int L, int shard_rows) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    const int wl = st.last[rows - 1] - 1;
    st.counters[2] = wl;
    int bad = (wl <= 0) ? 1 : 0;
    if (shard_rows > 0)
      for (int end = shard_rows - 1; end < rows; end += shard_rows) bad |= (st.last[end] - 1 <= 0) ? 1 : 0;
    st.counters[3] = bad;
  }
  if (i >= rows * L) return;
  int end = rows - 1;
  if (shard_rows > 0) end = min(rows - 1, (i / L / shard_rows + 1) * shard_rows - 1);
  st.vis_fill[i] = st.last[end] - 1;
}

// ---------------------------------------------------------------------------------------------
// Vocab epilogue: log_softmax over V (AttModel.py:206-209), greedy first-max argmax with torch's
// NaN-is-maximal rule (CaptionModel.py:389-390) and tail padding seq[b, sum(phrase_length[b]):] = 0
// (AttModel.py:422-423).  One CTA per (row, slot); logits are read from the 16-byte-pitched
// workspace, log-probs are written with coalesced 4-byte stores to the caller's [rows, L, V] tensor
// (V = 9491 rows are not 16-byte aligned).
// ---------------------------------------------------------------------------------------------
// torch.max semantics: a NaN is the maximum, ties (and several NaNs) go to the lowest index.  (A branch-free form -- the
// value mapped to a sortable unsigned key, candidates chosen with selects -- was measured SLOWER in the vocabulary
// epilogue, 0.52 vs 0.44 ms: the common case here is "the running best stays", which the branches leave after two compares.)
struct ArgMax { float v; int i; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an || bn) {
    if (an && bn) return a.i < b.i ? a : b;
    return an ? a : b;
  }
  if (a.v > b.v) return a;
  if (b.v > a.v) return b;
  return a.i < b.i ? a : b;
}
// {last K values of prev, first 4 - K of own} on the output's 16-byte grid (see vocab_epilogue_kernel)
template <int K>
__device__ __forceinline__ void store_realigned(float* __restrict__ o, int c4, int n4, int V, const float (&prev)[4], const float (&own)[4]) {
  const int c = 4 * c4;
  if constexpr (K == 0) {
    if (c + 3 < V) __stcs(reinterpret_cast<float4*>(o + c), make_float4(own[0], own[1], own[2], own[3]));
    else {
#pragma unroll
      for (int q = 0; q < 4; ++q) if (c + q < V) __stcs(o + c + q, own[q]);
    }
  } else {
    float u[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) u[q] = q < K ? prev[q + 4 - K] : own[q - K];
    const int s0 = c - K;
    if (c4 > 0 && s0 + 3 < V) __stcs(reinterpret_cast<float4*>(o + s0), make_float4(u[0], u[1], u[2], u[3]));
    else {
#pragma unroll
      for (int q = 0; q < 4; ++q) if (s0 + q >= 0 && s0 + q < V) __stcs(o + s0 + q, u[q]);
    }
    if (c4 == n4 - 1) {          // the columns after the last aligned vector belong to nobody's "next" float4
#pragma unroll
      for (int q = 4 - K; q < 4; ++q) if (c + q < V) __stcs(o + c + q, own[q]);
    }
  }
}

constexpr int kVocabThreads = 512;
constexpr int kVocabVec = 5;    // float4 per thread: 512 threads x 5 x 4 = 10240 >= Vpad (9504)
#ifndef BOFI_VOCAB_OCC
#define BOFI_VOCAB_OCC 3
#endif
__global__ void __launch_bounds__(kVocabThreads, BOFI_VOCAB_OCC)
vocab_epilogue_kernel(const float* __restrict__ logits, int ldl, int V, float* __restrict__ logp_out,
                      long long* __restrict__ seq_out, const int* __restrict__ total_len, int total_off, int L,
                      int do_logsoftmax, int* __restrict__ tok_out_i32, Sampler sp, float* __restrict__ slot_entropy,
                      float* __restrict__ slot_logp, int row0, int fast_exp, int ldo) {
  pdl_enter();
  // The whole row lives in registers: logits are read from HBM exactly once (128-bit loads, all independent),
  // max / argmax / sum-exp are block reductions, and the log-probs are written once.
  __shared__ ArgMax s_am[kVocabThreads / 32];
  __shared__ float s_sum[kVocabThreads / 32];
  // `logits` holds the rows [row0, row0 + gridDim.x) of the step (the NAIC projection runs in L2-sized row chunks);
  // every other array is indexed by the global row
  const int row = row0 + blockIdx.x;       // b * L + t
  const int b = row / L, t = row - b * L;
  const float4* z4 = reinterpret_cast<const float4*>(logits + (size_t)blockIdx.x * ldl);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n4 = (V + 3) >> 2;             // ldl >= 4 * n4 (padded pitch); columns >= V are ignored below
  float4 v[kVocabVec];
#pragma unroll
  for (int i = 0; i < kVocabVec; ++i) {
    const int c4 = tid + i * kVocabThreads;
    v[i] = (c4 < n4) ? z4[c4] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (c4 == n4 - 1) {        // pad columns of the last float4 (never written by the GEMM): neutral for max and sum
      if (4 * c4 + 1 >= V) v[i].y = -INFINITY;
      if (4 * c4 + 2 >= V) v[i].z = -INFINITY;
      if (4 * c4 + 3 >= V) v[i].w = -INFINITY;
    }
  }
  // This thread's candidate: the maximum with FMNMX (fmaxf skips NaNs) and a NaN flag, then the lowest index holding it
  // (descending scan, equal values overwrite).  Pad / filler elements are -inf with indices above every real one, so
  // they can only stand when the whole row is -inf and are then overwritten by lower real indices.  A thread that saw a
  // NaN (torch.max: NaN wins, first one) takes the general compare chain; the cross-lane reduction uses it in any case.
  float tmax = -INFINITY;
  bool tnan = false;
#pragma unroll
  for (int i = 0; i < kVocabVec; ++i) {
    tmax = fmaxf(fmaxf(tmax, v[i].x), fmaxf(v[i].y, fmaxf(v[i].z, v[i].w)));
    tnan |= (v[i].x != v[i].x) | (v[i].y != v[i].y) | (v[i].z != v[i].z) | (v[i].w != v[i].w);
  }
  ArgMax am = {tmax, 0x7fffffff};
  if (!tnan) {
#pragma unroll
    for (int i = kVocabVec - 1; i >= 0; --i) {
      const int c = (tid + i * kVocabThreads) * 4;
      if (v[i].w == tmax) am.i = c + 3;
      if (v[i].z == tmax) am.i = c + 2;
      if (v[i].y == tmax) am.i = c + 1;
      if (v[i].x == tmax) am.i = c;
    }
  } else {
    am = ArgMax{-INFINITY, 0x7fffffff};
#pragma unroll
    for (int i = 0; i < kVocabVec; ++i) {
      const int c = (tid + i * kVocabThreads) * 4;
      const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c + q < V) am = better(am, ArgMax{e[q], c + q});
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax other = {__shfl_xor_sync(0xffffffffu, am.v, o), __shfl_xor_sync(0xffffffffu, am.i, o)};
    am = better(am, other);
  }
  if (lane == 0) s_am[warp] = am;
  __syncthreads();
  am = s_am[0];
#pragma unroll
  for (int w = 1; w < kVocabThreads / 32; ++w) am = better(am, s_am[w]);
  const float mx = am.v;
  int picked = am.i;
  if (sp.enabled) {                      // sampled token: a second block-wide argmax over the perturbed scores
    ArgMax sm = {-INFINITY, 0x7fffffff};
#pragma unroll
    for (int i = 0; i < kVocabVec; ++i) {
      const int c = (tid + i * kVocabThreads) * 4;
      const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c + q < V) sm = better(sm, ArgMax{gumbel_score(sp, e[q], row, c + q), c + q});
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax other = {__shfl_xor_sync(0xffffffffu, sm.v, o), __shfl_xor_sync(0xffffffffu, sm.i, o)};
      sm = better(sm, other);
    }
    __syncthreads();                     // s_am is reused
    if (lane == 0) s_am[warp] = sm;
    __syncthreads();
    sm = s_am[0];
#pragma unroll
    for (int w = 1; w < kVocabThreads / 32; ++w) sm = better(sm, s_am[w]);
    picked = sm.i;
  }
  float lse = 0.f;
  if ((logp_out && do_logsoftmax) || slot_entropy) {
    {
      float s = 0.f;                  // pad / filler elements are -inf: exp(-inf - mx) adds 0 (NaN exactly when the real ones do)
      if (fast_exp) {               // bf16 engine: the logits carry bf16 rounding (1e-2); ex2.approx is far inside that
#pragma unroll
        for (int i = 0; i < kVocabVec; ++i) s += __expf(v[i].x - mx) + __expf(v[i].y - mx) + __expf(v[i].z - mx) + __expf(v[i].w - mx);
      } else {
#pragma unroll
        for (int i = 0; i < kVocabVec; ++i) s += expf(v[i].x - mx) + expf(v[i].y - mx) + expf(v[i].z - mx) + expf(v[i].w - mx);
      }
      s = warp_sum(s);
      if (lane == 0) s_sum[warp] = s;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kVocabThreads / 32; ++w) tot += s_sum[w];
      lse = logf(tot);
    }
  }
  if (slot_entropy) {
    // eval_split post-processing (captioning/utils/eval_utils.py:183-184) without the [B, L, V] tensor:
    //   slot_entropy = -sum_v p_v log p_v ,  slot_logp = log p at the token written to seq
    float h = 0.f;
#pragma unroll
    for (int i = 0; i < kVocabVec; ++i) {
      const int c = (tid + i * kVocabThreads) * 4;
      const float e[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (c + q < V) { const float lp = (e[q] - mx) - lse; h -= expf(lp) * lp; }
    }
    h = warp_sum(h);
    __syncthreads();
    if (lane == 0) s_sum[warp] = h;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kVocabThreads / 32; ++w) tot += s_sum[w];
      int tok = picked;
      if (total_len && t >= total_len[b] + total_off) tok = 0;
      slot_entropy[row] = tot;
      slot_logp[row] = (logits[(size_t)blockIdx.x * ldl + tok] - mx) - lse;
    }
    __syncthreads();
  }
  if (logp_out) {
    float* o = logp_out + (size_t)row * (size_t)(ldo > 0 ? ldo : V);     // ldo: pitch of the caller's rows (bofi_decode_ex)
    // Rows of the caller's [rows, L, V] tensor are only 4-byte aligned (V = 9491).  Every thread re-cuts its float4 on the
    // 16-byte grid of the OUTPUT: with k = (address of the row) mod 4 floats, the aligned vector that ends inside this
    // thread's float4 is {last k values of the previous float4, first 4 - k of its own}.  The previous float4 comes from
    // the neighbouring lane (lane 0 re-reads it: an L1/L2 hit); 128-bit streamed stores, no staging, no barrier.
    const int k = (int)((reinterpret_cast<uintptr_t>(o) >> 2) & 3);
    auto xform = [&](float4 w) {
      if (do_logsoftmax) { w.x = (w.x - mx) - lse; w.y = (w.y - mx) - lse; w.z = (w.z - mx) - lse; w.w = (w.w - mx) - lse; }
      return w;
    };
#pragma unroll
    for (int i = 0; i < kVocabVec; ++i) {
      const int c4 = tid + i * kVocabThreads;
      const float4 w = xform(v[i]);
      float4 p;
      p.x = __shfl_up_sync(0xffffffffu, w.x, 1); p.y = __shfl_up_sync(0xffffffffu, w.y, 1);
      p.z = __shfl_up_sync(0xffffffffu, w.z, 1); p.w = __shfl_up_sync(0xffffffffu, w.w, 1);
      if (c4 >= n4) continue;
      if (lane == 0 && c4 > 0 && k != 0) p = xform(z4[c4 - 1]);
      const float own[4] = {w.x, w.y, w.z, w.w}, prev[4] = {p.x, p.y, p.z, p.w};
      switch (k) {               // uniform over the CTA
        case 0: store_realigned<0>(o, c4, n4, V, prev, own); break;
        case 1: store_realigned<1>(o, c4, n4, V, prev, own); break;
        case 2: store_realigned<2>(o, c4, n4, V, prev, own); break;
        default: store_realigned<3>(o, c4, n4, V, prev, own); break;
      }
    }
  }
  if (tid == 0) {
    int tok = picked;
    if (tok_out_i32) tok_out_i32[row] = tok;            // SAIC keeps the unpadded pick
    if (seq_out) {
      if (total_len && t >= total_len[b] + total_off) tok = 0;
      seq_out[row] = tok;
    }
  }
}

// Second half of the fused vocabulary projection (tc::VocabEpi, gemm_tc2.cuh): folds the 2 * ceil(V / 256) records of a row
// = (b, slot) into what vocab_epilogue_kernel computes from a materialised row -- the maximum with torch.max's rules (a NaN
// wins, the lowest index on ties), log sum exp(z - max), the greedy or Gumbel-max token, tail padding
// seq[b, sum(phrase_length[b]):] = 0, and optionally eval_split's per-slot entropy / log-prob.  One thread per row, every
// record array is read coalesced.
__global__ void __launch_bounds__(256)
vocab_merge_kernel(tc::VocabEpi ve, int nparts, int M, int L, float* __restrict__ mx_out, float* __restrict__ lse_out,
                   long long* __restrict__ seq_out, const int* __restrict__ total_len, int total_off, int* __restrict__ tok_out_i32,
                   float* __restrict__ slot_entropy, float* __restrict__ slot_logp) {
  pdl_enter();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  float m = -INFINITY, gv = -INFINITY, gz = 0.f;
  int idx = 0x7fffffff, nan_at = 0x7fffffff, gi = 0x7fffffff;
  for (int p = 0; p < nparts; ++p) {
    const size_t at = (size_t)p * M + row;
    const float mp = ve.pm[at];
    const int ip = ve.pi[at];
    nan_at = min(nan_at, ve.pn[at]);
    if (mp > m || (mp == m && ip < idx)) { m = mp; idx = ip; }     // the two halves of a tile interleave their columns
    if (ve.sp.enabled) {
      const float g = ve.pgv[at];
      const int gip = ve.pgi[at];
      if (g > gv || (g == gv && gip < gi)) { gv = g; gi = gip; gz = ve.pgz[at]; }
    }
  }
  float S = 0.f, Tz = 0.f;
  for (int p = 0; p < nparts; ++p) {
    const size_t at = (size_t)p * M + row;
    const float mp = ve.pm[at], sp_ = ve.ps[at];
    if (mp == -INFINITY && sp_ == 0.f) continue;                  // a record without valid columns (the last, partial tile)
    const float w = ve.fast_exp ? __expf(mp - m) : expf(mp - m);
    S = fmaf(sp_, w, S);
    if (ve.pt) Tz = fmaf(ve.pt[at], w, Tz);
  }
  const bool has_nan = nan_at != 0x7fffffff;
  const float mx = has_nan ? __int_as_float(0x7fc00000) : m;
  const float lse = ve.need_lse ? logf(S) : 0.f;
  mx_out[row] = mx;
  lse_out[row] = lse;
  int tok = ve.sp.enabled ? gi : (has_nan ? nan_at : idx);
  float ztok = ve.sp.enabled ? gz : mx;
  if (tok_out_i32) tok_out_i32[row] = tok;
  const int b = row / L, t = row - b * L;
  if (total_len && t >= total_len[b] + total_off) { tok = 0; ztok = ve.pz0 ? ve.pz0[row] : ztok; }
  if (seq_out) seq_out[row] = tok;
  if (slot_entropy) {
    // -sum p log p = (max + lse) - sum p z ;  log p(token) = (z_token - max) - lse
    slot_entropy[row] = (mx + lse) - Tz / S;
    slot_logp[row] = (ztok - mx) - lse;
  }
}

// ---------------------------------------------------------------------------------------------
// SAIC (core_SAIC, TransformerModel.py:1878-1986).  counters[4] is the live-row count at the START of the
// step: a row that finishes by clipping in this step still has its last phrase decoded (:1917-1922), so the
// kernels of a step test the snapshot, not the running counter.  counters[5] is the "phrase nan!" flag.
// ---------------------------------------------------------------------------------------------
__global__ void saic_snapshot_kernel(DecodeState st) {
  pdl_enter();
  st.counters[4] = st.counters[0];
  st.counters[6] = 0;          // compact (row, slot) count of the step
}

// Decoder input of the phrase accepted at step i (:1928-1948): syn label + position-wise copy of the previous
// phrase's words (last n words when n <= m, otherwise each word stretched ct or ct+1 times), and the
// phrase-block-causal mask phrase_mask[p:, :p+n] = True in its prefix-count form.
__global__ void saic_prepare_kernel(DecodeState st, int rows, int Lb, int L, int step) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= rows) return;
  const int n = st.step_len[b];
  if (n == 0) return;
  const int p = st.last[b], m = st.phrase_length[b * Lb + step - 1], s0 = st.seq_last[b];
  const int syn = st.phrase_syn[b * Lb + step];
  for (int k = 0; k < n; ++k) st.ext_syn[b * Lb + p + k] = syn;
  if (n <= m) {
    for (int k = 0; k < n; ++k) st.ext_word[b * Lb + p + k] = st.seq22[b * Lb + s0 + (m - n) + k];
  } else {
    const int pre_less = m - (n % m), ct = n / m;
    int w = 0;
    for (int k = 0; k < m; ++k) {
      const int rep = (k < pre_less) ? ct : ct + 1;
      const int tokv = st.seq22[b * Lb + s0 + k];
      for (int r = 0; r < rep; ++r) st.ext_word[b * Lb + p + w++] = tokv;
    }
  }
  // decoder rows q = slot-1 >= p-1 see decoder keys k < p+n-1  (phrase_mask[1:-1, 1:-1])
  for (int q = p - 1; q < L; ++q) st.vis_fill[b * L + q] = p + n - 1;
}

// Per (row, slot): first-max argmax (NaN maximal), max and log-sum-exp of the logits; raises the NaN flag
// (the reference aborts the whole batch when any log-prob of the step is NaN, :1956-1958).
__global__ void __launch_bounds__(256)
vocab_stats_kernel(const float* __restrict__ logits, int ldl, int V, int* __restrict__ tok, float* __restrict__ mx_out,
                   float* __restrict__ lse_out, DecodeState st, Sampler sp, const int* __restrict__ cidx) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  __shared__ ArgMax s_am[8];
  __shared__ float s_sum[8];
  // compact step (cidx != nullptr): logits / mx / lse rows are compact, the pick goes to the dense (row, slot) index
  if (cidx && (int)blockIdx.x >= st.counters[6]) return;
  const int zrow = blockIdx.x;
  const int row = cidx ? cidx[zrow] : zrow;
  const float* z = logits + (size_t)zrow * ldl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  ArgMax am = {-INFINITY, 0x7fffffff};
  for (int c = tid; c < V; c += 256) am = better(am, ArgMax{z[c], c});
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax other = {__shfl_xor_sync(0xffffffffu, am.v, o), __shfl_xor_sync(0xffffffffu, am.i, o)};
    am = better(am, other);
  }
  if (lane == 0) s_am[warp] = am;
  __syncthreads();
  am = s_am[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) am = better(am, s_am[w]);
  float s = 0.f;
  for (int c = tid; c < V; c += 256) s += expf(z[c] - am.v);
  s = warp_sum(s);
  if (lane == 0) s_sum[warp] = s;
  __syncthreads();
  int picked = am.i;
  if (sp.enabled) {
    ArgMax sm = {-INFINITY, 0x7fffffff};
    for (int c = tid; c < V; c += 256) sm = better(sm, ArgMax{gumbel_score(sp, z[c], row, c), c});
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ArgMax other = {__shfl_xor_sync(0xffffffffu, sm.v, o), __shfl_xor_sync(0xffffffffu, sm.i, o)};
      sm = better(sm, other);
    }
    __syncthreads();
    if (lane == 0) s_am[warp] = sm;
    __syncthreads();
    sm = s_am[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) sm = better(sm, s_am[w]);
    picked = sm.i;
  }
  if (tid == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_sum[w];
    tok[row] = picked;
    mx_out[zrow] = am.v;
    lse_out[zrow] = logf(tot);
    if (am.v != am.v) atomicExch(&st.counters[5], 1);
  }
}

// seq_logprobs[j, p:p+n] = phrase_logprobs[j, p-1:p-1+n]  (:1972) -- only the slots committed this step are
// ever written; the caller's tensor is zero-filled once per decode.
__global__ void __launch_bounds__(256)
saic_write_logp_kernel(const float* __restrict__ logits, int ldl, int V, const float* __restrict__ mx, const float* __restrict__ lse,
                       float* __restrict__ logp_out, DecodeState st, int L, int do_logsoftmax, const int* __restrict__ tok,
                       float* __restrict__ slot_entropy, float* __restrict__ slot_logp, const int* __restrict__ cidx) {
  pdl_enter();
  if (st.counters[4] == 0 || st.counters[5] != 0) return;
  int zrow = blockIdx.x, row = zrow;
  if (cidx) {                            // compact step: every compact row is a committed slot
    if (zrow >= st.counters[6]) return;
    row = cidx[zrow];
  } else {
    const int b = row / L, t = row - b * L;
    const int n = st.step_len[b], p = st.last[b];
    if (n == 0 || t < p - 1 || t >= p - 1 + n) return;
  }
  const float* z = logits + (size_t)zrow * ldl;
  float* o = logp_out + (size_t)row * V;
  const float m = mx[zrow], l = lse[zrow];
  if (logp_out) {
    if (do_logsoftmax) {
      for (int c = threadIdx.x; c < V; c += 256) o[c] = (z[c] - m) - l;
    } else {
      for (int c = threadIdx.x; c < V; c += 256) o[c] = z[c];
    }
  }
  if (slot_entropy) {
    __shared__ float s_h[8];
    float h = 0.f;
    for (int c = threadIdx.x; c < V; c += 256) { const float lp = (z[c] - m) - l; h -= expf(lp) * lp; }
    h = warp_sum(h);
    if ((threadIdx.x & 31) == 0) s_h[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) tot += s_h[w];
      // the slot this row's pick is committed to: seq[:, 1:-1] column t (TransformerModel.py:1968-1972)
      slot_entropy[row] = tot;
      slot_logp[row] = (z[tok[row]] - m) - l;
    }
  }
}

// Commit of the step (:1968-1977): generated words into seq / the bounding-head input, len_mask, counters.
__global__ void saic_advance_kernel(const int* __restrict__ tok, DecodeState st, int rows, int Lb, int L, int step) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (st.counters[5] != 0) {           // "phrase nan!": return the state as it is, no further steps
    if (b == 0) { st.counters[0] = 0; st.counters[3] = 1; }
    return;
  }
  if (b >= rows) return;
  const int n = st.step_len[b];
  if (n == 0) return;
  const int p = st.last[b];
  for (int k = 0; k < n; ++k) {
    const int w = tok[b * L + p - 1 + k];
    st.seq22[b * Lb + p + k] = w;
    st.ext[b * Lb + p + k] = w;
  }
  for (int r = p; r < Lb; ++r) st.vis[b * Lb + r] = p + n;   // len_mask[j, p:, :p+n] = True
  st.vis[b * Lb] = p + n;                                     // len_mask[j, 0, :phrase_last] = True
  st.last[b] = p + n;
  st.seq_last[b] += st.phrase_length[b * Lb + step - 1];
}

// ---- incremental SAIC (decode_saic_incremental, engine.cu) -------------------------------------------------------
// Under the phrase-block-causal mask a decoder slot only sees slots of its own and earlier phrases, and its input is
// fixed once its phrase is placed: hidden states of earlier phrases never change.  Each step therefore only computes the
// slots of the phrase accepted in that step (compact list of (row, slot)), with per-layer K/V caches for the rest.
// cidx[k] = row * L + slot for every new decoder slot of the step; counters[6] = their number (any order).
__global__ void saic_compact_kernel(DecodeState st, int rows, int L, int step, int* __restrict__ cidx) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= rows) return;
  const int n = st.step_len[b];
  // A row without a first phrase keeps an all-False phrase mask: the reference's decoder turns its rows into NaN and
  // core_SAIC aborts the whole batch at step 1 ("phrase nan!", TransformerModel.py:1956-1958).  Rows that did get a
  // first phrase see >= 1 key in every slot from then on, so this is the only mask-induced NaN.
  if (step == 1 && n == 0) atomicExch(&st.counters[5], 1);
  if (n == 0) return;
  const int p = st.last[b];
  const int base = atomicAdd(&st.counters[6], n);
  for (int k = 0; k < n; ++k) cidx[base + k] = b * L + p - 1 + k;
}

// x_c[k] = word_lut[w]*sqrt(d) + syn_lut[s]*sqrt(d) + pe[slot]   for the compact rows (decode_SA input, :520-523)
__global__ void __launch_bounds__(256)
embed_compact_kernel(const float* __restrict__ word_lut, const float* __restrict__ syn_lut, const float* __restrict__ pe,
                     DecodeState st, int Lb, int L, const int* __restrict__ cidx, float sqrt_d, float* __restrict__ x) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  const int lane = threadIdx.x & 31, mc = st.counters[6];
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < mc; k += (gridDim.x * blockDim.x) >> 5) {
    const int d = cidx[k], b = d / L, q = d - b * L;
    const int w = st.ext_word[b * Lb + 1 + q], sy = st.ext_syn[b * Lb + 1 + q];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 e = load4(word_lut + (size_t)w * kD + c);
      e.x *= sqrt_d; e.y *= sqrt_d; e.z *= sqrt_d; e.w *= sqrt_d;
      const float4 g = load4(syn_lut + (size_t)sy * kD + c);
      e.x += g.x * sqrt_d; e.y += g.y * sqrt_d; e.z += g.z * sqrt_d; e.w += g.w * sqrt_d;
      const float4 pp = load4(pe + (size_t)q * kD + c);
      e.x += pp.x; e.y += pp.y; e.z += pp.z; e.w += pp.w;
      store4(x + (size_t)k * kD + c, e);
    }
  }
}

// cache[cidx[k], 0:1024] = qkv_c[k, 512:1536]   (K | V of the new slots into the layer's cache)
template <typename T>
__global__ void __launch_bounds__(256)
saic_scatter_kv_kernel(const T* __restrict__ qkv_c, const int* __restrict__ cidx, DecodeState st, T* __restrict__ cache) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  const int lane = threadIdx.x & 31, mc = st.counters[6];
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < mc; k += (gridDim.x * blockDim.x) >> 5) {
    const T* src = qkv_c + (size_t)k * 3 * kD + kD;
    T* dst = cache + (size_t)cidx[k] * 2 * kD;
#pragma unroll
    for (int i = 0; i < 8; ++i) store4(dst + (i * 32 + lane) * 4, load4(src + (i * 32 + lane) * 4));
  }
}

// Self-attention of the compact query rows against their sequence's cached K/V (keys < vis_fill[row, slot]).
// One CTA (8 warps = 8 heads) per compact row; arithmetic order as attention_kernel.
template <typename T>
__global__ void __launch_bounds__(256)
saic_self_attn_kernel(const T* __restrict__ qkv_c, const T* __restrict__ cache, const int* __restrict__ cidx, DecodeState st, int L,
                      T* __restrict__ O, float scale) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  __shared__ float qs[8][kHeadDim];
  const int head = threadIdx.x >> 5, lane = threadIdx.x & 31, mc = st.counters[6];
  for (int k = blockIdx.x; k < mc; k += gridDim.x) {
  const int d = cidx[k], b = d / L;
  const int nvis = min(st.vis_fill[d], L);
  const T* qg = qkv_c + (size_t)k * 3 * kD + head * kHeadDim;
  qs[head][lane] = to_float<T>(qg[lane]);
  qs[head][lane + 32] = to_float<T>(qg[lane + 32]);
  __syncwarp();
  const T* base = cache + (size_t)b * L * 2 * kD;
  float s = -INFINITY;
  if (lane < nvis) {
    const T* kr = base + (size_t)lane * 2 * kD + head * kHeadDim;
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < kHeadDim; c += 4) {
      const float4 kq = load4(kr + c);
      dsum = fmaf(qs[head][c], kq.x, dsum);
      dsum = fmaf(qs[head][c + 1], kq.y, dsum);
      dsum = fmaf(qs[head][c + 2], kq.z, dsum);
      dsum = fmaf(qs[head][c + 3], kq.w, dsum);
    }
    s = dsum * scale;
  }
  const float mx = warp_max(s);
  const float e = (lane < nvis) ? expf(s - mx) : 0.f;
  const float sum = warp_sum(e);
  const float p = e / sum;
  float o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < nvis; ++j) {
    const float pj = __shfl_sync(0xffffffffu, p, j);
    const T* vr = base + (size_t)j * 2 * kD + kD + head * kHeadDim;
    o0 = fmaf(pj, to_float<T>(vr[lane]), o0);
    o1 = fmaf(pj, to_float<T>(vr[lane + 32]), o1);
  }
  if (nvis <= 0) o0 = o1 = __int_as_float(0x7fc00000);
  T* og = O + (size_t)k * kD + head * kHeadDim;
  og[lane] = from_float<T>(o0);
  og[lane + 32] = from_float<T>(o1);
  __syncwarp();                      // qs[head] is rewritten by the next compact row
  }
}

// ---- incremental SAIC bounding (N_len == 1): the bounding layer's input rows are the words generated so far, so its
// LN + K|V projection is cached per (row, slot) as well and only the [LEN] query row is evaluated per step.
// x0[0:512] = tgt_embed[len_idx]*sqrt(d) + pe[0]   (the constant [LEN] input row, TransformerModel.py:518)
__global__ void saic_len_row_kernel(const float* __restrict__ word_lut, const float* __restrict__ pe, int len_idx, float sqrt_d,
                                    float* __restrict__ x0) {
  pdl_enter();
  for (int c = threadIdx.x; c < kD; c += blockDim.x) x0[c] = word_lut[(size_t)len_idx * kD + c] * sqrt_d + pe[c];
}
// bcache[b*Lb + 0, :] = K|V of the [LEN] row for every sequence b
template <typename T>
__global__ void __launch_bounds__(256)
saic_bcache_init_kernel(const T* __restrict__ qkv_row, T* __restrict__ bcache, int rows, int Lb) {
  pdl_enter();
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= rows) return;
  T* dst = bcache + (size_t)b * Lb * 2 * kD;
#pragma unroll
  for (int i = 0; i < 8; ++i) store4(dst + (i * 32 + lane) * 4, load4(qkv_row + kD + (i * 32 + lane) * 4));
}
// x_c[k] = tgt_embed[tok]*sqrt(d) + pe[slot + 1] for the words picked in this step (bounding-layer input rows)
__global__ void __launch_bounds__(256)
embed_bound_compact_kernel(const float* __restrict__ word_lut, const float* __restrict__ pe, const int* __restrict__ tok, DecodeState st,
                           int L, const int* __restrict__ cidx, float sqrt_d, float* __restrict__ x) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  const int lane = threadIdx.x & 31, mc = st.counters[6];
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < mc; k += (gridDim.x * blockDim.x) >> 5) {
    const int d = cidx[k], b = d / L, q = d - b * L;
    const int w = tok[d];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 e = load4(word_lut + (size_t)w * kD + c);
      e.x *= sqrt_d; e.y *= sqrt_d; e.z *= sqrt_d; e.w *= sqrt_d;
      const float4 pp = load4(pe + (size_t)(q + 1) * kD + c);
      e.x += pp.x; e.y += pp.y; e.z += pp.z; e.w += pp.w;
      store4(x + (size_t)k * kD + c, e);
    }
  }
}
// bcache[(b*Lb + slot + 1), :] = kv_c[k, :]
template <typename T>
__global__ void __launch_bounds__(256)
saic_scatter_bkv_kernel(const T* __restrict__ kv_c, const int* __restrict__ cidx, DecodeState st, int L, int Lb, T* __restrict__ bcache) {
  pdl_enter();
  if (st.counters[4] == 0) return;
  const int lane = threadIdx.x & 31, mc = st.counters[6];
  for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < mc; k += (gridDim.x * blockDim.x) >> 5) {
    const int d = cidx[k], b = d / L, q = d - b * L;
    const T* src = kv_c + (size_t)k * 2 * kD;
    T* dst = bcache + ((size_t)b * Lb + q + 1) * 2 * kD;
#pragma unroll
    for (int i = 0; i < 8; ++i) store4(dst + (i * 32 + lane) * 4, load4(src + (i * 32 + lane) * 4));
  }
}
// Self-attention of the [LEN] row (constant query q0) over the cached K/V of the words generated so far (keys < last[b],
// len_mask[j, 0, :phrase_last], TransformerModel.py:1975).  One CTA per sequence, one warp per head.
template <typename T>
__global__ void __launch_bounds__(256)
saic_bound_self_attn_kernel(const T* __restrict__ q0, const T* __restrict__ bcache, int Lb, const int* __restrict__ last, T* __restrict__ O,
                            float scale, const int* live_rows, const int* __restrict__ finished) {
  pdl_enter();
  if (step_is_dead(live_rows)) return;
  if (finished[blockIdx.x]) return;
  __shared__ float qs[8][kHeadDim];
  const int b = blockIdx.x, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvis = min(last[b], Lb);
  qs[head][lane] = to_float<T>(q0[head * kHeadDim + lane]);
  qs[head][lane + 32] = to_float<T>(q0[head * kHeadDim + lane + 32]);
  __syncwarp();
  const T* base = bcache + (size_t)b * Lb * 2 * kD;
  float s = -INFINITY;
  if (lane < nvis) {
    const T* kr = base + (size_t)lane * 2 * kD + head * kHeadDim;
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
  const float mx = warp_max(s);
  const float e = (lane < nvis) ? expf(s - mx) : 0.f;
  const float sum = warp_sum(e);
  const float p = e / sum;
  float o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < nvis; ++j) {
    const float pj = __shfl_sync(0xffffffffu, p, j);
    const T* vr = base + (size_t)j * 2 * kD + kD + head * kHeadDim;
    o0 = fmaf(pj, to_float<T>(vr[lane]), o0);
    o1 = fmaf(pj, to_float<T>(vr[lane + 32]), o1);
  }
  T* og = O + (size_t)b * kD + head * kHeadDim;
  og[lane] = from_float<T>(o0);
  og[lane + 32] = from_float<T>(o1);
}

__global__ void export_seq_kernel(DecodeState st, int rows, int Lb, int L, long long* __restrict__ seq) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const int b = i / L, c = i - b * L;
  seq[i] = st.seq22[b * Lb + 1 + c];
}

// phrase_length[:, :L] / phrase_syn[:, :L] (i64) / phrase_num to the caller's tensors.
// col0 = 0 for NAIC (phrase_length[:, :-2]), 1 for SAIC (phrase_length[:, 1:-1]).
__global__ void export_boxes_kernel(DecodeState st, int rows, int Lb, int L, int col0, int* __restrict__ phrase_num,
                                    int* __restrict__ phrase_length, long long* __restrict__ phrase_syn) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * L) return;
  const int b = i / L, c = i - b * L;
  phrase_length[i] = st.phrase_length[b * Lb + col0 + c];
  phrase_syn[i] = st.phrase_syn[b * Lb + col0 + c];
  if (c == 0) phrase_num[b] = st.phrase_num[b];
}

}  // namespace bofi
