// Kernels of the XE-training path: the teacher-forced forward pieces that the decode
// path does not have, and every backward kernel that is not a GEMM.  Dense contractions (forward, dgrad, wgrad)
// go through the same tcgen05 / FFMA GEMMs as inference (train.inl).
// Reference semantics: TransformerModel.py:413-565 (forward), :1338-1349 (LayerNorm), :1421-1432 (attention),
// losses.py:315-369 (criterion); gradients are those torch autograd derives from that code.
#pragma once
#include "common.cuh"

namespace bofi {

// ---------------------------------------------------------------------------------------------
// Index bookkeeping of get_predict_phrase_length_syn_SA / _NA (TransformerModel.py:476-513, :532-565):
//   vis_b[n, p]  keys visible to the [LEN] row at bounding pass p: 1 + sum_{i=1..min(p, phrase_num[n]-1)} phrase_length[n,i]
//   na_vis[n]    = last[n] - 1, the key window of the NA filling pass (:559-561)
//   word_seq     = labels with column 0 replaced by len_idx (:478-479)
// ---------------------------------------------------------------------------------------------
__global__ void xe_prepare_kernel(const int* __restrict__ labels, const int* __restrict__ phrase_num, const int* __restrict__ phrase_length,
                                  int N, int Tb, int P, int len_idx, int* __restrict__ word_seq, int* __restrict__ vis_b,
                                  int* __restrict__ na_vis) {
  pdl_enter();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  for (int t = 0; t < Tb; ++t) word_seq[n * Tb + t] = (t == 0) ? len_idx : labels[n * Tb + t];
  int last = 1;
  vis_b[n * P] = 1;
  for (int p = 1; p < P; ++p) {
    if (phrase_num[n] > p) last += phrase_length[n * Tb + p];
    vis_b[n * P + p] = last;
  }
  na_vis[n] = last - 1;
}

// x[(b*T + r), :] = [word_lut[w]*sqrt(d)] (+ syn_lut[s]*sqrt(d)) + pe[r]      (Embeddings :1480-1487, PositionalEncoding :1489-1507)
//   w = word_ids ? word_ids[b*w_stride + w_off + r] : const_word   (const_word < 0: no word term)
//   s = syn_ids  ? syn_ids [b*s_stride + s_off + r] : none
__global__ void __launch_bounds__(256)
embed_xe_kernel(const float* __restrict__ word_lut, const float* __restrict__ syn_lut, const float* __restrict__ pe,
                const int* __restrict__ word_ids, int w_stride, int w_off, int const_word,
                const int* __restrict__ syn_ids, int s_stride, int s_off, float sqrt_d, float* __restrict__ x, int rows, int T, Drop drop) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / T, r = row - b * T;
  const int w = word_ids ? word_ids[(size_t)b * w_stride + w_off + r] : const_word;
  const int s = syn_ids ? syn_ids[(size_t)b * s_stride + s_off + r] : -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (w >= 0) {
      e = load4(word_lut + (size_t)w * kD + c);
      e.x *= sqrt_d; e.y *= sqrt_d; e.z *= sqrt_d; e.w *= sqrt_d;
    }
    if (s >= 0) {
      const float4 g = load4(syn_lut + (size_t)s * kD + c);
      if (w >= 0) { e.x += g.x * sqrt_d; e.y += g.y * sqrt_d; e.z += g.z * sqrt_d; e.w += g.w * sqrt_d; }
      else { e.x = g.x * sqrt_d; e.y = g.y * sqrt_d; e.z = g.z * sqrt_d; e.w = g.w * sqrt_d; }
    }
    const float4 pp = load4(pe + (size_t)r * kD + c);
    e.x += pp.x; e.y += pp.y; e.z += pp.z; e.w += pp.w;
    if (drop.thresh) {                                  // PositionalEncoding.dropout (:1506)
      const uint32_t i0 = (uint32_t)row * kD + c;
      e.x *= drop_mul(drop, i0); e.y *= drop_mul(drop, i0 + 1); e.z *= drop_mul(drop, i0 + 2); e.w *= drop_mul(drop, i0 + 3);
    }
    store4(x + (size_t)row * kD + c, e);
  }
}

// x[i] *= mask(i) / (1 - p), in place (FFN inner dropout :1478, att_embed dropout :1646, head dropout :376).
template <typename T>
__global__ void __launch_bounds__(256) dropout_inplace_kernel(T* __restrict__ x, size_t n4, Drop drop) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 v = load4(x + i * 4);
    const uint32_t i0 = (uint32_t)(i * 4);
    v.x *= drop_mul(drop, i0); v.y *= drop_mul(drop, i0 + 1); v.z *= drop_mul(drop, i0 + 2); v.w *= drop_mul(drop, i0 + 3);
    store4(x + i * 4, v);
  }
}
// out[i] = in[i] * mask(i) / (1 - p)   (gradient of a dropout site; TIn / TOut = fp32 or the GEMM operand type)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) dropout_apply_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, size_t n4, Drop drop) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 v = load4(in + i * 4);
    const uint32_t i0 = (uint32_t)(i * 4);
    v.x *= drop_mul(drop, i0); v.y *= drop_mul(drop, i0 + 1); v.z *= drop_mul(drop, i0 + 2); v.w *= drop_mul(drop, i0 + 3);
    store4(out + i * 4, v);
  }
}
// SublayerConnection (:1351-1363) with dropout: x_out = x_in + dropout(z), z = the sub-layer's output (GEMM operand type)
template <typename T>
__global__ void __launch_bounds__(256)
drop_add_kernel(const T* __restrict__ z, const float* __restrict__ x_in, float* __restrict__ x_out, size_t n4, Drop drop) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 v = load4(z + i * 4);
    const float4 r = load4(x_in + i * 4);
    const uint32_t i0 = (uint32_t)(i * 4);
    v.x = r.x + v.x * drop_mul(drop, i0); v.y = r.y + v.y * drop_mul(drop, i0 + 1);
    v.z = r.z + v.z * drop_mul(drop, i0 + 2); v.w = r.w + v.w * drop_mul(drop, i0 + 3);
    store4(x_out + i * 4, v);
  }
}

// out[(n*P + p), :] = in[n * in_row_stride, 0:512]   (the [LEN] row of sequence n, repeated for every pass)
template <typename T>
__global__ void __launch_bounds__(256)
repeat_row_kernel(const T* __restrict__ in, size_t in_row_stride, T* __restrict__ out, int N, int P) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= N * P) return;
  const T* src = in + (size_t)(row / P) * in_row_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) store4(out + (size_t)row * kD + (i * 32 + lane) * 4, load4(src + (i * 32 + lane) * 4));
}

// Self-attention of the [LEN] row of sequence n at pass p against the first vis_b[n,p] rows of the sequence
// (tgt_mask[j, 0, :last], TransformerModel.py:497-499): q = qkv[n*Tb, 0:512], K/V = qkv[n*Tb + j, 512: / 1024:].
// One CTA per (n, p), one warp per head.  Same arithmetic order as attention_kernel.
template <typename T>
__global__ void __launch_bounds__(256)
xe_bound_self_attn_kernel(const T* __restrict__ qkv, int Tb, int P, const int* __restrict__ vis_b, T* __restrict__ O, float scale, Drop drop) {
  pdl_enter();
  __shared__ float qs[8][kHeadDim];
  const int row = blockIdx.x, n = row / P, head = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvis = min(vis_b[row], Tb);
  const T* base = qkv + (size_t)n * Tb * 3 * kD;
  const T* qg = base + head * kHeadDim;
  qs[head][lane] = to_float<T>(qg[lane]);
  qs[head][lane + 32] = to_float<T>(qg[lane + 32]);
  __syncwarp();
  float s = -INFINITY;
  if (lane < nvis) {
    const T* kr = base + (size_t)lane * 3 * kD + kD + head * kHeadDim;
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
  const float p = e / sum * drop_mul(drop, att_idx(row, head, lane));
  float o0 = 0.f, o1 = 0.f;
  for (int j = 0; j < nvis; ++j) {
    const float pj = __shfl_sync(0xffffffffu, p, j);
    const T* vr = base + (size_t)j * 3 * kD + 2 * kD + head * kHeadDim;
    o0 = fmaf(pj, to_float<T>(vr[lane]), o0);
    o1 = fmaf(pj, to_float<T>(vr[lane + 32]), o1);
  }
  T* og = O + (size_t)row * kD + head * kHeadDim;
  og[lane] = from_float<T>(o0);
  og[lane + 32] = from_float<T>(o1);
}

// classifier2 of both heads + log_softmax (TransformerModel.py:376-379), training layout:
//   hid [Mb, 200] = relu(classifier1) of the final-normalised [LEN] row, rows ordered (n, p)
//   len_logp[n, p, :20], syn_logp[n, p, :10] (row pitch (Tb-1) per sequence; slots p >= P stay zero)
__global__ void __launch_bounds__(128)
xe_head_logp_kernel(const float* __restrict__ hid, const float* __restrict__ w_len, const float* __restrict__ b_len,
                    const float* __restrict__ w_syn, const float* __restrict__ b_syn, int Hh, int n_len, int n_syn,
                    float* __restrict__ len_logp, float* __restrict__ syn_logp, int Mb, int P, int slots) {
  pdl_enter();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= Mb) return;
  const float* hr = hid + (size_t)row * 2 * Hh;
  float z = -INFINITY;
  const bool is_len = lane < n_len, valid = lane < n_len + n_syn;
  if (valid) {
    const float* w = is_len ? w_len + lane * Hh : w_syn + (lane - n_len) * Hh;
    const float* hh = is_len ? hr : hr + Hh;
    float acc = 0.f;
    for (int c = 0; c < Hh; ++c) acc = fmaf(hh[c], w[c], acc);
    z = acc + (is_len ? b_len[lane] : b_syn[lane - n_len]);
  }
  // two independent log-softmaxes inside one warp: lanes [0,n_len) and [n_len, n_len+n_syn)
  const float zl = is_len ? z : -INFINITY, zs = (valid && !is_len) ? z : -INFINITY;
  const float ml = warp_max(zl), ms = warp_max(zs);
  const float el = is_len ? expf(z - ml) : 0.f, es = (valid && !is_len) ? expf(z - ms) : 0.f;
  const float ll = logf(warp_sum(el)), ls = logf(warp_sum(es));
  const int n = row / P, p = row - n * P;
  if (is_len) len_logp[((size_t)n * slots + p) * n_len + lane] = (z - ml) - ll;
  else if (valid) syn_logp[((size_t)n * slots + p) * n_syn + (lane - n_len)] = (z - ms) - ls;
}

// Backward of the two log-softmax heads: dz = g - softmax * sum(g) for both heads of row (n,p) -> dz [Mb, 32]
// (cols 0..19 length head, 20..29 syntax head, 30..31 zero).
__global__ void __launch_bounds__(128)
xe_head_dz_kernel(const float* __restrict__ g_len, const float* __restrict__ g_syn, const float* __restrict__ len_logp,
                  const float* __restrict__ syn_logp, int n_len, int n_syn, float* __restrict__ dz, int Mb, int P, int slots) {
  pdl_enter();
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= Mb) return;
  const int n = row / P, p = row - n * P;
  const bool is_len = lane < n_len, valid = lane < n_len + n_syn;
  float g = 0.f, lp = 0.f;
  if (is_len) {
    g = g_len[((size_t)n * slots + p) * n_len + lane];
    lp = len_logp[((size_t)n * slots + p) * n_len + lane];
  } else if (valid) {
    g = g_syn[((size_t)n * slots + p) * n_syn + lane - n_len];
    lp = syn_logp[((size_t)n * slots + p) * n_syn + lane - n_len];
  }
  const float sl = warp_sum(is_len ? g : 0.f), ss = warp_sum((valid && !is_len) ? g : 0.f);
  float out = 0.f;
  if (is_len) out = g - expf(lp) * sl;
  else if (valid) out = g - expf(lp) * ss;
  dz[(size_t)row * 32 + lane] = out;
}

// classifier2 backward, two fixed-order stages.  Stage 1: one CTA per (output unit o, row chunk):
//   partial[chunk][o][c] = sum_{rows of chunk} dz[row, o] * hid[row, c (+Hh)]   (c == Hh holds the bias term sum dz[row, o])
// Stage 2: gW2[o, c] += sum over chunks, gb2[o] likewise.
__global__ void __launch_bounds__(128)
xe_head2_wgrad_partial_kernel(const float* __restrict__ dz, const float* __restrict__ hid, int Hh, int n_len, int Mb,
                              float* __restrict__ partial) {
  pdl_enter();
  const int o = blockIdx.x, c = threadIdx.x, n_out = gridDim.x;
  const bool is_len = o < n_len;
  const float* h = hid + (is_len ? 0 : Hh);
  const int per = (Mb + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(Mb, r0 + per);
  float acc = 0.f;
  for (int row = r0; row < r1; ++row) {
    const float d = dz[(size_t)row * 32 + o];
    acc = fmaf(d, c < Hh ? h[(size_t)row * 2 * Hh + c] : 1.0f, acc);
  }
  if (c <= Hh) partial[((size_t)blockIdx.y * n_out + o) * 128 + c] = acc;
}
__global__ void __launch_bounds__(128)
xe_head2_wgrad_reduce_kernel(const float* __restrict__ partial, int chunks, int Hh, int n_len, float* __restrict__ gw_len,
                             float* __restrict__ gb_len, float* __restrict__ gw_syn, float* __restrict__ gb_syn) {
  pdl_enter();
  const int o = blockIdx.x, c = threadIdx.x, n_out = gridDim.x;
  if (c > Hh) return;
  float t = 0.f;
  for (int ch = 0; ch < chunks; ++ch) t += partial[((size_t)ch * n_out + o) * 128 + c];
  const bool is_len = o < n_len;
  if (c < Hh) {
    float* gw = is_len ? gw_len + (size_t)o * Hh : gw_syn + (size_t)(o - n_len) * Hh;
    gw[c] += t;
  } else {
    if (is_len) gb_len[o] += t; else gb_syn[o - n_len] += t;
  }
}

// dhid[row, c] = (sum_o dz[row, o] * W2[o, c]) * (hid[row, c] > 0)      (classifier2 dgrad + ReLU backward)
__global__ void __launch_bounds__(256)
xe_head2_dgrad_kernel(const float* __restrict__ dz, const float* __restrict__ hid, const float* __restrict__ w_len,
                      const float* __restrict__ w_syn, int Hh, int n_len, int n_syn, float* __restrict__ dhid, int ld_out, int Mb, float gain) {
  pdl_enter();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Mb * 2 * Hh) return;
  const int row = (int)(i / (2 * Hh)), c = (int)(i - (size_t)row * 2 * Hh);
  float acc = 0.f;
  if (c < Hh) {
    for (int o = 0; o < n_len; ++o) acc = fmaf(dz[(size_t)row * 32 + o], w_len[o * Hh + c], acc);
  } else {
    for (int o = 0; o < n_syn; ++o) acc = fmaf(dz[(size_t)row * 32 + n_len + o], w_syn[o * Hh + c - Hh], acc);
  }
  dhid[(size_t)row * ld_out + c] = hid[i] > 0.f ? acc * gain : 0.f;
}

// rows (b, t) with t >= total[b] + off are set to zero (SAIC: slots that were never committed return zero log-probs,
// TransformerModel.py:1883; their gradient is zero as well).  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256)
zero_tail_rows_kernel_t(T* __restrict__ x, const int* __restrict__ total, int off, int rows, int Tt, int ld) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / Tt, t = row - b * Tt;
  if (t < total[b] + off) return;
  T* p = x + (size_t)row * ld;
  for (int c = lane; c < ld; c += 32) p[c] = from_float<T>(0.f);
}
__global__ void __launch_bounds__(256)
zero_tail_rows_kernel(float* __restrict__ x, const int* __restrict__ total, int off, int rows, int Tt, int ld) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / Tt, t = row - b * Tt;
  if (t < total[b] + off) return;
  float* p = x + (size_t)row * ld;
  for (int c = lane; c < ld; c += 32) p[c] = 0.f;
}
// Glancing inputs (EncoderDecoder_UIC.forward :445-461): per caption row, mismatch = (#words - #correct predictions) / #words over
// its real word slots, keep_prob = mismatch * glat_p there (0 elsewhere); slot t takes the ground-truth word when
// u(row, t) < keep_prob and bos otherwise.  u is the library's counter-based uniform (the reference draws torch.rand).  One warp
// per row.  labels [N, Tb] (words at 1..T), plen [N, Tb] (slot 0 = the bos pseudo-phrase), tok i64 [N, T] the no-grad predictions.
__global__ void __launch_bounds__(256)
glat_input_kernel(const int* __restrict__ labels, const int* __restrict__ plen, const long long* __restrict__ tok, int N, int T, int Tb, int bos,
                  float glat_p, uint32_t key, int* __restrict__ words) {
  pdl_enter();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  int len = 0;
  for (int j = lane; j < Tb; j += 32) len += plen[(size_t)n * Tb + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len += __shfl_xor_sync(0xffffffffu, len, o);
  len -= 1;
  int same = 0;
  for (int t = lane; t < T; t += 32) same += (t < len && (int)tok[(size_t)n * T + t] == labels[(size_t)n * Tb + 1 + t]) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) same += __shfl_xor_sync(0xffffffffu, same, o);
  const float mismatch = len > 0 ? (float)(len - same) / (float)len : 0.f;
  for (int t = lane; t < T; t += 32) {
    const float keep_prob = (t < len) ? mismatch * glat_p : 0.f;
    const float u = ((float)drop_hash(key, (uint32_t)(n * T + t)) + 0.5f) * (1.0f / 4294967296.0f);
    words[(size_t)n * T + t] = (u < keep_prob) ? labels[(size_t)n * Tb + 1 + t] : bos;
  }
}

// ---- N_len >= 2: the bounding passes in the reference's own formulation (every pass a full stack over all Tb rows) -------------
// Visible-key counts of pass p (get_predict_phrase_length_syn_{SA,NA}, TransformerModel.py:476-513 / :532-565): the mask grows by
// `tgt_mask[j, last:, :last+len] = True; tgt_mask[j, 0, :last+len] = True`, so at pass p row 0 and the rows at or beyond the frontier
// see vis_b[n, p] keys, and a row inside phrase k <= p keeps what it saw when its phrase was added: vis_b[n, k].
//   vis_b [N, P] (xe_prepare_kernel)  ->  vis_full [P][N][Tb]
__global__ void xe_vis_full_kernel(const int* __restrict__ vis_b, int N, int P, int Tb, int* __restrict__ vis_full) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * N * Tb) return;
  const int r = i % Tb, n = (i / Tb) % N, p = i / (Tb * N);
  const int front = vis_b[n * P + p];
  int v = front;
  if (r != 0 && r < front) {
    for (int k = 0; k <= p; ++k) {
      const int lk = vis_b[n * P + k];
      if (r < lk) { v = lk; break; }
    }
  }
  vis_full[i] = v;
}
// x3[(n, p)] = x[n, row 0]: the [LEN] row of pass p (LengthPredictor_UIC.forward :375 `output[:, 0, :]`)
__global__ void gather_len_rows_kernel(const float* __restrict__ x, int Tb, float* __restrict__ x3, int N, int P, int p) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;          // float4 index
  if (i >= N * (kD / 4)) return;
  const int n = i / (kD / 4), c = (i % (kD / 4)) * 4;
  store4(x3 + ((size_t)n * P + p) * kD + c, load4(x + (size_t)n * Tb * kD + c));
}
// Gradient of a pass's stack output: zero except row 0 of every caption, which takes d x3[(n, p)]; fp32 and operand-type copies.
template <typename T>
__global__ void scatter_len_rows_kernel(const float* __restrict__ dlen, int Tb, float* __restrict__ dx, T* __restrict__ dxT, int N, int P, int p) {
  pdl_enter();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index
  if (i >= (size_t)N * Tb * (kD / 4)) return;
  const int c = (int)(i % (kD / 4)) * 4;
  const size_t row = i / (kD / 4);
  const int r = (int)(row % Tb), n = (int)(row / Tb);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r == 0) v = load4(dlen + ((size_t)n * P + p) * kD + c);
  store4(dx + row * kD + c, v);
  store4(dxT + row * kD + c, v);
}
__global__ void add_inplace_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n4, int overwrite) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 a = load4(src + i * 4);
    if (!overwrite) { const float4 b = load4(dst + i * 4); a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
    store4(dst + i * 4, a);
  }
}

__global__ void clamp_min_i32_kernel(int* __restrict__ x, int lo, int n) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = max(x[i], lo);
}
__global__ void fill_i32_kernel(int* __restrict__ x, int v, int n) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

// Backward of log_softmax over the vocabulary (Generator, TransformerModel.py:1315-1323):
//   dz[row, v] = g[row, v] - exp(logp[row, v]) * sum_v g[row, v],  written as T with row pitch ldz (pad columns = 0).
template <typename T>
__global__ void __launch_bounds__(512)
logsoftmax_bwd_kernel(const float* __restrict__ g, const float* __restrict__ logp, int V, T* __restrict__ dz, int ldz) {
  pdl_enter();
  __shared__ float s_sum[16];
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* gr = g + (size_t)row * V;
  const float* lr = logp + (size_t)row * V;
  float s = 0.f;
  for (int c = tid; c < V; c += 512) s += gr[c];
  s = warp_sum(s);
  if (lane == 0) s_sum[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) tot += s_sum[w];
  T* o = dz + (size_t)row * ldz;
  for (int c = tid; c < ldz; c += 512) o[c] = from_float<T>(c < V ? gr[c] - expf(lr[c]) * tot : 0.f);
}

// Fused criterion gradient for the word log-probs (LanguageModelCriterion_UIC, losses.py:329-331 + Generator):
// loss = -sum_rows mask[row] * logp[row, label[row]] / denom  =>  dz = (softmax - onehot) * mask / denom.
// Never materialises a [rows, V] gradient tensor.  Also accumulates the loss numerator per row.
template <typename T>
__global__ void __launch_bounds__(512)
xe_word_loss_bwd_kernel(const float* __restrict__ logits, int ldl, int V, const int* __restrict__ labels, int lab_stride, int lab_off,
                        const int* __restrict__ n_words, int Tt, const int* __restrict__ total_words, T* __restrict__ dz, int ldz,
                        float* __restrict__ row_nll) {
  pdl_enter();
  __shared__ float s_red[16];
  const float inv_denom = 1.0f / (float)(*total_words);
  const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = row / Tt, t = row - n * Tt;
  const bool live = t < n_words[n];
  T* o = dz + (size_t)row * ldz;
  if (!live) {
    for (int c = tid; c < ldz; c += 512) o[c] = from_float<T>(0.f);
    if (tid == 0) row_nll[row] = 0.f;
    return;
  }
  const float* z = logits + (size_t)row * ldl;
  float mx = -INFINITY;
  for (int c = tid; c < V; c += 512) mx = fmaxf(mx, z[c]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = s_red[0];
#pragma unroll
  for (int w = 1; w < 16; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  float s = 0.f;
  for (int c = tid; c < V; c += 512) s += expf(z[c] - mx);
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 16; ++w) tot += s_red[w];
  const float lse = logf(tot);
  const int lab = labels[(size_t)n * lab_stride + lab_off + t];
  for (int c = tid; c < ldz; c += 512) {
    float v = 0.f;
    if (c < V) v = (expf((z[c] - mx) - lse) - (c == lab ? 1.f : 0.f)) * inv_denom;
    o[c] = from_float<T>(v);
  }
  if (tid == 0) row_nll[row] = -((z[lab] - mx) - lse);
}

// out[c][r] = in[r][c] for r < rows, 0 for rows <= r < rows_pad   (in: [rows, cols] pitch ld_in; out: [cols, rows_pad])
template <typename TIn, typename TOut>
__global__ void transpose_pad_kernel(const TIn* __restrict__ in, int ld_in, TOut* __restrict__ out, int rows, int cols, int rows_pad) {
  pdl_enter();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? to_float<TIn>(in[(size_t)r * ld_in + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows_pad) out[(size_t)c * rows_pad + r] = from_float<TOut>(tile[threadIdx.x][i]);
  }
}

// Bias gradients gb[c] += sum_r dY[r, c], one launch, fixed summation order:
//   every CTA (256 columns x one row chunk; 8 row lanes, 8- / 16-byte loads of 4 columns per thread) writes its partial
//   sums; the LAST CTA of a column block to finish (ticket counter, self-resetting) adds the chunk partials in chunk order.
// `cols4` = columns rounded up to 4 (pad columns of dY are zero).
template <typename T>
__global__ void __launch_bounds__(512)
colsum_kernel(const T* __restrict__ dY, int ld, int rows, int cols, int cols4, float* __restrict__ partial, int* __restrict__ tickets,
              float* __restrict__ out) {
  pdl_enter();
  __shared__ float4 part[8][64];
  __shared__ int s_last;
  const int cg = threadIdx.x & 63, ry = threadIdx.x >> 6;
  const int c = (blockIdx.x * 64 + cg) * 4;
  const int per = (rows + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols4) {
    int r = r0 + ry;
    for (; r + 24 < r1; r += 32) {                     // four independent loads in flight
      const float4 v0 = load4(dY + (size_t)r * ld + c), v1 = load4(dY + (size_t)(r + 8) * ld + c);
      const float4 v2 = load4(dY + (size_t)(r + 16) * ld + c), v3 = load4(dY + (size_t)(r + 24) * ld + c);
      acc.x += (v0.x + v1.x) + (v2.x + v3.x); acc.y += (v0.y + v1.y) + (v2.y + v3.y);
      acc.z += (v0.z + v1.z) + (v2.z + v3.z); acc.w += (v0.w + v1.w) + (v2.w + v3.w);
    }
    for (; r < r1; r += 8) {
      const float4 v = load4(dY + (size_t)r * ld + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part[ry][cg] = acc;
  __syncthreads();
  if (ry == 0 && c < cols4) {
    float4 t = part[0][cg];
#pragma unroll
    for (int i = 1; i < 8; ++i) { t.x += part[i][cg].x; t.y += part[i][cg].y; t.z += part[i][cg].z; t.w += part[i][cg].w; }
    store4(partial + (size_t)blockIdx.y * cols4 + c, t);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&tickets[blockIdx.x], 1) == (int)gridDim.y - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (threadIdx.x < 256 && col < cols) {
    float t = 0.f;
#pragma unroll 8
    for (int i = 0; i < (int)gridDim.y; ++i) t += __ldcg(partial + (size_t)i * cols4 + col);
    out[col] += t;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0;
}

// ---------------------------------------------------------------------------------------------
// LayerNorm backward (forward: y = a * (x - mean) / (std_unbiased + eps) + b, TransformerModel.py:1338-1349).
// With c = x - mean, s = std, t = s + eps, g = dy * a:
//   dx = (g - mean(g)) / t - c * sum(g * c) / (t^2 * s * (D - 1)),   da = sum_rows dy * c / t,   db = sum_rows dy
// One warp per row (grid-stride); dx_out = dres + dx is written as fp32 (residual-stream gradient) and as T (operand
// of the next GEMMs).  Parameter gradients: per-CTA partial sums -> ln_param_reduce_kernel (fixed order).
// ---------------------------------------------------------------------------------------------
template <typename TG, typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ a2, const TG* __restrict__ dy, const float* __restrict__ dres,
                     float* __restrict__ dx_out, T* __restrict__ dx_out_t, int rows, float* __restrict__ partial) {
  pdl_enter();
  __shared__ float s_da[8][kD], s_db[8][kD];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float da[16], db[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) da[i] = db[i] = 0.f;
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const float* xr = x + (size_t)row * kD;
    float4 v[4], g[4];
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
    const float sd = sqrtf(warp_sum(ss) * (1.0f / (kD - 1)));
    const float t = sd + 1e-6f, inv_t = 1.0f / t;
    float sg = 0.f, sgc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 d = load4(dy + (size_t)row * kD + c), a = load4(a2 + c);
      da[4 * i + 0] += d.x * v[i].x * inv_t; da[4 * i + 1] += d.y * v[i].y * inv_t;
      da[4 * i + 2] += d.z * v[i].z * inv_t; da[4 * i + 3] += d.w * v[i].w * inv_t;
      db[4 * i + 0] += d.x; db[4 * i + 1] += d.y; db[4 * i + 2] += d.z; db[4 * i + 3] += d.w;
      g[i] = make_float4(d.x * a.x, d.y * a.y, d.z * a.z, d.w * a.w);
      sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
      sgc += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
    }
    const float mg = warp_sum(sg) * (1.0f / kD);
    const float k2 = (sd > 0.f) ? warp_sum(sgc) / (t * t * sd * (kD - 1)) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o;
      o.x = (g[i].x - mg) * inv_t - v[i].x * k2;
      o.y = (g[i].y - mg) * inv_t - v[i].y * k2;
      o.z = (g[i].z - mg) * inv_t - v[i].z * k2;
      o.w = (g[i].w - mg) * inv_t - v[i].w * k2;
      if (dres) {
        const float4 r = load4(dres + (size_t)row * kD + c);
        o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
      }
      if (dx_out) store4(dx_out + (size_t)row * kD + c, o);
      if (dx_out_t) store4(dx_out_t + (size_t)row * kD + c, o);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) { s_da[warp][c + q] = da[4 * i + q]; s_db[warp][c + q] = db[4 * i + q]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kD; c += 256) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { ta += s_da[w][c]; tb += s_db[w][c]; }
    partial[(size_t)blockIdx.x * 2 * kD + c] = ta;
    partial[(size_t)blockIdx.x * 2 * kD + kD + c] = tb;
  }
}

__global__ void __launch_bounds__(256)
ln_param_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ ga, float* __restrict__ gb) {
  pdl_enter();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x, py = threadIdx.y;       // c in [0, 2*kD)
  float t = 0.f;
#pragma unroll 4
  for (int p = py; p < nparts; p += 8) t += partial[(size_t)p * 2 * kD + c];
  red[py][threadIdx.x] = t;
  __syncthreads();
  if (py == 0) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][threadIdx.x];
    if (c < kD) ga[c] += v; else gb[c - kD] += v;
  }
}

// ---------------------------------------------------------------------------------------------
// Attention backward for the short sequences of this model (prefix masks in visible-count form, as the forward).
// One CTA per (head, key/value block).  A K/V block is one sequence (self-attention) or one image (cross-attention);
// `qpk * Tq` consecutive query rows attend to it (qpk sequences per block).  Queries are processed in chunks of 32:
//   phase A (warp per query): s = q.K^T*scale, p = softmax(s[:nvis]), dp = dO.V^T, ds = p*(dp - sum(p*dp)); dq = ds.K*scale
//   phase B (thread per (key, dim)): dK += ds^T.q*scale, dV += p^T.dO      -- accumulated in registers, no atomics
// fp32 FFMA throughout (the problem is a few hundred kFLOP per CTA); dQ/dK/dV are written as T.
// ---------------------------------------------------------------------------------------------
constexpr int kBwdChunk = 32;
template <typename T, int KPT>      // KPT = keys per thread in phase B (4 key lanes): Tk <= 4*KPT
__global__ void __launch_bounds__(256)
attention_bwd_kernel(const T* __restrict__ Q, int ldq, const T* __restrict__ K, const T* __restrict__ V, int ldkv,
                     const T* __restrict__ dO, int ldo, T* __restrict__ dQ, int lddq, T* __restrict__ dK, T* __restrict__ dV, int lddkv,
                     int Tq, int Tk, int qpk, const int* __restrict__ vis, int vis_bs, int vis_qs, int vis_div, float scale,
                     int accumulate_kv, Drop drop) {
  pdl_enter();
  extern __shared__ float bsm[];
  const int TKP = (Tk + 3) & ~3;
  float* Ks = bsm;                                  // [Tk][65]
  float* Vs = Ks + Tk * 65;                         // [Tk][65]
  float* Qs = Vs + Tk * 65;                         // [32][64]
  float* Os = Qs + kBwdChunk * 64;                  // [32][64]  dO
  float* Ps = Os + kBwdChunk * 64;                  // [32][TKP]
  float* Ds = Ps + kBwdChunk * TKP;                 // [32][TKP] dS
  const int head = blockIdx.x, kb = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < Tk * 16; idx += 256) {
    const int j = idx >> 4, c = (idx & 15) * 4;
    const float4 kq = load4(K + ((size_t)kb * Tk + j) * ldkv + head * kHeadDim + c);
    const float4 vq = load4(V + ((size_t)kb * Tk + j) * ldkv + head * kHeadDim + c);
    float* kd = Ks + j * 65 + c;
    float* vd = Vs + j * 65 + c;
    kd[0] = kq.x; kd[1] = kq.y; kd[2] = kq.z; kd[3] = kq.w;
    vd[0] = vq.x; vd[1] = vq.y; vd[2] = vq.z; vd[3] = vq.w;
  }
  float dk[KPT], dv[KPT];
#pragma unroll
  for (int i = 0; i < KPT; ++i) dk[i] = dv[i] = 0.f;
  const int nq = qpk * Tq;
  const int d_b = tid & 63, jg = tid >> 6;
  for (int q0 = 0; q0 < nq; q0 += kBwdChunk) {
    const int qc = min(kBwdChunk, nq - q0);
    __syncthreads();                                 // previous chunk's phase B done (and K/V staged on the first trip)
    for (int idx = tid; idx < qc * 16; idx += 256) {
      const int i = idx >> 4, c = (idx & 15) * 4;
      const size_t r = (size_t)kb * nq + q0 + i;
      *reinterpret_cast<float4*>(Qs + i * 64 + c) = load4(Q + r * ldq + head * kHeadDim + c);
      *reinterpret_cast<float4*>(Os + i * 64 + c) = load4(dO + r * ldo + head * kHeadDim + c);
    }
    __syncthreads();
    // ---- phase A
    for (int i = warp; i < qc; i += 8) {
      const int qi = q0 + i, bq = kb * qpk + qi / Tq, t = qi % Tq;
      int nvis = vis ? vis[(size_t)(bq / vis_div) * vis_bs + (size_t)t * vis_qs] : Tk;
      nvis = min(nvis, Tk);
      const float* q = Qs + i * 64;
      const float* go = Os + i * 64;
      float sc[4], dp[4];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = jj * 32 + lane;
        sc[jj] = -INFINITY;
        dp[jj] = 0.f;
        if (j < nvis) {
          const float* kr = Ks + j * 65;
          const float* vr = Vs + j * 65;
          float d = 0.f, e = 0.f;
#pragma unroll 16
          for (int c = 0; c < kHeadDim; ++c) { d = fmaf(q[c], kr[c], d); e = fmaf(go[c], vr[c], e); }
          sc[jj] = d * scale;
          dp[jj] = e;
        }
        mx = fmaxf(mx, sc[jj]);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = jj * 32 + lane;
        sc[jj] = (j < nvis) ? expf(sc[jj] - mx) : 0.f;
        sum += sc[jj];
      }
      sum = warp_sum(sum);
      float dsum = 0.f, km[4];
      const int qrow = kb * nq + qi;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        km[jj] = drop_mul(drop, att_idx(qrow, head, jj * 32 + lane));       // dropout on the probabilities
        sc[jj] = sc[jj] / sum;
        dp[jj] *= km[jj];
        dsum += sc[jj] * dp[jj];
      }
      dsum = warp_sum(dsum);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = jj * 32 + lane;
        if (j < TKP) {
          Ps[i * TKP + j] = (j < nvis) ? sc[jj] * km[jj] : 0.f;
          Ds[i * TKP + j] = (j < nvis) ? sc[jj] * (dp[jj] - dsum) : 0.f;
        }
      }
      __syncwarp();
      float q0v = 0.f, q1v = 0.f;
      for (int j = 0; j < nvis; ++j) {
        const float ds = Ds[i * TKP + j];
        q0v = fmaf(ds, Ks[j * 65 + lane], q0v);
        q1v = fmaf(ds, Ks[j * 65 + lane + 32], q1v);
      }
      T* dq = dQ + ((size_t)kb * nq + qi) * lddq + head * kHeadDim;
      dq[lane] = from_float<T>(q0v * scale);
      dq[lane + 32] = from_float<T>(q1v * scale);
    }
    __syncthreads();
    // ---- phase B
#pragma unroll
    for (int u = 0; u < KPT; ++u) {
      const int j = jg + 4 * u;
      if (j < Tk) {
        float ak = 0.f, av = 0.f;
        for (int i = 0; i < qc; ++i) {
          ak = fmaf(Ds[i * TKP + j], Qs[i * 64 + d_b], ak);
          av = fmaf(Ps[i * TKP + j], Os[i * 64 + d_b], av);
        }
        dk[u] += ak;
        dv[u] += av;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < KPT; ++u) {
    const int j = jg + 4 * u;
    if (j < Tk) {
      T* pk = dK + ((size_t)kb * Tk + j) * lddkv + head * kHeadDim + d_b;
      T* pv = dV + ((size_t)kb * Tk + j) * lddkv + head * kHeadDim + d_b;
      float vk = dk[u] * scale, vv = dv[u];
      if (accumulate_kv) { vk += to_float<T>(*pk); vv += to_float<T>(*pv); }
      *pk = from_float<T>(vk);
      *pv = from_float<T>(vv);
    }
  }
}
inline size_t attention_bwd_smem_bytes(int Tk) {
  const int TKP = (Tk + 3) & ~3;
  return sizeof(float) * ((size_t)2 * Tk * 65 + 2 * kBwdChunk * 64 + 2 * kBwdChunk * TKP);
}

// d_ffh *= (ffh > 0)   (ReLU backward; ffh is the saved post-ReLU activation)
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(T* __restrict__ d, const T* __restrict__ act, size_t n4, float gain) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 g = load4(d + i * 4);
    const float4 a = load4(act + i * 4);
    // `act` is the saved post-ReLU (and post-dropout) activation: zero where either killed it; gain = 1/(1-p) of the dropout
    g.x = a.x > 0.f ? g.x * gain : 0.f; g.y = a.y > 0.f ? g.y * gain : 0.f; g.z = a.z > 0.f ? g.z * gain : 0.f; g.w = a.w > 0.f ? g.w * gain : 0.f;
    store4(d + i * 4, g);
  }
}
// out(T) = dx * (x0 > 0)   (att_embed ReLU backward, fp32 residual gradient in, GEMM operand out)
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_cast_kernel(const float* __restrict__ dx, const float* __restrict__ act, T* __restrict__ out, size_t n4, float gain) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 g = load4(dx + i * 4);
    const float4 a = load4(act + i * 4);
    g.x = a.x > 0.f ? g.x * gain : 0.f; g.y = a.y > 0.f ? g.y * gain : 0.f; g.z = a.z > 0.f ? g.z * gain : 0.f; g.w = a.w > 0.f ? g.w * gain : 0.f;
    store4(out + i * 4, g);
  }
}

// a += b (fp32, n % 4 == 0)
__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ a, const float* __restrict__ b, size_t n4) {
  pdl_enter();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 x = load4(a + i * 4);
    const float4 y = load4(b + i * 4);
    x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w;
    store4(a + i * 4, x);
  }
}

// out[n*out_stride : +512] (+)= sum_p in[(n*P + p), :]    (gradient of a row that was repeated over the passes)
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
sum_over_passes_kernel(const TIn* __restrict__ in, int P, TOut* __restrict__ out, size_t out_stride, int N, int accumulate) {
  pdl_enter();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    float4 acc = accumulate ? load4(out + (size_t)n * out_stride + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < P; ++p) {
      const float4 v = load4(in + ((size_t)n * P + p) * kD + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    store4(out + (size_t)n * out_stride + c, acc);
  }
}

// Embedding backward for a big table (tgt_embed): g_lut[id, :] += sqrt(d) * dx[row, :] with fp32 atomics; rows whose
// gradient is exactly zero (padding slots: no loss, never a visible key) are skipped.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ ids, int ids_stride, int ids_off, int T, int rows,
                 float sqrt_d, float* __restrict__ g_lut) {
  pdl_enter();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / T, r = row - b * T;
  const int id = ids[(size_t)b * ids_stride + ids_off + r];
  float4 v[4];
  bool nz = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = load4(dx + (size_t)row * kD + (i * 32 + lane) * 4);
    nz |= (v[i].x != 0.f) | (v[i].y != 0.f) | (v[i].z != 0.f) | (v[i].w != 0.f);
  }
  if (!__any_sync(0xffffffffu, nz)) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* g = g_lut + (size_t)id * kD + (i * 32 + lane) * 4;
    atomicAdd(g + 0, v[i].x * sqrt_d); atomicAdd(g + 1, v[i].y * sqrt_d);
    atomicAdd(g + 2, v[i].z * sqrt_d); atomicAdd(g + 3, v[i].w * sqrt_d);
  }
}

// Embedding backward for a tiny id space (syn_embed: 10 ids; constant word: 1 id): deterministic, two stages.
//   stage 1: CTA = 64 columns x one row chunk, 4 row lanes accumulate per-id sums in shared memory
//            -> partial[chunk][id][512]
//   stage 2: g_lut[id, c] += sqrt(d) * sum over chunks.      ids == nullptr: every row has id `const_id` (slot 0).
constexpr int kSmallIds = 10;
__global__ void __launch_bounds__(256)
embed_small_partial_kernel(const float* __restrict__ dx, const int* __restrict__ ids, int ids_stride, int ids_off, int T, int rows,
                           float* __restrict__ partial) {
  pdl_enter();
  __shared__ float acc[4][kSmallIds][64];
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6, col = blockIdx.x * 64 + c;
  for (int i = 0; i < kSmallIds; ++i) acc[rl][i][c] = 0.f;
  const int per = (rows + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  for (int row = r0 + rl; row < r1; row += 4) {
    int slot = 0;
    if (ids) {
      const int b = row / T, r = row - b * T;
      slot = ids[(size_t)b * ids_stride + ids_off + r];
    }
    if (slot >= 0 && slot < kSmallIds) acc[rl][slot][c] += dx[(size_t)row * kD + col];
  }
  __syncthreads();
  if (rl == 0) {
    for (int i = 0; i < kSmallIds; ++i)
      partial[((size_t)blockIdx.y * kSmallIds + i) * kD + col] = ((acc[0][i][c] + acc[1][i][c]) + acc[2][i][c]) + acc[3][i][c];
  }
}
__global__ void __launch_bounds__(256)
embed_small_reduce_kernel(const float* __restrict__ partial, int chunks, int n_ids, int const_id, float sqrt_d, float* __restrict__ g_lut) {
  pdl_enter();
  const int i = blockIdx.x * 256 + threadIdx.x;                        // (id slot, column)
  if (i >= n_ids * kD) return;
  const int slot = i / kD, col = i - slot * kD;
  float t = 0.f;
#pragma unroll 4
  for (int ch = 0; ch < chunks; ++ch) t += partial[((size_t)ch * kSmallIds + slot) * kD + col];
  const int id = const_id >= 0 ? const_id : slot;
  g_lut[(size_t)id * kD + col] += t * sqrt_d;
}

// Criterion for the bounding heads fused with its gradient (losses.py:340-349): for row (n, p), p < phrase_num[n]:
//   loss += -logp[n, p, target[n, p+1]] / denom ;  g[n, p, target] = -1/denom, everything else 0.
__global__ void xe_box_loss_grad_kernel(const float* __restrict__ len_logp, const float* __restrict__ syn_logp,
                                        const int* __restrict__ phrase_num, const int* __restrict__ phrase_length,
                                        const int* __restrict__ phrase_syn, int N, int Tb, int n_len, int n_syn,
                                        const int* __restrict__ total_words, float* __restrict__ g_len, float* __restrict__ g_syn,
                                        float* __restrict__ row_nll_len, float* __restrict__ row_nll_syn) {
  pdl_enter();
  const float inv_denom = 1.0f / (float)(*total_words);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int slots = Tb - 1;
  if (i >= N * slots) return;
  const int n = i / slots, p = i - n * slots;
  float nl = 0.f, ns = 0.f;
  for (int o = 0; o < n_len; ++o) g_len[(size_t)i * n_len + o] = 0.f;
  for (int o = 0; o < n_syn; ++o) g_syn[(size_t)i * n_syn + o] = 0.f;
  if (p < phrase_num[n]) {
    const int tl = phrase_length[n * Tb + p + 1], ts = phrase_syn[n * Tb + p + 1];
    nl = -len_logp[(size_t)i * n_len + tl];
    ns = -syn_logp[(size_t)i * n_syn + ts];
    g_len[(size_t)i * n_len + tl] = -inv_denom;
    g_syn[(size_t)i * n_syn + ts] = -inv_denom;
  }
  row_nll_len[i] = nl;
  row_nll_syn[i] = ns;
}

// out[0] = sum(v[0..n)) in a fixed order (single CTA; n is a few tens of thousands)
__global__ void __launch_bounds__(1024) sum_kernel(const float* __restrict__ v, int n, float* __restrict__ out, float scale) {
  pdl_enter();
  __shared__ float s[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) acc += v[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_sum(s[threadIdx.x]);
    if (threadIdx.x == 0) out[0] = t * scale;
  }
}

// losses[1..6] = sums[0..5] / total_words, losses[0] = their sum
// (reference order: total, SA_length, SA_phrase, SA_syn, NA_length, NA_phrase, NA_syn; losses.py:359-369)
__global__ void xe_finish_losses_kernel(const float* __restrict__ sums, const int* __restrict__ total_words, float* __restrict__ losses) {
  pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float inv = 1.0f / (float)(*total_words);
  float tot = 0.f;
  for (int i = 0; i < 6; ++i) { losses[1 + i] = sums[i] * inv; tot += sums[i] * inv; }
  losses[0] = tot;
}

// n_words[n] = sum(phrase_length[n, :]) - 1   (word slots that carry loss, losses.py:329-330)
__global__ void xe_count_words_kernel(const int* __restrict__ phrase_length, int N, int Tb, int* __restrict__ n_words, int* __restrict__ total) {
  pdl_enter();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int s = 0;
  for (int t = 0; t < Tb; ++t) s += phrase_length[n * Tb + t];
  n_words[n] = s - 1;
  atomicAdd(total, s - 1);
}

}  // namespace bofi
