// XE-training path, included by engine.cu: teacher-forced forward of
// EncoderDecoder_UIC.forward + Generator (TransformerModel.py:413-468 with glat_p < 0, :1713-1775 train_mode UIC),
// the criterion LanguageModelCriterion_UIC (losses.py:315-369) and the backward pass of both.
//
// Formulation (N_len == 1, the uic_sd.yml configuration):
//   * the encoder runs once per IMAGE; the reference repeats the features seq_per_img times first (same values);
//   * memory K/V of every decoder-style layer are projected once and shared by the SA / NA passes and by the
//     seq_per_img captions of an image (K/V block = image);
//   * the bounding head only reads its [LEN] row (:375) and its layer input is the same at every bounding pass
//     (teacher forcing), so LN + QKV of the Tb input rows are computed once and all P = max(phrase_num) passes are
//     ONE batch of N*P single-query rows (row order (n, p)) through self-attention / cross-attention / FFN / heads.
// Every activation needed by the backward pass is kept in a bump arena (no recomputation except LayerNorm
// statistics): about 6 GB at 256 images x 5 captions, which is what the 180 GB of HBM are for.
// Dropout: the reference's train() mode applies dropout; parity is defined with dropout off (eval()), which is how
// the golden gradients were recorded (oracle/make_golden_xe.py).  See DESIGN.md.

struct Arena {
  std::vector<DevBuf> chunks;
  size_t cur = 0, off = 0, used_total = 0;
  void reset() {
    if (chunks.size() > 1) {          // coalesce: next step fits in one chunk
      size_t total = 0;
      for (DevBuf& c : chunks) { total += c.cap; c.release(); }
      chunks.clear();
      chunks.emplace_back();
      chunks[0].reserve(total + (total >> 2));
    }
    cur = 0;
    off = 0;
    used_total = 0;
  }
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (chunks.empty()) chunks.emplace_back();
    while (true) {
      if (cur < chunks.size() && off + bytes <= chunks[cur].cap) {
        void* p = (char*)chunks[cur].p + off;
        off += bytes;
        used_total += bytes;
        return p;
      }
      if (cur + 1 < chunks.size()) { ++cur; off = 0; continue; }
      chunks.emplace_back();
      cur = chunks.size() - 1;
      off = 0;
      if (chunks[cur].reserve(std::max<size_t>(bytes, (size_t)512 << 20)) != BOFI_OK) return nullptr;
    }
  }
  void release() {
    for (DevBuf& c : chunks) c.release();
    chunks.clear();
    cur = off = 0;
  }
  // scratch passes (the no-grad glancing pass): everything allocated after mark() is handed back by rewind()
  struct Mark { size_t cur, off, used_total; };
  Mark mark() const { return Mark{cur, off, used_total}; }
  void rewind(const Mark& m) { cur = m.cur; off = m.off; used_total = m.used_total; }
};

struct LayerTape {        // T-typed buffers are void*
  const float* x_in = nullptr;
  void *y0 = nullptr, *qkv = nullptr, *ao = nullptr;
  float* x1 = nullptr;
  void *y1 = nullptr, *q = nullptr, *ao2 = nullptr;
  float* x2 = nullptr;
  void *y2 = nullptr, *ffh = nullptr;
  float* x_out = nullptr;
  Drop d_sa_att, d_sa_out, d_ca_att, d_ca_out, d_ffh, d_ffn_out;     // dropout sites of the layer (disabled when p == 0)
};
struct BoundTape {
  float* x_in = nullptr;          // [N*Tb, 512]
  void *y0 = nullptr, *qkv = nullptr;   // [N*Tb, .]
  void *q_rep = nullptr;          // [Mb, 512] the [LEN] query repeated per pass (backward only needs it)
  LayerTape lt;                   // rows (n, p): ao, x1, y1, q, ao2, x2, y2, ffh, x_out
  std::vector<LayerTape> full;    // N_len >= 2: [P][N_len] full-layer tapes over the N*Tb rows of every pass (t_bound_fwd_full)
  float* hn = nullptr;            // [Mb, 512] length_predictor.norm output (fp32: the heads are fp32)
  float* hid = nullptr;           // [Mb, 200]
  Drop d_embed, d_hid;
};
struct DecTape {
  float* x_in = nullptr;
  std::vector<LayerTape> layers;
  void* y_final = nullptr;        // [N*T, 512]
  float* logits = nullptr;        // [N*T, Vpad] (fused-loss path) or nullptr
  Drop d_embed;
};
struct TrainState {
  bool valid = false;
  int B = 0, R = 0, spi = 0, N = 0, T = 0, Tb = 0, P = 0, Mb = 0;
  bool have_len = false;
  Arena arena;
  int *labels = nullptr, *pnum = nullptr, *plen = nullptr, *psyn = nullptr, *ext_syn = nullptr, *ext_seq = nullptr, *sa_vis = nullptr;
  int *word_seq = nullptr, *vis_b = nullptr, *na_vis = nullptr, *n_words = nullptr, *total_words = nullptr;
  int* vis_full = nullptr;        // N_len >= 2: [P][N][Tb] visible keys of every row at every bounding pass
  void* attT = nullptr;           // [B*R, F] GEMM operand copy of the features
  float* x0 = nullptr;            // att_embed output
  std::vector<LayerTape> enc;
  const float* enc_x_final = nullptr;
  void* memT = nullptr;           // encoder output (GEMM operand type)
  int* att_len = nullptr;         // valid regions per image (nullptr: all)
  std::vector<void*> kv;          // [B*R, 1024] per bounding / decoder layer
  BoundTape sa_b, na_b;
  DecTape sa_d, na_d;
  DevBuf tr_a, tr_b, tr_w, zeros, ln_partial, cs_partial, cs_tickets, scratch_f32, dkv, dmem, zbuf, dlen;
  // Dropout (train() mode of the reference): p_sub = opt.dropout at every sub-layer output, attention probability,
  // FFN hidden, positional encoding and head hidden; p_att = opt.drop_prob_lm after att_embed's ReLU.  Sites are numbered
  // in forward order from 0 every step; key(site) = drop_hash(seed, site); the backward pass reuses the keys.
  float p_sub = 0.f, p_att = 0.f;
  // self-critical sampling (bofi_sc_sample): decoder inputs of the taped pass, copied out of the decode state
  bool sc = false;
  int sc_mode = 0;
  int *sc_words = nullptr, *sc_vis = nullptr, *sc_total = nullptr;     // [N, T] word ids (SAIC) / visible keys; [N] committed words + 1
  // glancing training (EncoderDecoder_UIC.forward with glat_p >= 0, :437-464): the NA decoder's word inputs
  float glat_p = -1.f;
  uint32_t glat_seed = 0;
  int* glat_words = nullptr;      // [N, T] bos or the ground-truth word per slot (nullptr: glancing off, constant bos)
  cudaEvent_t grad_event = nullptr;   // bofi_train_set_grad_event: recorded once every gradient outside the encoder is final
  std::vector<cudaEvent_t> layer_events;   // bofi_train_set_layer_event: [l] recorded once encoder layer l's gradients are final
  uint32_t seed = 0, site = 0;
  Drop d_att_embed;
  Drop next_drop(float p) {
    Drop d;
    if (p > 0.f) {
      d.key = drop_hash(seed, site);
      d.thresh = (uint32_t)std::min(4294967295.0, (double)p * 4294967296.0);
      d.scale = 1.0f / (1.0f - p);
    }
    ++site;
    return d;
  }
};

template <typename T>
static int sublayer_out(bofi_engine* e, cudaStream_t s, TrainState* ts, const T* a, int lda, const Lin& lin, const float* x_in, float* x_out,
                        int rows, const Drop& d);

static TrainState* train_state(bofi_engine* e);

template <typename T> static T* aalloc(TrainState* ts, size_t n) { return reinterpret_cast<T*>(ts->arena.alloc(n * sizeof(T))); }
#define A_TRY(ptr) do { if (!(ptr)) return fail(BOFI_ERR_NOMEM, "training arena allocation failed"); } while (0)

// ---- dropout helpers ------------------------------------------------------------------------------------------
template <typename T>
static int dropout_inplace(bofi_engine* e, cudaStream_t s, T* x, size_t n, const Drop& d) {
  if (!d.thresh) return BOFI_OK;
  ProfScope prof(e, s, PC_OTHER, 0.0, 2.0 * n * sizeof(T));
  launch_k(dropout_inplace_kernel<T>, 148 * 8, 256, 0, s, x, n / 4, d);
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}
template <typename TIn, typename TOut>
static int dropout_apply(bofi_engine* e, cudaStream_t s, const TIn* in, TOut* out, size_t n, const Drop& d) {
  ProfScope prof(e, s, PC_OTHER, 0.0, (double)n * (sizeof(TIn) + sizeof(TOut)));
  launch_k(dropout_apply_kernel<TIn, TOut>, 148 * 8, 256, 0, s, in, out, n / 4, d);
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}
// SublayerConnection output: x_out = x_in + dropout(a . W^T + b).  Without dropout the residual add is the GEMM epilogue;
// with dropout the GEMM writes the sub-layer output z and one fused pass applies mask, scale and residual.
template <typename T>
static int sublayer_out(bofi_engine* e, cudaStream_t s, TrainState* ts, const T* a, int lda, const Lin& lin, const float* x_in, float* x_out,
                        int rows, const Drop& d) {
  if (!d.thresh) return linear<T, float>(e, s, a, lda, lin, x_in, kD, x_out, kD, rows, 0, nullptr);
  RC_TRY(ts->zbuf.reserve((size_t)rows * kD * sizeof(T)));
  T* z = ts->zbuf.as<T>();
  RC_TRY((linear<T, T>(e, s, a, lda, lin, nullptr, 0, z, kD, rows, 0, nullptr)));
  ProfScope prof(e, s, PC_OTHER, 0.0, (double)rows * kD * (sizeof(T) + 8));
  launch_k(drop_add_kernel<T>, 148 * 8, 256, 0, s, (const T*)z, x_in, x_out, (size_t)rows * kD / 4, d);
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}
// gradient w.r.t. the sub-layer output z: dz = dxT o mask / (1 - p)  (dxT itself when dropout is off)
template <typename T>
static int sublayer_grad(bofi_engine* e, cudaStream_t s, TrainState* ts, const T* dxT, int rows, const Drop& d, const T** dz) {
  if (!d.thresh) { *dz = dxT; return BOFI_OK; }
  RC_TRY(ts->zbuf.reserve((size_t)rows * kD * sizeof(T)));
  RC_TRY((dropout_apply<T, T>(e, s, dxT, ts->zbuf.as<T>(), (size_t)rows * kD, d)));
  *dz = ts->zbuf.as<T>();
  return BOFI_OK;
}

// ---- forward pieces -----------------------------------------------------------------------------------------
// One pre-norm layer with every intermediate kept (same kernels as run_layer).
template <typename T>
static int t_layer_fwd(bofi_engine* e, cudaStream_t s, TrainState* ts, const Layer& ly, LayerTape& tp, const float* x_in, int nb, int Tq,
                       const int* self_vis, int vis_bs, int vis_qs, const T* kvmem, int R, const int* mem_len, int kv_div) {
  const int rows = nb * Tq, dff = e->cfg.d_ff;
  tp.x_in = x_in;
  T* y0 = aalloc<T>(ts, (size_t)rows * kD); A_TRY(y0);
  T* qkv = aalloc<T>(ts, (size_t)rows * 3 * kD); A_TRY(qkv);
  T* ao = aalloc<T>(ts, (size_t)rows * kD); A_TRY(ao);
  float* x1 = aalloc<float>(ts, (size_t)rows * kD); A_TRY(x1);
  tp.y0 = y0; tp.qkv = qkv; tp.ao = ao; tp.x1 = x1;
  RC_TRY(layernorm<T>(e, s, x_in, kD, ly.ln[0], y0, kD, rows, nullptr, nullptr));
  RC_TRY((linear<T, T>(e, s, y0, kD, ly.sa.qkv, nullptr, 0, qkv, 3 * kD, rows, 0, nullptr)));
  tp.d_sa_att = ts->next_drop(ts->p_sub);
  tp.d_sa_out = ts->next_drop(ts->p_sub);
  RC_TRY(attention<T>(e, s, qkv, 3 * kD, qkv + kD, qkv + 2 * kD, 3 * kD, ao, kD, nb, Tq, Tq, self_vis, vis_bs, vis_qs, 1, 1, nullptr, nullptr,
                      tp.d_sa_att));
  RC_TRY(sublayer_out<T>(e, s, ts, ao, kD, ly.sa.o, x_in, x1, rows, tp.d_sa_out));
  const float* xm = x1;
  int f = 1;
  if (ly.cross) {
    T* y1 = aalloc<T>(ts, (size_t)rows * kD); A_TRY(y1);
    T* q = aalloc<T>(ts, (size_t)rows * kD); A_TRY(q);
    T* ao2 = aalloc<T>(ts, (size_t)rows * kD); A_TRY(ao2);
    float* x2 = aalloc<float>(ts, (size_t)rows * kD); A_TRY(x2);
    tp.y1 = y1; tp.q = q; tp.ao2 = ao2; tp.x2 = x2;
    RC_TRY(layernorm<T>(e, s, x1, kD, ly.ln[1], y1, kD, rows, nullptr, nullptr));
    RC_TRY((linear<T, T>(e, s, y1, kD, ly.ca.q, nullptr, 0, q, kD, rows, 0, nullptr)));
    tp.d_ca_att = ts->next_drop(ts->p_sub);
    tp.d_ca_out = ts->next_drop(ts->p_sub);
    RC_TRY(attention<T>(e, s, q, kD, kvmem, kvmem + kD, 2 * kD, ao2, kD, nb, Tq, R, mem_len, 1, 0, kv_div, kv_div, nullptr, nullptr, tp.d_ca_att));
    RC_TRY(sublayer_out<T>(e, s, ts, ao2, kD, ly.ca.o, x1, x2, rows, tp.d_ca_out));
    xm = x2;
    f = 2;
  }
  T* y2 = aalloc<T>(ts, (size_t)rows * kD); A_TRY(y2);
  T* ffh = aalloc<T>(ts, (size_t)rows * dff); A_TRY(ffh);
  float* xo = aalloc<float>(ts, (size_t)rows * kD); A_TRY(xo);
  tp.y2 = y2; tp.ffh = ffh; tp.x_out = xo;
  RC_TRY(layernorm<T>(e, s, xm, kD, ly.ln[f], y2, kD, rows, nullptr, nullptr));
  tp.d_ffh = ts->next_drop(ts->p_sub);
  tp.d_ffn_out = ts->next_drop(ts->p_sub);
  RC_TRY((linear<T, T>(e, s, y2, kD, ly.w1, nullptr, 0, ffh, dff, rows, 1, nullptr)));
  RC_TRY(dropout_inplace<T>(e, s, ffh, (size_t)rows * dff, tp.d_ffh));
  RC_TRY(sublayer_out<T>(e, s, ts, ffh, dff, ly.w2, xm, xo, rows, tp.d_ffn_out));
  return BOFI_OK;
}

// get_predict_phrase_length_syn_{SA,NA}: all P passes as one batch of (n, p) rows.
template <typename T>
static int t_bound_fwd(bofi_engine* e, cudaStream_t s, TrainState* ts, BoundTape& bt, const int* word_ids, const int* syn_ids,
                       float* len_logp, float* syn_logp) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, Tb = ts->Tb, P = ts->P, Mb = ts->Mb, dff = c.d_ff;
  const Layer& ly = e->lp[0];
  const int* mem_len = ts->att_len;
  const float scale = 1.0f / sqrtf((float)kHeadDim);
  bt.x_in = aalloc<float>(ts, (size_t)N * Tb * kD); A_TRY(bt.x_in);
  T* y0 = aalloc<T>(ts, (size_t)N * Tb * kD); A_TRY(y0);
  T* qkv = aalloc<T>(ts, (size_t)N * Tb * 3 * kD); A_TRY(qkv);
  bt.y0 = y0; bt.qkv = qkv;
  bt.d_embed = ts->next_drop(ts->p_sub);
  launch_k(embed_xe_kernel, ceil_div(N * Tb, 8), 256, 0, s, W(e, "model.tgt_embed.lut.weight"), W(e, "model.syn_embed.lut.weight"),
           W(e, "model.pos_embed.pe"), word_ids, Tb, 0, -1, syn_ids, Tb, 0, sqrtf((float)kD), bt.x_in, N * Tb, Tb, bt.d_embed);
  CU_TRY(cudaGetLastError());
  RC_TRY(layernorm<T>(e, s, bt.x_in, kD, ly.ln[0], y0, kD, N * Tb, nullptr, nullptr));
  RC_TRY((linear<T, T>(e, s, y0, kD, ly.sa.qkv, nullptr, 0, qkv, 3 * kD, N * Tb, 0, nullptr)));
  LayerTape& tp = bt.lt;
  T* ao = aalloc<T>(ts, (size_t)Mb * kD); A_TRY(ao);
  float* xr = aalloc<float>(ts, (size_t)Mb * kD); A_TRY(xr);       // [LEN] input row repeated per pass
  float* x1 = aalloc<float>(ts, (size_t)Mb * kD); A_TRY(x1);
  T* y1 = aalloc<T>(ts, (size_t)Mb * kD); A_TRY(y1);
  T* q = aalloc<T>(ts, (size_t)Mb * kD); A_TRY(q);
  T* ao2 = aalloc<T>(ts, (size_t)Mb * kD); A_TRY(ao2);
  float* x2 = aalloc<float>(ts, (size_t)Mb * kD); A_TRY(x2);
  T* y2 = aalloc<T>(ts, (size_t)Mb * kD); A_TRY(y2);
  T* ffh = aalloc<T>(ts, (size_t)Mb * dff); A_TRY(ffh);
  float* x3 = aalloc<float>(ts, (size_t)Mb * kD); A_TRY(x3);
  bt.hn = aalloc<float>(ts, (size_t)Mb * kD); A_TRY(bt.hn);
  bt.hid = aalloc<float>(ts, (size_t)Mb * 200); A_TRY(bt.hid);
  bt.q_rep = aalloc<T>(ts, (size_t)Mb * kD); A_TRY(bt.q_rep);
  tp.x_in = xr; tp.ao = ao; tp.x1 = x1; tp.y1 = y1; tp.q = q; tp.ao2 = ao2; tp.x2 = x2; tp.y2 = y2; tp.ffh = ffh; tp.x_out = x3;
  tp.d_sa_att = ts->next_drop(ts->p_sub);
  tp.d_sa_out = ts->next_drop(ts->p_sub);
  tp.d_ca_att = ts->next_drop(ts->p_sub);
  tp.d_ca_out = ts->next_drop(ts->p_sub);
  tp.d_ffh = ts->next_drop(ts->p_sub);
  tp.d_ffn_out = ts->next_drop(ts->p_sub);
  bt.d_hid = ts->next_drop(ts->p_sub);
  launch_k(xe_bound_self_attn_kernel<T>, Mb, 256, 0, s, (const T*)qkv, Tb, P, ts->vis_b, ao, scale, tp.d_sa_att);
  CU_TRY(cudaGetLastError());
  launch_k(repeat_row_kernel<float>, ceil_div(Mb, 8), 256, 0, s, (const float*)bt.x_in, (size_t)Tb * kD, xr, N, P);
  CU_TRY(cudaGetLastError());
  RC_TRY(sublayer_out<T>(e, s, ts, ao, kD, ly.sa.o, xr, x1, Mb, tp.d_sa_out));
  RC_TRY(layernorm<T>(e, s, x1, kD, ly.ln[1], y1, kD, Mb, nullptr, nullptr));
  RC_TRY((linear<T, T>(e, s, y1, kD, ly.ca.q, nullptr, 0, q, kD, Mb, 0, nullptr)));
  const T* kvm = (const T*)ts->kv[0];
  RC_TRY(attention<T>(e, s, q, kD, kvm, kvm + kD, 2 * kD, ao2, kD, Mb, 1, ts->R, mem_len, 1, 0, ts->spi * P, ts->spi * P, nullptr, nullptr,
                      tp.d_ca_att));
  RC_TRY(sublayer_out<T>(e, s, ts, ao2, kD, ly.ca.o, x1, x2, Mb, tp.d_ca_out));
  RC_TRY(layernorm<T>(e, s, x2, kD, ly.ln[2], y2, kD, Mb, nullptr, nullptr));
  RC_TRY((linear<T, T>(e, s, y2, kD, ly.w1, nullptr, 0, ffh, dff, Mb, 1, nullptr)));
  RC_TRY(dropout_inplace<T>(e, s, ffh, (size_t)Mb * dff, tp.d_ffh));
  RC_TRY(sublayer_out<T>(e, s, ts, ffh, dff, ly.w2, x2, x3, Mb, tp.d_ffn_out));
  RC_TRY(layernorm<float>(e, s, x3, kD, e->lp_norm, bt.hn, kD, Mb, nullptr, nullptr));
  {
    cudaError_t err = gemm_simt<float, float>(s, bt.hn, kD, e->head1.w32, kD, e->head1.b, nullptr, 0, bt.hid, 200, Mb, 200, kD, 1, nullptr);
    if (err != cudaSuccess) return fail(BOFI_ERR_CUDA, "head GEMM: %s", cudaGetErrorString(err));
    e->launches++;
  }
  RC_TRY(dropout_inplace<float>(e, s, bt.hid, (size_t)Mb * 200, bt.d_hid));
  launch_k(xe_head_logp_kernel, ceil_div(Mb, 4), 128, 0, s, (const float*)bt.hid, e->w_len2, e->b_len2, e->w_syn2, e->b_syn2, 100, 20, 10,
           len_logp, syn_logp, Mb, P, Tb - 1);
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

// N_len >= 2 (configs/uic_sd_N2.yml): the passes in the reference's own formulation -- every pass p runs the whole LengthPredictor stack
// over all Tb rows of every caption under that pass's mask (vis_full[p]), the [LEN] row of the last layer goes to the heads.  (The
// [LEN]-row shortcut of t_bound_fwd needs the other rows only as keys of ONE layer; with two layers the first layer's outputs of all
// rows are the second layer's keys.)  P * N_len full-layer tapes; the heads are those of t_bound_fwd.
template <typename T>
static int t_bound_fwd_full(bofi_engine* e, cudaStream_t s, TrainState* ts, BoundTape& bt, const int* word_ids, const int* syn_ids,
                            float* len_logp, float* syn_logp) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, Tb = ts->Tb, P = ts->P, Mb = ts->Mb, NL = c.n_len;
  const int* mem_len = ts->att_len;
  bt.x_in = aalloc<float>(ts, (size_t)N * Tb * kD); A_TRY(bt.x_in);
  bt.d_embed = ts->next_drop(ts->p_sub);
  launch_k(embed_xe_kernel, ceil_div(N * Tb, 8), 256, 0, s, W(e, "model.tgt_embed.lut.weight"), W(e, "model.syn_embed.lut.weight"),
           W(e, "model.pos_embed.pe"), word_ids, Tb, 0, -1, syn_ids, Tb, 0, sqrtf((float)kD), bt.x_in, N * Tb, Tb, bt.d_embed);
  CU_TRY(cudaGetLastError());
  float* x3 = aalloc<float>(ts, (size_t)Mb * kD); A_TRY(x3);
  bt.hn = aalloc<float>(ts, (size_t)Mb * kD); A_TRY(bt.hn);
  bt.hid = aalloc<float>(ts, (size_t)Mb * 200); A_TRY(bt.hid);
  bt.lt = LayerTape();
  bt.lt.x_out = x3;                                   // what length_predictor.norm normalises (the head backward reads it)
  bt.full.assign((size_t)P * NL, LayerTape());
  for (int p = 0; p < P; ++p) {
    const int* vis = ts->vis_full + (size_t)p * N * Tb;
    const float* x = bt.x_in;
    for (int l = 0; l < NL; ++l) {
      LayerTape& tp = bt.full[(size_t)p * NL + l];
      RC_TRY(t_layer_fwd<T>(e, s, ts, e->lp[l], tp, x, N, Tb, vis, Tb, 1, (const T*)ts->kv[l], ts->R, mem_len, ts->spi));
      x = tp.x_out;
    }
    launch_k(gather_len_rows_kernel, ceil_div(N * (kD / 4), 256), 256, 0, s, x, Tb, x3, N, P, p);
    CU_TRY(cudaGetLastError());
  }
  bt.d_hid = ts->next_drop(ts->p_sub);
  RC_TRY(layernorm<float>(e, s, x3, kD, e->lp_norm, bt.hn, kD, Mb, nullptr, nullptr));
  {
    cudaError_t err = gemm_simt<float, float>(s, bt.hn, kD, e->head1.w32, kD, e->head1.b, nullptr, 0, bt.hid, 200, Mb, 200, kD, 1, nullptr);
    if (err != cudaSuccess) return fail(BOFI_ERR_CUDA, "head GEMM: %s", cudaGetErrorString(err));
    e->launches++;
  }
  RC_TRY(dropout_inplace<float>(e, s, bt.hid, (size_t)Mb * 200, bt.d_hid));
  launch_k(xe_head_logp_kernel, ceil_div(Mb, 4), 128, 0, s, (const float*)bt.hid, e->w_len2, e->b_len2, e->w_syn2, e->b_syn2, 100, 20, 10,
           len_logp, syn_logp, Mb, P, Tb - 1);
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

// decode_SA / decode_NA (:520-530, :570-587) + decoder stack + final LayerNorm + vocab projection.
template <typename T>
static int t_dec_fwd(bofi_engine* e, cudaStream_t s, TrainState* ts, DecTape& dt, const int* word_ids, int const_word, const int* self_vis,
                     int vis_bs, int vis_qs, float* logp_out, bool keep_logits, Sampler sampler = Sampler(), long long* seq_out = nullptr,
                     const int* total_len = nullptr) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, T_ = ts->T, Tb = ts->Tb, rows = N * T_;
  const int* mem_len = ts->att_len;
  const int nb_layers = std::max(1, c.n_len);
  dt.x_in = aalloc<float>(ts, (size_t)rows * kD); A_TRY(dt.x_in);
  dt.d_embed = ts->next_drop(ts->p_sub);
  launch_k(embed_xe_kernel, ceil_div(rows, 8), 256, 0, s, W(e, "model.tgt_embed.lut.weight"), W(e, "model.syn_embed.lut.weight"),
           W(e, "model.pos_embed.pe"), word_ids, T_, 0, const_word, (const int*)ts->ext_syn, Tb, 1, sqrtf((float)kD), dt.x_in, rows, T_,
           dt.d_embed);
  CU_TRY(cudaGetLastError());
  dt.layers.assign(c.n_dec, LayerTape());
  const float* x = dt.x_in;
  for (int l = 0; l < c.n_dec; ++l) {
    RC_TRY(t_layer_fwd<T>(e, s, ts, e->dec[l], dt.layers[l], x, N, T_, self_vis, vis_bs, vis_qs, (const T*)ts->kv[nb_layers + l], ts->R,
                          mem_len, ts->spi));
    x = dt.layers[l].x_out;
  }
  T* yf = aalloc<T>(ts, (size_t)rows * kD); A_TRY(yf);
  dt.y_final = yf;
  RC_TRY(layernorm<T>(e, s, x, kD, e->dec_norm, yf, kD, rows, nullptr, nullptr));
  float* logits;
  if (keep_logits) {
    logits = aalloc<float>(ts, (size_t)rows * e->Vpad); A_TRY(logits);
    dt.logits = logits;
  } else {
    RC_TRY(e->logits.reserve((size_t)rows * e->Vpad * 4));
    logits = e->logits.as<float>();
    dt.logits = nullptr;
  }
  RC_TRY((linear<T, float>(e, s, yf, kD, e->generator, nullptr, 0, logits, e->Vpad, rows, 0, nullptr)));
  if (logp_out || seq_out) {
    launch_k(vocab_epilogue_kernel, rows, kVocabThreads, 0, s, (const float*)logits, e->Vpad, e->V, logp_out, seq_out,
             total_len, -1, T_, 1, (int*)nullptr, sampler, (float*)nullptr, (float*)nullptr, 0, 0, 0);
    CU_TRY(cudaGetLastError());
  }
  return BOFI_OK;
}

// _prepare_feature_forward + encoder (once per image) + the memory K/V of every decoder-style layer, with tape and dropout.
template <typename T>
static int t_encode_fwd(bofi_engine* e, cudaStream_t s, TrainState* ts, const float* att, const int* att_len) {
  const bofi_config_t& c = e->cfg;
  const int B = ts->B, R = ts->R, M = B * R, F = c.att_feat_size;
  // ---- _prepare_feature_forward + encoder (once per image) --------------------------------------------------
  const int* len_dev = nullptr;
  ts->have_len = (att_len != nullptr);
  ts->att_len = nullptr;
  if (att_len) {
    ts->att_len = aalloc<int>(ts, (size_t)B); A_TRY(ts->att_len);
    CU_TRY(cudaMemcpyAsync(ts->att_len, att_len, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
    len_dev = ts->att_len;
  }
  const T* a_in;
  if constexpr (std::is_same<T, bf16>::value) {
    T* attT = aalloc<T>(ts, (size_t)M * F); A_TRY(attT);
    const size_t n4 = (size_t)M * F / 4;
    launch_k(cast_kernel<bf16>, (int)std::min<size_t>((n4 + 255) / 256, 148 * 16), 256, 0, s, att, attT, n4);
    CU_TRY(cudaGetLastError());
    ts->attT = attT;
    a_in = attT;
  } else {
    ts->attT = (void*)att;         // fp32: the caller's tensor is the operand (must stay alive until backward)
    a_in = att;
  }
  ts->x0 = aalloc<float>(ts, (size_t)M * kD); A_TRY(ts->x0);
  ts->site = 0;
  ts->d_att_embed = ts->next_drop(ts->p_att);
  RC_TRY((linear<T, float>(e, s, a_in, F, e->att_embed, nullptr, 0, ts->x0, kD, M, 1, nullptr)));
  RC_TRY(dropout_inplace<float>(e, s, ts->x0, (size_t)M * kD, ts->d_att_embed));
  if (len_dev) {
    launch_k(zero_padded_rows_kernel, ceil_div(M, 8), 256, 0, s, ts->x0, len_dev, B, R);
    CU_TRY(cudaGetLastError());
  }
  ts->enc.assign(c.n_enc, LayerTape());
  const float* x = ts->x0;
  for (int l = 0; l < c.n_enc; ++l) {
    RC_TRY(t_layer_fwd<T>(e, s, ts, e->enc[l], ts->enc[l], x, B, R, len_dev, 1, 0, (const T*)nullptr, 0, nullptr, 1));
    x = ts->enc[l].x_out;
  }
  ts->enc_x_final = x;
  T* memT = aalloc<T>(ts, (size_t)M * kD); A_TRY(memT);
  ts->memT = memT;
  RC_TRY(layernorm<T>(e, s, x, kD, e->enc_norm, memT, kD, M, nullptr, nullptr));
  // ---- memory K/V of the bounding layer and of every decoder layer, once -------------------------------------
  const int nb_layers = std::max(1, c.n_len);
  ts->sc = false;
  ts->kv.assign(nb_layers + c.n_dec, nullptr);
  for (int l = 0; l < nb_layers + c.n_dec; ++l) {
    T* kv = aalloc<T>(ts, (size_t)M * 2 * kD); A_TRY(kv);
    ts->kv[l] = kv;
    const Lin& w = (l < nb_layers) ? e->lp[l].ca.kv : e->dec[l - nb_layers].ca.kv;
    RC_TRY((linear<T, T>(e, s, memT, kD, w, nullptr, 0, kv, 2 * kD, M, 0, nullptr)));
  }
  return BOFI_OK;
}

template <typename T>
static int train_forward_impl(bofi_engine* e, cudaStream_t s, TrainState* ts, const float* att, const int* att_len, float* sa_len,
                              float* sa_syn, float* sa_logp, float* na_len, float* na_syn, float* na_logp, bool fused_loss) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, Tb = ts->Tb;
  RC_TRY(t_encode_fwd<T>(e, s, ts, att, att_len));
  // ---- index bookkeeping ---------------------------------------------------------------------------------------
  launch_k(xe_prepare_kernel, ceil_div(N, 128), 128, 0, s, (const int*)ts->labels, (const int*)ts->pnum, (const int*)ts->plen, N, Tb, ts->P,
           c.len_idx, ts->word_seq, ts->vis_b, ts->na_vis);
  CU_TRY(cudaGetLastError());
  const size_t slots = (size_t)N * (Tb - 1);
  CU_TRY(cudaMemsetAsync(sa_len, 0, slots * 20 * 4, s));
  CU_TRY(cudaMemsetAsync(sa_syn, 0, slots * 10 * 4, s));
  CU_TRY(cudaMemsetAsync(na_len, 0, slots * 20 * 4, s));
  CU_TRY(cudaMemsetAsync(na_syn, 0, slots * 10 * 4, s));
  // ---- SA: bounding on the ground-truth words, decoder on the position-wise copied words ------------------
  const bool full_bound = c.n_len >= 2;
  if (full_bound) {
    ts->vis_full = aalloc<int>(ts, (size_t)ts->P * N * Tb); A_TRY(ts->vis_full);
    launch_k(xe_vis_full_kernel, ceil_div(ts->P * N * Tb, 256), 256, 0, s, (const int*)ts->vis_b, N, ts->P, Tb, ts->vis_full);
    CU_TRY(cudaGetLastError());
    RC_TRY(t_bound_fwd_full<T>(e, s, ts, ts->sa_b, ts->word_seq, nullptr, sa_len, sa_syn));
  } else {
    RC_TRY(t_bound_fwd<T>(e, s, ts, ts->sa_b, ts->word_seq, nullptr, sa_len, sa_syn));
  }
  RC_TRY(t_dec_fwd<T>(e, s, ts, ts->sa_d, ts->ext_seq, -1, ts->sa_vis, ts->T, 1, sa_logp, fused_loss));
  // ---- NA: bounding on the syn labels, decoder on BOS + syn labels ------------------------------------------
  if (full_bound) RC_TRY(t_bound_fwd_full<T>(e, s, ts, ts->na_b, nullptr, ts->ext_syn, na_len, na_syn));
  else RC_TRY(t_bound_fwd<T>(e, s, ts, ts->na_b, nullptr, ts->ext_syn, na_len, na_syn));
  ts->glat_words = nullptr;
  if (ts->glat_p >= 0.f) {
    // Glancing (:437-464): a no-grad NA pass on constant bos inputs predicts the words; the fraction of mismatches times glat_p is
    // the probability with which a slot's ground-truth word replaces bos in the input of the real (taped) NA pass.  The no-grad
    // pass draws its own dropout masks like the reference's (it runs in train() mode there too) and leaves nothing on the tape.
    const int T_ = ts->T;
    int* words = aalloc<int>(ts, (size_t)N * T_); A_TRY(words);
    long long* tok = aalloc<long long>(ts, (size_t)N * T_); A_TRY(tok);
    const Arena::Mark mark = ts->arena.mark();
    DecTape scratch;
    RC_TRY(t_dec_fwd<T>(e, s, ts, scratch, nullptr, c.bos_idx, ts->na_vis, 1, 0, nullptr, false, Sampler(), tok, nullptr));
    ts->arena.rewind(mark);
    launch_k(glat_input_kernel, ceil_div(N, 8), 256, 0, s, (const int*)ts->labels, (const int*)ts->plen, (const long long*)tok, N, T_, Tb, c.bos_idx,
             ts->glat_p, drop_hash(ts->glat_seed, 0x474C4154u), words);
    CU_TRY(cudaGetLastError());
    ts->glat_words = words;
    RC_TRY(t_dec_fwd<T>(e, s, ts, ts->na_d, words, -1, ts->na_vis, 1, 0, na_logp, fused_loss));
  } else {
    RC_TRY(t_dec_fwd<T>(e, s, ts, ts->na_d, nullptr, c.bos_idx, ts->na_vis, 1, 0, na_logp, fused_loss));
  }
  return BOFI_OK;
}

// ---- backward helpers ----------------------------------------------------------------------------------------
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// C[M,N] (TOut) = A[M,K] . Bm[N,K]^T (+ resid), no bias: the generic NT contraction behind dgrad and wgrad.
template <typename T, typename TOut>
static int gemm_nt(bofi_engine* e, cudaStream_t s, TrainState* ts, const T* A, int lda, const T* Bm, int ldb, int M, int N, int K,
                   TOut* out, int ldc, const float* resid, int ldr) {
  Lin l;
  l.N = N;
  l.K = ldb;          // linear() passes l.K as the weight pitch ...
  if constexpr (std::is_same<T, float>::value) l.w32 = Bm; else l.w16 = Bm;
  l.b = ts->zeros.as<float>();
  // ... and as the contraction length; pitches are always >= K and the pad columns are zero on both operands
  if (ldb != K) return fail(BOFI_ERR_INVALID, "gemm_nt: operand pitch %d != contraction %d", ldb, K);
  return linear<T, TOut>(e, s, A, lda, l, resid, ldr, out, ldc, M, 0, nullptr);
}

template <typename TIn, typename TOut>
static int transpose_pad(bofi_engine* e, cudaStream_t s, const TIn* in, int ld_in, TOut* out, int rows, int cols, int rows_pad) {
  ProfScope prof(e, s, PC_OTHER, 0.0, (double)rows * cols * sizeof(TIn) + (double)rows_pad * cols * sizeof(TOut));
  launch_k(transpose_pad_kernel<TIn, TOut>, dim3(ceil_div(cols, 32), ceil_div(rows_pad, 32)), dim3(32, 8), 0, s, in, ld_in, out, rows, cols,
           rows_pad);
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

// gb[c] += sum_r dY[r, c] in two fixed-order stages (row chunks, then the chunk partials).
template <typename T>
static int colsum(bofi_engine* e, cudaStream_t s, TrainState* ts, const T* dY, int ldy, int M, int N, float* gb) {
  const int cols4 = round_up(N, 4);
  if (cols4 > ldy) return fail(BOFI_ERR_INVALID, "colsum: pitch %d < %d", ldy, cols4);
  const int col_blocks = ceil_div(cols4, 256);
  const int chunks = std::max(1, std::min(std::min(64, M / 64), (4 * 148) / col_blocks));
  RC_TRY(ts->cs_partial.reserve((size_t)chunks * cols4 * 4));
  if (ts->cs_tickets.cap == 0) {
    RC_TRY(ts->cs_tickets.reserve(256 * 4));
    CU_TRY(cudaMemsetAsync(ts->cs_tickets.p, 0, 256 * 4, s));
  }
  {
    ProfScope prof(e, s, PC_OTHER, 0.0, (double)M * N * sizeof(T));
    launch_k(colsum_kernel<T>, dim3(col_blocks, chunks), 512, 0, s, dY, ldy, M, N, cols4, ts->cs_partial.as<float>(), ts->cs_tickets.as<int>(), gb);
  }
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

// g_lut[id] += sqrt(d) * sum of the dx rows with that id (tiny id space), two fixed-order stages.
static int embed_small_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, const float* dx, const int* ids, int ids_stride, int ids_off,
                           int const_id, int T_, int rows, float* g_lut) {
  const int chunks = std::max(1, std::min(64, rows / 128));
  RC_TRY(ts->cs_partial.reserve((size_t)chunks * kSmallIds * kD * 4));
  {
    ProfScope prof(e, s, PC_OTHER, 0.0, (double)rows * kD * 4);
    launch_k(embed_small_partial_kernel, dim3(kD / 64, chunks), 256, 0, s, dx, ids, ids_stride, ids_off, T_, rows, ts->cs_partial.as<float>());
  }
  CU_TRY(cudaGetLastError());
  {
    ProfScope prof(e, s, PC_OTHER, 0.0, 0.0);
    const int n_ids = ids ? kSmallIds : 1;
    launch_k(embed_small_reduce_kernel, ceil_div(n_ids * kD, 256), 256, 0, s, (const float*)ts->cs_partial.as<float>(), chunks, n_ids,
             ids ? -1 : const_id, sqrtf((float)kD), g_lut);
  }
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

// Backward of y = x . W^T + b for one (possibly fused) Linear:  gW += dY^T . X,  gb += colsum(dY),  dX = dY . W.
//   X [M, K] pitch ldx, dY [M, N] pitch ldy.
// bf16 (tcgen05): both contractions read their operands as stored -- MN-major UMMA descriptors, no transposed copies.
// fp32 (FFMA parity backend): explicit zero-padded transposes feed the NT kernel (dY pad columns up to
// round_up(N, 64) must be zero when N % 64 != 0).
template <typename T>
static int lin_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, const Lin& lin, const T* X, int ldx, const T* dY, int ldy, int M,
                   T* dX, int lddx) {
  const int N = lin.N, K = lin.K, Mp = round_up(M, 64);
  const bool tc = std::is_same<T, bf16>::value && e->use_tc;
  if (lin.gw) {
    if constexpr (std::is_same<T, bf16>::value) {
      if (tc) {
        ProfScope prof(e, s, PC_GEMM_TC, 2.0 * M * N * K, 2.0 * ((double)M * N + (double)M * K) + 8.0 * N * K, N, K, M);
        cudaError_t err = tc::gemm_tc_wgrad(s, dY, ldy, X, ldx, ts->zeros.as<float>(), lin.gw, K, N, K, M);
        if (err != cudaSuccess) return fail(BOFI_ERR_CUDA, "wgrad GEMM %dx%dx%d: %s", N, K, M, cudaGetErrorString(err));
      }
    }
    if (!tc) {
      RC_TRY(ts->tr_a.reserve((size_t)N * Mp * sizeof(T)));
      RC_TRY(ts->tr_b.reserve((size_t)K * Mp * sizeof(T)));
      RC_TRY((transpose_pad<T, T>(e, s, dY, ldy, ts->tr_a.as<T>(), M, N, Mp)));
      RC_TRY((transpose_pad<T, T>(e, s, X, ldx, ts->tr_b.as<T>(), M, K, Mp)));
      RC_TRY((gemm_nt<T, float>(e, s, ts, ts->tr_a.as<T>(), Mp, ts->tr_b.as<T>(), Mp, N, K, Mp, lin.gw, K, lin.gw, K)));
    }
    RC_TRY(colsum<T>(e, s, ts, dY, ldy, M, N, lin.gb));
  }
  if (dX) {
    if constexpr (std::is_same<T, bf16>::value) {
      if (tc) {
        ProfScope prof(e, s, PC_GEMM_TC, 2.0 * M * N * K, 2.0 * ((double)M * N + (double)N * K + (double)M * K), M, K, N);
        cudaError_t err = tc::gemm_tc_dgrad(s, dY, ldy, lin.w16, K, ts->zeros.as<float>(), dX, lddx, M, K, N);
        if (err != cudaSuccess) return fail(BOFI_ERR_CUDA, "dgrad GEMM %dx%dx%d: %s", M, K, N, cudaGetErrorString(err));
        return BOFI_OK;
      }
    }
    const int Np = round_up(N, 64);
    if (ldy < Np) return fail(BOFI_ERR_INVALID, "lin_bwd: dY pitch %d < padded N %d", ldy, Np);
    RC_TRY(ts->tr_w.reserve((size_t)K * Np * sizeof(T)));
    if constexpr (std::is_same<T, float>::value) {
      RC_TRY((transpose_pad<float, float>(e, s, lin.w32, K, ts->tr_w.as<float>(), N, K, Np)));
    } else {
      RC_TRY((transpose_pad<bf16, bf16>(e, s, lin.w16, K, ts->tr_w.as<bf16>(), N, K, Np)));
    }
    RC_TRY((gemm_nt<T, T>(e, s, ts, dY, ldy, ts->tr_w.as<T>(), Np, M, K, Np, dX, lddx, nullptr, 0)));
  }
  return BOFI_OK;
}

// (dx, dxT) = dres + LayerNorm backward; parameter gradients through a fixed-order two-stage reduction.
template <typename TG, typename T>
static int ln_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, const Norm& n, const float* x, const TG* dy, const float* dres,
                  float* dx_out, T* dx_out_t, int rows) {
  const int grid = std::min(ceil_div(rows, 8), 148 * 2);
  RC_TRY(ts->ln_partial.reserve((size_t)grid * 2 * kD * 4));
  {
    ProfScope prof(e, s, PC_LAYERNORM, 0.0, (double)rows * kD * (4 + sizeof(TG) + 4 + 4 + sizeof(T)));
    launch_k(layernorm_bwd_kernel<TG, T>, grid, 256, 0, s, x, n.a, dy, dres, dx_out, dx_out_t, rows, ts->ln_partial.as<float>());
  }
  CU_TRY(cudaGetLastError());
  {
    ProfScope prof(e, s, PC_LAYERNORM, 0.0, 0.0);
    launch_k(ln_param_reduce_kernel, 2 * kD / 32, dim3(32, 8), 0, s, (const float*)ts->ln_partial.as<float>(), grid, n.ga, n.gb);
  }
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

template <typename T>
static int attn_bwd(bofi_engine* e, cudaStream_t s, const T* Q, int ldq, const T* K, const T* V, int ldkv, const T* dO, int ldo, T* dQ,
                    int lddq, T* dK, T* dV, int lddkv, int n_kv_blocks, int Tq, int Tk, int qpk, const int* vis, int vis_bs, int vis_qs,
                    int vis_div, int accumulate_kv, Drop drop = Drop()) {
  if (n_kv_blocks <= 0) return BOFI_OK;
  if (Tk > kMaxKeys) return fail(BOFI_ERR_INVALID, "attention backward over %d keys (max %d)", Tk, kMaxKeys);
  const float scale = 1.0f / sqrtf((float)kHeadDim);
  const size_t smem = attention_bwd_smem_bytes(Tk);
  dim3 grid(e->cfg.heads, n_kv_blocks);
  ProfScope prof(e, s, PC_ATTENTION, 10.0 * n_kv_blocks * qpk * Tq * Tk * kD, 0.0, n_kv_blocks, qpk * Tq, Tk);
  if constexpr (std::is_same<T, bf16>::value) {
    if (!e->attn_simt_only) {
      const int KT = (Tk + 15) / 16;
      const size_t sm = attention_bwd_mma_smem_bytes(KT);
      cudaError_t err = cudaErrorInvalidValue;
#define BOFI_ATTBM(n)                                                                                                                  \
  case n: {                                                                                                                            \
    static PerDevice<bool> configured_dev;                                                                                             \
    bool& configured = configured_dev.get();                                                                                           \
    if (!configured) {                                                                                                                 \
      CU_TRY(cudaFuncSetAttribute(attention_bwd_mma_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));                 \
      configured = true;                                                                                                               \
    }                                                                                                                                  \
    err = launch_k(attention_bwd_mma_kernel<n>, grid, 128, sm, s, Q, ldq, K, V, ldkv, dO, ldo, dQ, lddq, dK, dV, lddkv, Tq, Tk, qpk, vis, \
                   vis_bs, vis_qs, vis_div, scale, accumulate_kv, drop);                                                               \
  } break;
      switch (KT) { BOFI_ATTBM(1) BOFI_ATTBM(2) BOFI_ATTBM(3) BOFI_ATTBM(4) BOFI_ATTBM(5) BOFI_ATTBM(6) BOFI_ATTBM(7) BOFI_ATTBM(8) }
#undef BOFI_ATTBM
      if (err != cudaSuccess) return fail(BOFI_ERR_CUDA, "attention_bwd_mma: %s", cudaGetErrorString(err));
      return BOFI_OK;
    }
  }
#define BOFI_ATTB(KPT_)                                                                                                             \
  do {                                                                                                                              \
    static PerDevice<size_t> configured_dev;                                                                                        \
    size_t& configured = configured_dev.get();                                                                                      \
    if (smem > configured) {                                                                                                        \
      CU_TRY(cudaFuncSetAttribute(attention_bwd_kernel<T, KPT_>, cudaFuncAttributeMaxDynamicSharedMemorySize,                       \
                                  (int)attention_bwd_smem_bytes(4 * KPT_)));                                                        \
      configured = attention_bwd_smem_bytes(4 * KPT_);                                                                              \
    }                                                                                                                               \
    launch_k(attention_bwd_kernel<T, KPT_>, grid, 256, smem, s, Q, ldq, K, V, ldkv, dO, ldo, dQ, lddq, dK, dV, lddkv, Tq, Tk, qpk, vis, \
             vis_bs, vis_qs, vis_div, scale, accumulate_kv, drop);                                                                  \
  } while (0)
  if (Tk <= 24) BOFI_ATTB(6);
  else if (Tk <= 40) BOFI_ATTB(10);
  else if (Tk <= 64) BOFI_ATTB(16);
  else BOFI_ATTB(32);
#undef BOFI_ATTB
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

// Backward of one pre-norm layer.  In: dx (fp32) / dxT (T) = gradient of the layer output.  Out: the same pair for the
// layer input (written into dx / dxT in place).  Cross-attention K/V gradients accumulate into dkv_mem [B*R, 1024].
// Row-only mode (`bound`): the self-attention sublayer is handled by the caller.
template <typename T>
static int t_layer_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, const Layer& ly, const LayerTape& tp, float* dx, T* dxT, T* g1, T* g2,
                       int nb, int Tq, const int* self_vis, int vis_bs, int vis_qs, const T* kvmem, T* dkv_mem, int dkv_accumulate, int R,
                       const int* mem_len, int qpk, bool skip_self) {
  const int rows = nb * Tq, dff = e->cfg.d_ff;
  const float* xm = ly.cross ? tp.x2 : tp.x1;
  const int f = ly.cross ? 2 : 1;
  // FFN: x_out = xm + drop(w2(drop(relu(w1(LN(xm))))))
  const T* dz = nullptr;
  RC_TRY(sublayer_grad<T>(e, s, ts, dxT, rows, tp.d_ffn_out, &dz));
  RC_TRY(lin_bwd<T>(e, s, ts, ly.w2, (const T*)tp.ffh, dff, dz, kD, rows, g2, dff));            // g2 = d ffh [rows, dff]
  e->launches++;
  launch_k(relu_bwd_kernel<T>, 148 * 8, 256, 0, s, g2, (const T*)tp.ffh, (size_t)rows * dff / 4, tp.d_ffh.scale);
  CU_TRY(cudaGetLastError());
  RC_TRY(lin_bwd<T>(e, s, ts, ly.w1, (const T*)tp.y2, kD, g2, dff, rows, g1, kD));               // g1 = d y2
  RC_TRY((ln_bwd<T, T>(e, s, ts, ly.ln[f], xm, g1, dx, dx, dxT, rows)));
  if (ly.cross) {
    // x2 = x1 + o(attn(q(LN(x1)), kv(mem)))
    RC_TRY(sublayer_grad<T>(e, s, ts, dxT, rows, tp.d_ca_out, &dz));
    RC_TRY(lin_bwd<T>(e, s, ts, ly.ca.o, (const T*)tp.ao2, kD, dz, kD, rows, g1, kD));           // g1 = d ao2
    // g2 reused as dq [rows, 512]
    RC_TRY(attn_bwd<T>(e, s, (const T*)tp.q, kD, kvmem, kvmem + kD, 2 * kD, g1, kD, g2, kD, dkv_mem, dkv_mem + kD, 2 * kD,
                       rows / (qpk * Tq), Tq, R, qpk, mem_len, 1, 0, qpk, dkv_accumulate, tp.d_ca_att));
    RC_TRY(lin_bwd<T>(e, s, ts, ly.ca.q, (const T*)tp.y1, kD, g2, kD, rows, g1, kD));            // g1 = d y1
    RC_TRY((ln_bwd<T, T>(e, s, ts, ly.ln[1], tp.x1, g1, dx, dx, dxT, rows)));
  }
  if (skip_self) return BOFI_OK;
  // x1 = x_in + o(attn(qkv(LN(x_in))))
  RC_TRY(sublayer_grad<T>(e, s, ts, dxT, rows, tp.d_sa_out, &dz));
  RC_TRY(lin_bwd<T>(e, s, ts, ly.sa.o, (const T*)tp.ao, kD, dz, kD, rows, g1, kD));               // g1 = d ao
  const T* qkv = (const T*)tp.qkv;
  RC_TRY(attn_bwd<T>(e, s, qkv, 3 * kD, qkv + kD, qkv + 2 * kD, 3 * kD, g1, kD, g2, 3 * kD, g2 + kD, g2 + 2 * kD, 3 * kD, nb, Tq, Tq, 1,
                     self_vis, vis_bs, vis_qs, 1, 0, tp.d_sa_att));                              // g2 = d qkv [rows, 1536]
  RC_TRY(lin_bwd<T>(e, s, ts, ly.sa.qkv, (const T*)tp.y0, kD, g2, 3 * kD, rows, g1, kD));        // g1 = d y0
  RC_TRY((ln_bwd<T, T>(e, s, ts, ly.ln[0], tp.x_in, g1, dx, dx, dxT, rows)));
  return BOFI_OK;
}

// Backward of one decoder pass.  dz [rows, Vp64] = gradient w.r.t. the logits (pad columns zero).
template <typename T>
static int t_dec_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, DecTape& dt, T* dz, int ldz, const int* word_ids, bool const_word,
                     const int* self_vis, int vis_bs, int vis_qs, float* dx, T* dxT, T* g1, T* g2, int first_pass, int word_stride = -1,
                     int word_off = 0) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, T_ = ts->T, Tb = ts->Tb, rows = N * T_;
  const int* mem_len = ts->att_len;
  const int nb_layers = std::max(1, c.n_len);
  RC_TRY(lin_bwd<T>(e, s, ts, e->generator, (const T*)dt.y_final, kD, dz, ldz, rows, g1, kD));     // g1 = d y_final
  const float* x_last = dt.layers.back().x_out;
  RC_TRY((ln_bwd<T, T>(e, s, ts, e->dec_norm, x_last, g1, nullptr, dx, dxT, rows)));
  for (int l = c.n_dec - 1; l >= 0; --l) {
    T* dkv = ts->dkv.as<T>() + (size_t)(nb_layers + l) * ts->B * ts->R * 2 * kD;
    RC_TRY(t_layer_bwd<T>(e, s, ts, e->dec[l], dt.layers[l], dx, dxT, g1, g2, N, T_, self_vis, vis_bs, vis_qs, (const T*)ts->kv[nb_layers + l],
                          dkv, first_pass ? 0 : 1, ts->R, mem_len, ts->spi, false));
  }
  // input embeddings: x = drop(tgt_embed(word)*sqrt(d) + syn_embed(syn)*sqrt(d) + pe)
  if (dt.d_embed.thresh) RC_TRY((dropout_apply<float, float>(e, s, dx, dx, (size_t)rows * kD, dt.d_embed)));
  const float sq = sqrtf((float)kD);
  if (const_word) {
    RC_TRY(embed_small_bwd(e, s, ts, dx, nullptr, 0, 0, c.bos_idx, T_, rows, G(e, "model.tgt_embed.lut.weight")));
  } else {
    ProfScope prof(e, s, PC_OTHER, 0.0, (double)rows * kD * 8);
    launch_k(embed_bwd_kernel, ceil_div(rows, 8), 256, 0, s, (const float*)dx, word_ids, word_stride < 0 ? T_ : word_stride, word_off, T_, rows, sq,
             G(e, "model.tgt_embed.lut.weight"));
  }
  CU_TRY(cudaGetLastError());
  RC_TRY(embed_small_bwd(e, s, ts, dx, ts->ext_syn, Tb, 1, 0, T_, rows, G(e, "model.syn_embed.lut.weight")));
  return BOFI_OK;
}

// N_len >= 2: backward of t_bound_fwd_full below the heads.  dx [Mb, 512] holds d x3 (the [LEN] rows of every pass) on entry.
template <typename T>
static int t_bound_bwd_full_tail(bofi_engine* e, cudaStream_t s, TrainState* ts, BoundTape& bt, const int* word_ids, const int* syn_ids,
                                 float* dx, T* dxT, T* g1, T* g2, int first_pass) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, Tb = ts->Tb, P = ts->P, Mb = ts->Mb, NL = c.n_len;
  const int* mem_len = ts->att_len;
  const size_t M = (size_t)ts->B * ts->R;
  RC_TRY(ts->dlen.reserve((size_t)Mb * kD * 4));
  float* dlen = ts->dlen.as<float>();
  CU_TRY(cudaMemcpyAsync(dlen, dx, (size_t)Mb * kD * 4, cudaMemcpyDeviceToDevice, s));
  RC_TRY(ts->dmem.reserve((size_t)std::max(N * Tb, ts->B * ts->R) * kD * 4));
  float* dxin = ts->dmem.as<float>();
  const size_t n4 = (size_t)N * Tb * (kD / 4);
  for (int p = P - 1; p >= 0; --p) {
    const int* vis = ts->vis_full + (size_t)p * N * Tb;
    e->launches += 2;
    launch_k(scatter_len_rows_kernel<T>, (unsigned)ceil_div(n4, (size_t)256), 256, 0, s, (const float*)dlen, Tb, dx, dxT, N, P, p);
    CU_TRY(cudaGetLastError());
    for (int l = NL - 1; l >= 0; --l) {
      T* dkv = ts->dkv.as<T>() + (size_t)l * M * 2 * kD;
      const int accumulate = (first_pass && p == P - 1) ? 0 : 1;       // the first pass to touch a layer's K/V gradient overwrites
      RC_TRY(t_layer_bwd<T>(e, s, ts, e->lp[l], bt.full[(size_t)p * NL + l], dx, dxT, g1, g2, N, Tb, vis, Tb, 1, (const T*)ts->kv[l], dkv,
                            accumulate, ts->R, mem_len, ts->spi, false));
    }
    launch_k(add_inplace_f32_kernel, 148 * 8, 256, 0, s, dxin, (const float*)dx, n4, p == P - 1 ? 1 : 0);   // d x_in: sum over the passes
    CU_TRY(cudaGetLastError());
  }
  if (bt.d_embed.thresh) RC_TRY((dropout_apply<float, float>(e, s, dxin, dxin, (size_t)N * Tb * kD, bt.d_embed)));
  const float sq = sqrtf((float)kD);
  if (word_ids) {
    ProfScope prof(e, s, PC_OTHER, 0.0, (double)N * Tb * kD * 8);
    launch_k(embed_bwd_kernel, ceil_div(N * Tb, 8), 256, 0, s, (const float*)dxin, word_ids, Tb, 0, Tb, N * Tb, sq, G(e, "model.tgt_embed.lut.weight"));
  } else {
    RC_TRY(embed_small_bwd(e, s, ts, dxin, syn_ids, Tb, 0, 0, Tb, N * Tb, G(e, "model.syn_embed.lut.weight")));
  }
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

// Backward of one teacher-forced bounding pass batch.  g_len / g_syn: gradients of the [N, Tb-1, .] outputs.
template <typename T>
static int t_bound_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, BoundTape& bt, const float* g_len, const float* g_syn,
                       const float* len_logp, const float* syn_logp, const int* word_ids, const int* syn_ids, float* dx, T* dxT, T* g1, T* g2,
                       int first_pass) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, Tb = ts->Tb, P = ts->P, Mb = ts->Mb;
  const Layer& ly = e->lp[0];
  const int* mem_len = ts->att_len;
  const std::string lp = "model.length_predictor";
  // heads (fp32): dz -> classifier2 grads -> d hid -> classifier1 grads -> d hn
  RC_TRY(ts->scratch_f32.reserve(((size_t)Mb * 32 + (size_t)Mb * 256 + (size_t)Mb * kD) * 4));
  float* dzh = ts->scratch_f32.as<float>();
  float* dhid = dzh + (size_t)Mb * 32;               // [Mb, 256] pitch (pad columns zero)
  float* dhn = dhid + (size_t)Mb * 256;              // [Mb, 512]
  e->launches += 3;
  launch_k(xe_head_dz_kernel, ceil_div(Mb, 4), 128, 0, s, g_len, g_syn, len_logp, syn_logp, 20, 10, dzh, Mb, P, Tb - 1);
  CU_TRY(cudaGetLastError());
  {
    const int chunks = std::max(1, std::min(64, Mb / 128));
    RC_TRY(ts->cs_partial.reserve((size_t)chunks * 30 * 128 * 4));
    launch_k(xe_head2_wgrad_partial_kernel, dim3(30, chunks), 128, 0, s, (const float*)dzh, (const float*)bt.hid, 100, 20, Mb,
             ts->cs_partial.as<float>());
    CU_TRY(cudaGetLastError());
    launch_k(xe_head2_wgrad_reduce_kernel, 30, 128, 0, s, (const float*)ts->cs_partial.as<float>(), chunks, 100, 20,
             G(e, lp + ".Length_classifier2.weight"), G(e, lp + ".Length_classifier2.bias"), G(e, lp + ".Syntactic_classifier2.weight"),
             G(e, lp + ".Syntactic_classifier2.bias"));
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaMemsetAsync(dhid, 0, (size_t)Mb * 256 * 4, s));
  launch_k(xe_head2_dgrad_kernel, ceil_div((size_t)Mb * 200, 256), 256, 0, s, (const float*)dzh, (const float*)bt.hid, e->w_len2, e->w_syn2, 100,
           20, 10, dhid, 256, Mb, bt.d_hid.scale);
  CU_TRY(cudaGetLastError());
  const LayerTape& tp = bt.lt;
  bool head_tc = false;
  if constexpr (std::is_same<T, bf16>::value) {
    if (e->use_tc && e->head1_w16) {
      // classifier1 backward on the tensor cores: bf16 copies of d hid / hn (two small casts), fp32 accumulation into gW
      head_tc = true;
      bf16* dhid16 = g2;                               // [Mb, 256]
      bf16* hn16 = g2 + (size_t)Mb * 256;              // [Mb, 512]
      bf16* dhn16 = g1;                                // [Mb, 512]
      e->launches += 2;
      launch_k(cast_kernel<bf16>, 148 * 4, 256, 0, s, (const float*)dhid, dhid16, (size_t)Mb * 256 / 4);
      CU_TRY(cudaGetLastError());
      launch_k(cast_kernel<bf16>, 148 * 4, 256, 0, s, (const float*)bt.hn, hn16, (size_t)Mb * kD / 4);
      CU_TRY(cudaGetLastError());
      {
        ProfScope prof(e, s, PC_GEMM_TC, 2.0 * Mb * 200 * kD, 0.0, 200, kD, Mb);
        cudaError_t err = tc::gemm_tc_wgrad(s, dhid16, 256, hn16, kD, ts->zeros.as<float>(), e->head1.gw, kD, 200, kD, Mb);
        if (err != cudaSuccess) return fail(BOFI_ERR_CUDA, "head wgrad: %s", cudaGetErrorString(err));
      }
      RC_TRY(colsum<float>(e, s, ts, dhid, 256, Mb, 200, e->head1.gb));
      {
        ProfScope prof(e, s, PC_GEMM_TC, 2.0 * Mb * 200 * kD, 0.0, Mb, kD, 200);
        cudaError_t err = tc::gemm_tc_dgrad(s, dhid16, 256, e->head1_w16, kD, ts->zeros.as<float>(), dhn16, kD, Mb, kD, 200);
        if (err != cudaSuccess) return fail(BOFI_ERR_CUDA, "head dgrad: %s", cudaGetErrorString(err));
      }
      RC_TRY((ln_bwd<bf16, T>(e, s, ts, e->lp_norm, tp.x_out, dhn16, nullptr, dx, dxT, Mb)));
    }
  }
  if (!head_tc) {
    RC_TRY(lin_bwd<float>(e, s, ts, e->head1, bt.hn, kD, dhid, 256, Mb, dhn, kD));
    RC_TRY((ln_bwd<float, T>(e, s, ts, e->lp_norm, tp.x_out, dhn, nullptr, dx, dxT, Mb)));
  }
  if (c.n_len >= 2) return t_bound_bwd_full_tail<T>(e, s, ts, bt, word_ids, syn_ids, dx, dxT, g1, g2, first_pass);
  // FFN + cross-attention of the (n, p) rows; K/V block = image, spi*P single-query rows per block
  T* dkv = ts->dkv.as<T>();
  RC_TRY(t_layer_bwd<T>(e, s, ts, ly, tp, dx, dxT, g1, g2, Mb, 1, nullptr, 0, 0, (const T*)ts->kv[0], dkv, first_pass ? 0 : 1, ts->R, mem_len,
                        ts->spi * P, true));
  // self-attention sublayer: x1[(n,p)] = x_in[n, 0] + o(attn(q[n,0], K/V[n, :vis_b[n,p]]))
  const T* dz = nullptr;
  RC_TRY(sublayer_grad<T>(e, s, ts, dxT, Mb, tp.d_sa_out, &dz));
  RC_TRY(lin_bwd<T>(e, s, ts, ly.sa.o, (const T*)tp.ao, kD, dz, kD, Mb, g1, kD));                 // g1 = d ao [Mb, 512]
  T* q_rep = (T*)bt.q_rep;
  const T* qkv = (const T*)bt.qkv;
  e->launches += 2;
  launch_k(repeat_row_kernel<T>, ceil_div(Mb, 8), 256, 0, s, qkv, (size_t)Tb * 3 * kD, q_rep, N, P);
  CU_TRY(cudaGetLastError());
  // g2 layout: [0, Mb*512) = dq per (n,p) ; then d qkv [N*Tb, 1536]
  T* dq_rep = g2;
  T* dqkv = g2 + (size_t)Mb * kD;
  CU_TRY(cudaMemsetAsync(dqkv, 0, (size_t)N * Tb * 3 * kD * sizeof(T), s));
  RC_TRY(attn_bwd<T>(e, s, q_rep, kD, qkv + kD, qkv + 2 * kD, 3 * kD, g1, kD, dq_rep, kD, dqkv + kD, dqkv + 2 * kD, 3 * kD, N, 1, Tb, P, ts->vis_b,
                     1, 0, 1, 0, tp.d_sa_att));
  launch_k(sum_over_passes_kernel<T, T>, ceil_div(N, 8), 256, 0, s, (const T*)dq_rep, P, dqkv, (size_t)Tb * 3 * kD, N, 0);
  CU_TRY(cudaGetLastError());
  RC_TRY(lin_bwd<T>(e, s, ts, ly.sa.qkv, (const T*)bt.y0, kD, dqkv, 3 * kD, N * Tb, g1, kD));     // g1 = d y0 [N*Tb, 512]
  // d x_in = LN0 backward (all rows) + the residual of row 0: sum over passes of dx
  RC_TRY(ts->dmem.reserve((size_t)std::max(N * Tb, ts->B * ts->R) * kD * 4));
  float* dxin = ts->dmem.as<float>();
  RC_TRY((ln_bwd<T, T>(e, s, ts, ly.ln[0], bt.x_in, g1, nullptr, dxin, (T*)nullptr, N * Tb)));
  e->launches += 1;
  launch_k(sum_over_passes_kernel<float, float>, ceil_div(N, 8), 256, 0, s, (const float*)dx, P, dxin, (size_t)Tb * kD, N, 1);
  CU_TRY(cudaGetLastError());
  if (bt.d_embed.thresh) RC_TRY((dropout_apply<float, float>(e, s, dxin, dxin, (size_t)N * Tb * kD, bt.d_embed)));
  const float sq = sqrtf((float)kD);
  if (word_ids) {
    ProfScope prof(e, s, PC_OTHER, 0.0, (double)N * Tb * kD * 8);
    launch_k(embed_bwd_kernel, ceil_div(N * Tb, 8), 256, 0, s, (const float*)dxin, word_ids, Tb, 0, Tb, N * Tb, sq, G(e, "model.tgt_embed.lut.weight"));
  } else {
    RC_TRY(embed_small_bwd(e, s, ts, dxin, syn_ids, Tb, 0, 0, Tb, N * Tb, G(e, "model.syn_embed.lut.weight")));
  }
  CU_TRY(cudaGetLastError());
  return BOFI_OK;
}

template <typename T>
static int t_encode_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, float* dx, T* dxT, T* g1, T* g2, int first_kv);

// Backward of the whole step.  dz_sa / dz_na: logit gradients [N*T, Vp64] (T-typed, pad columns zero).
template <typename T>
static int train_backward_impl(bofi_engine* e, cudaStream_t s, TrainState* ts, T* dz_sa, T* dz_na, int ldz, const float* g_sa_len,
                               const float* g_sa_syn, const float* g_na_len, const float* g_na_syn, const float* sa_len, const float* sa_syn,
                               const float* na_len, const float* na_syn) {
  const bofi_config_t& c = e->cfg;
  const int B = ts->B, R = ts->R, M = B * R, N = ts->N, T_ = ts->T, Tb = ts->Tb, Mb = ts->Mb, F = c.att_feat_size;
  const int nb_layers = std::max(1, c.n_len);
  const size_t big = std::max(std::max((size_t)N * T_, (size_t)Mb), std::max((size_t)M, (size_t)N * Tb));
  // workspace: dx fp32, dxT, g1 [big,512], g2 [big, max(dff, 1536) + ...]
  RC_TRY(e->x.reserve(big * kD * 4));
  RC_TRY(e->y.reserve(big * kD * sizeof(T)));
  RC_TRY(e->q.reserve(big * kD * sizeof(T)));
  RC_TRY(e->ffh.reserve((big * std::max(c.d_ff, 3 * kD) + (size_t)Mb * kD) * sizeof(T)));
  RC_TRY(ts->dkv.reserve((size_t)(nb_layers + c.n_dec) * M * 2 * kD * sizeof(T)));
  float* dx = e->x.as<float>();
  T* dxT = e->y.as<T>();
  T* g1 = e->q.as<T>();
  T* g2 = e->ffh.as<T>();
  // the four decoder-style consumers of `memory`; the first one to touch a K/V gradient overwrites, the rest add
  RC_TRY(t_dec_bwd<T>(e, s, ts, ts->na_d, dz_na, ldz, ts->glat_words, ts->glat_words == nullptr, ts->na_vis, 1, 0, dx, dxT, g1, g2, 1));
  RC_TRY(t_dec_bwd<T>(e, s, ts, ts->sa_d, dz_sa, ldz, ts->ext_seq, false, ts->sa_vis, T_, 1, dx, dxT, g1, g2, 0));
  RC_TRY(t_bound_bwd<T>(e, s, ts, ts->na_b, g_na_len, g_na_syn, na_len, na_syn, nullptr, ts->ext_syn, dx, dxT, g1, g2, 1));
  RC_TRY(t_bound_bwd<T>(e, s, ts, ts->sa_b, g_sa_len, g_sa_syn, sa_len, sa_syn, ts->word_seq, nullptr, dx, dxT, g1, g2, 0));
  return t_encode_bwd<T>(e, s, ts, dx, dxT, g1, g2, 0);
}

// Backward of everything t_encode_fwd computed: the memory K/V projections of the layers [first_kv, ...) whose K/V gradients
// the decoder-side passes left in ts->dkv, the encoder and att_embed.
template <typename T>
static int t_encode_bwd(bofi_engine* e, cudaStream_t s, TrainState* ts, float* dx, T* dxT, T* g1, T* g2, int first_kv) {
  const bofi_config_t& c = e->cfg;
  const int B = ts->B, R = ts->R, M = B * R, N = ts->N, Tb = ts->Tb, F = c.att_feat_size;
  const int nb_layers = std::max(1, c.n_len);
  // memory K/V projections: g(kv weights) += dkv^T . mem ; d mem += dkv . Wkv   (accumulated in fp32)
  RC_TRY(ts->dmem.reserve((size_t)std::max(N * Tb, M) * kD * 4));
  float* dmem = ts->dmem.as<float>();
  CU_TRY(cudaMemsetAsync(dmem, 0, (size_t)M * kD * 4, s));
  for (int l = first_kv; l < nb_layers + c.n_dec; ++l) {
    const Lin& w = (l < nb_layers) ? e->lp[l].ca.kv : e->dec[l - nb_layers].ca.kv;
    const T* dkv = ts->dkv.as<T>() + (size_t)l * M * 2 * kD;
    RC_TRY(lin_bwd<T>(e, s, ts, w, (const T*)ts->memT, kD, dkv, 2 * kD, M, g1, kD));
    // d mem (fp32) += g1
    e->launches++;
    if constexpr (std::is_same<T, float>::value) {
      launch_k(add_inplace_kernel, 148 * 8, 256, 0, s, dmem, (const float*)g1, (size_t)M * kD / 4);
    } else {
      launch_k(sum_over_passes_kernel<T, float>, ceil_div(M, 8), 256, 0, s, (const T*)g1, 1, dmem, (size_t)kD, M, 1);
    }
    CU_TRY(cudaGetLastError());
  }
  // Every gradient of the decoders, embeddings, generator and bounding head (flat entries from "model.decoder..." on) is
  // final here; what follows only touches the encoder / att_embed entries at the front of the flat buffer.  A data-parallel
  // caller starts the all-reduce of the back part on a side stream now, under the encoder's backward pass.
  if (ts->grad_event) CU_TRY(cudaEventRecord(ts->grad_event, s));
  // encoder: final LayerNorm, layers, att_embed
  RC_TRY((ln_bwd<float, T>(e, s, ts, e->enc_norm, ts->enc_x_final, dmem, nullptr, dx, dxT, M)));
  const int* len_dev = ts->att_len;
  for (int l = c.n_enc - 1; l >= 0; --l) {
    RC_TRY(t_layer_bwd<T>(e, s, ts, e->enc[l], ts->enc[l], dx, dxT, g1, g2, B, R, len_dev, 1, 0, (const T*)nullptr, (T*)nullptr, 0, 0, nullptr, 1,
                          false));
    // the gradients of encoder layer l (and, for the last layer, of model.encoder.norm) are final: its all-reduce may start
    if ((size_t)l < ts->layer_events.size() && ts->layer_events[l]) CU_TRY(cudaEventRecord(ts->layer_events[l], s));
  }
  // att_embed: x0 = relu(att . W^T + b), zero on padded rows (their relu mask is false as x0 == 0 there)
  e->launches++;
  launch_k(relu_bwd_cast_kernel<T>, 148 * 8, 256, 0, s, (const float*)dx, (const float*)ts->x0, g1, (size_t)M * kD / 4, ts->d_att_embed.scale);
  CU_TRY(cudaGetLastError());
  RC_TRY(lin_bwd<T>(e, s, ts, e->att_embed, (const T*)ts->attT, F, g1, kD, M, (T*)nullptr, 0));
  return BOFI_OK;
}
