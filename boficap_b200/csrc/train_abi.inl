// C ABI of the training path (declared in include/bofi_b200.h), included at the end of engine.cu.

extern "C" int64_t bofi_param_numel(bofi_handle_t e) { return e ? e->flat_numel : -1; }

extern "C" int bofi_param_offset(bofi_handle_t e, const char* name, int64_t* offset, int64_t* numel) {
  if (!e || !name || !offset || !numel) return fail(BOFI_ERR_INVALID, "null argument");
  auto it = e->weights.find(name);
  if (it == e->weights.end()) return fail(BOFI_ERR_INVALID, "unknown state_dict key '%s'", name);
  *offset = it->second.off;
  *numel = it->second.numel;
  return BOFI_OK;
}

extern "C" int bofi_train_bind(bofi_handle_t e, void* stream, float* flat_params, float* flat_grads) {
  if (!e || !flat_params || !flat_grads) return fail(BOFI_ERR_INVALID, "null argument");
  if (e->cfg.n_len < 1) return fail(BOFI_ERR_INVALID, "the training path is built for N_len >= 1 (uic_sd.yml: 1, uic_sd_N2.yml: 2); got N_len = %d", e->cfg.n_len);
  for (const std::string& n : e->order)
    if (!e->weights[n].loaded) return fail(BOFI_ERR_STATE, "missing state_dict key '%s'", n.c_str());
  CU_TRY(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (flat_params != e->flat_w) {
    CU_TRY(cudaMemcpyAsync(flat_params, e->flat_w, (size_t)e->flat_numel * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CU_TRY(cudaStreamSynchronize(s));
    // captured decode graphs bake in pointers derived from flat_w (biases, LayerNorm gains, fp32 weights): drop them
    // before the buffer they point into is freed (the graph keys carry flat_w / flat16 as well)
    if (e->g_bound.exec) { cudaGraphExecDestroy(e->g_bound.exec); e->g_bound = bofi_engine::GraphSlot(); }
    if (e->g_saic.exec) { cudaGraphExecDestroy(e->g_saic.exec); e->g_saic = bofi_engine::GraphSlot(); }
    e->flat_own.release();
    e->flat_w = flat_params;
  }
  e->flat_g = flat_grads;
  RC_TRY(refresh_derived(e, s));
  CU_TRY(cudaStreamSynchronize(s));
  e->finalized = true;
  return BOFI_OK;
}

extern "C" int bofi_refresh_weights(bofi_handle_t e, void* stream) {
  if (!e) return fail(BOFI_ERR_INVALID, "null handle");
  CU_TRY(cudaSetDevice(e->device));
  RC_TRY(refresh_derived(e, (cudaStream_t)stream));
  e->finalized = true;
  return BOFI_OK;
}

static int train_begin(bofi_handle_t e, cudaStream_t s, int32_t B, int32_t R, int32_t spi, int32_t Lt, int32_t P, const int32_t* labels,
                       const int32_t* phrase_num, const int32_t* phrase_length, const int32_t* phrase_syn, const int32_t* ext_syn,
                       const int32_t* ext_seq, const int32_t* sa_vis) {
  if (!e->finalized) return fail(BOFI_ERR_STATE, "weights not finalised");
  if (e->cfg.n_len < 1) return fail(BOFI_ERR_INVALID, "the training path is built for N_len >= 1");
  if (B <= 0 || R <= 0 || R > kMaxKeys || spi <= 0 || Lt < 1 || Lt + 2 > 32 || P < 1 || P > Lt + 1)
    return fail(BOFI_ERR_INVALID, "bad training batch B=%d R=%d seq_per_img=%d L=%d P=%d", B, R, spi, Lt, P);
  if (!labels || !phrase_num || !phrase_length || !ext_syn || !ext_seq || !sa_vis) return fail(BOFI_ERR_INVALID, "null argument");
  TrainState* ts = train_state(e);
  ts->valid = false;
  ts->arena.reset();
  ts->B = B; ts->R = R; ts->spi = spi; ts->N = B * spi; ts->T = Lt; ts->Tb = Lt + 2; ts->P = P; ts->Mb = ts->N * P;
  const int N = ts->N, Tb = ts->Tb;
  RC_TRY(ts->zeros.reserve(16384 * 4));
  CU_TRY(cudaMemsetAsync(ts->zeros.p, 0, 16384 * 4, s));
  auto copy_i32 = [&](int*& dst, const int32_t* src, size_t n) -> int {
    dst = aalloc<int>(ts, n);
    A_TRY(dst);
    CU_TRY(cudaMemcpyAsync(dst, src, n * 4, cudaMemcpyDeviceToDevice, s));
    return BOFI_OK;
  };
  RC_TRY(copy_i32(ts->labels, labels, (size_t)N * Tb));
  RC_TRY(copy_i32(ts->pnum, phrase_num, (size_t)N));
  RC_TRY(copy_i32(ts->plen, phrase_length, (size_t)N * Tb));
  ts->psyn = nullptr;
  if (phrase_syn) RC_TRY(copy_i32(ts->psyn, phrase_syn, (size_t)N * Tb));
  RC_TRY(copy_i32(ts->ext_syn, ext_syn, (size_t)N * Tb));
  RC_TRY(copy_i32(ts->ext_seq, ext_seq, (size_t)N * Lt));
  RC_TRY(copy_i32(ts->sa_vis, sa_vis, (size_t)N * Lt));
  ts->word_seq = aalloc<int>(ts, (size_t)N * Tb); A_TRY(ts->word_seq);
  ts->vis_b = aalloc<int>(ts, (size_t)N * P); A_TRY(ts->vis_b);
  ts->na_vis = aalloc<int>(ts, (size_t)N); A_TRY(ts->na_vis);
  ts->n_words = aalloc<int>(ts, (size_t)N); A_TRY(ts->n_words);
  ts->total_words = aalloc<int>(ts, 64); A_TRY(ts->total_words);
  return BOFI_OK;
}

extern "C" int bofi_train_forward(bofi_handle_t e, void* stream, const float* att_feats, const int32_t* att_len, int32_t B, int32_t R,
                       int32_t seq_per_img, int32_t Lt, int32_t P, const int32_t* labels, const int32_t* phrase_num,
                       const int32_t* phrase_length, const int32_t* ext_syn, const int32_t* ext_seq, const int32_t* sa_vis,
                       float* sa_len_logp, float* sa_syn_logp, float* sa_logp, float* na_len_logp, float* na_syn_logp, float* na_logp) {
  if (!e || !att_feats || !sa_len_logp || !sa_syn_logp || !sa_logp || !na_len_logp || !na_syn_logp || !na_logp)
    return fail(BOFI_ERR_INVALID, "null argument");
  CU_TRY(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  e->launches = 0;
  RC_TRY(train_begin(e, s, B, R, seq_per_img, Lt, P, labels, phrase_num, phrase_length, nullptr, ext_syn, ext_seq, sa_vis));
  TrainState* ts = train_state(e);
  int rc = e->bf16_mode ? train_forward_impl<bf16>(e, s, ts, att_feats, att_len, sa_len_logp, sa_syn_logp, sa_logp, na_len_logp, na_syn_logp, na_logp, false)
                        : train_forward_impl<float>(e, s, ts, att_feats, att_len, sa_len_logp, sa_syn_logp, sa_logp, na_len_logp, na_syn_logp, na_logp, false);
  ts->valid = (rc == BOFI_OK);
  return rc;
}

template <typename T>
static int train_backward_dense(bofi_engine* e, cudaStream_t s, TrainState* ts, const float* g_sa_len, const float* g_sa_syn, const float* g_sa_logp,
                                const float* g_na_len, const float* g_na_syn, const float* g_na_logp, const float* sa_len, const float* sa_syn,
                                const float* sa_logp, const float* na_len, const float* na_syn, const float* na_logp) {
  const int rows = ts->N * ts->T, ldz = round_up(e->V, 64);
  RC_TRY(e->logits.reserve((size_t)2 * rows * ldz * sizeof(T)));
  T* dz_sa = e->logits.as<T>();
  T* dz_na = dz_sa + (size_t)rows * ldz;
  e->launches += 2;
  launch_k(logsoftmax_bwd_kernel<T>, rows, 512, 0, s, g_sa_logp, sa_logp, e->V, dz_sa, ldz);
  CU_TRY(cudaGetLastError());
  launch_k(logsoftmax_bwd_kernel<T>, rows, 512, 0, s, g_na_logp, na_logp, e->V, dz_na, ldz);
  CU_TRY(cudaGetLastError());
  return train_backward_impl<T>(e, s, ts, dz_sa, dz_na, ldz, g_sa_len, g_sa_syn, g_na_len, g_na_syn, sa_len, sa_syn, na_len, na_syn);
}

extern "C" int bofi_train_backward(bofi_handle_t e, void* stream, const float* g_sa_len, const float* g_sa_syn, const float* g_sa_logp,
                        const float* g_na_len, const float* g_na_syn, const float* g_na_logp, const float* sa_len_logp,
                        const float* sa_syn_logp, const float* sa_logp, const float* na_len_logp, const float* na_syn_logp,
                        const float* na_logp) {
  if (!e || !g_sa_len || !g_sa_syn || !g_sa_logp || !g_na_len || !g_na_syn || !g_na_logp || !sa_len_logp || !sa_syn_logp || !sa_logp ||
      !na_len_logp || !na_syn_logp || !na_logp)
    return fail(BOFI_ERR_INVALID, "null argument");
  if (!e->flat_g) return fail(BOFI_ERR_STATE, "bofi_train_backward needs bofi_train_bind (gradient buffer)");
  TrainState* ts = train_state(e);
  if (!ts->valid) return fail(BOFI_ERR_STATE, "bofi_train_backward needs a preceding bofi_train_forward");
  CU_TRY(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  int rc = e->bf16_mode ? train_backward_dense<bf16>(e, s, ts, g_sa_len, g_sa_syn, g_sa_logp, g_na_len, g_na_syn, g_na_logp, sa_len_logp, sa_syn_logp,
                                                     sa_logp, na_len_logp, na_syn_logp, na_logp)
                        : train_backward_dense<float>(e, s, ts, g_sa_len, g_sa_syn, g_sa_logp, g_na_len, g_na_syn, g_na_logp, sa_len_logp, sa_syn_logp,
                                                      sa_logp, na_len_logp, na_syn_logp, na_logp);
  ts->valid = false;          // the tape is consumed
  return rc;
}

// Forward + LanguageModelCriterion_UIC (reduction='mean') + backward in one call; the [N, L, V] log-prob tensors and
// their gradients are never materialised (the criterion's gradient is formed straight from the logits).
template <typename T>
static int train_step_impl(bofi_engine* e, cudaStream_t s, TrainState* ts, const float* att, const int* att_len, float* losses) {
  const int N = ts->N, T_ = ts->T, Tb = ts->Tb, rows = N * T_, slots = N * (Tb - 1), ldz = round_up(e->V, 64);
  // small outputs + loss scratch: [sa_len, sa_syn, na_len, na_syn] log-probs, their gradients, per-row NLLs, 6 sums
  float* sa_len = aalloc<float>(ts, (size_t)slots * 20); A_TRY(sa_len);
  float* sa_syn = aalloc<float>(ts, (size_t)slots * 10); A_TRY(sa_syn);
  float* na_len = aalloc<float>(ts, (size_t)slots * 20); A_TRY(na_len);
  float* na_syn = aalloc<float>(ts, (size_t)slots * 10); A_TRY(na_syn);
  float* g_sa_len = aalloc<float>(ts, (size_t)slots * 20); A_TRY(g_sa_len);
  float* g_sa_syn = aalloc<float>(ts, (size_t)slots * 10); A_TRY(g_sa_syn);
  float* g_na_len = aalloc<float>(ts, (size_t)slots * 20); A_TRY(g_na_len);
  float* g_na_syn = aalloc<float>(ts, (size_t)slots * 10); A_TRY(g_na_syn);
  float* nll = aalloc<float>(ts, (size_t)2 * rows + 4 * slots); A_TRY(nll);
  float* sums = aalloc<float>(ts, 64); A_TRY(sums);
  CU_TRY(cudaMemsetAsync(ts->total_words, 0, 4, s));
  launch_k(xe_count_words_kernel, ceil_div(N, 128), 128, 0, s, (const int*)ts->plen, N, Tb, ts->n_words, ts->total_words);
  CU_TRY(cudaGetLastError());
  RC_TRY(train_forward_impl<T>(e, s, ts, att, att_len, sa_len, sa_syn, nullptr, na_len, na_syn, nullptr, true));
  RC_TRY(e->logits.reserve((size_t)2 * rows * ldz * sizeof(T)));
  T* dz_sa = e->logits.as<T>();
  T* dz_na = dz_sa + (size_t)rows * ldz;
  float* nll_sa = nll, *nll_na = nll + rows, *nl_sa_len = nll + 2 * rows, *nl_sa_syn = nl_sa_len + slots, *nl_na_len = nl_sa_syn + slots,
        *nl_na_syn = nl_na_len + slots;
  e->launches += 11;
  launch_k(xe_word_loss_bwd_kernel<T>, rows, 512, 0, s, (const float*)ts->sa_d.logits, e->Vpad, e->V, (const int*)ts->labels, Tb, 1, (const int*)ts->n_words,
           T_, (const int*)ts->total_words, dz_sa, ldz, nll_sa);
  CU_TRY(cudaGetLastError());
  launch_k(xe_word_loss_bwd_kernel<T>, rows, 512, 0, s, (const float*)ts->na_d.logits, e->Vpad, e->V, (const int*)ts->labels, Tb, 1, (const int*)ts->n_words,
           T_, (const int*)ts->total_words, dz_na, ldz, nll_na);
  CU_TRY(cudaGetLastError());
  launch_k(xe_box_loss_grad_kernel, ceil_div(slots, 128), 128, 0, s, (const float*)sa_len, (const float*)sa_syn, (const int*)ts->pnum, (const int*)ts->plen,
           (const int*)ts->psyn, N, Tb, 20, 10, (const int*)ts->total_words, g_sa_len, g_sa_syn, nl_sa_len, nl_sa_syn);
  CU_TRY(cudaGetLastError());
  launch_k(xe_box_loss_grad_kernel, ceil_div(slots, 128), 128, 0, s, (const float*)na_len, (const float*)na_syn, (const int*)ts->pnum, (const int*)ts->plen,
           (const int*)ts->psyn, N, Tb, 20, 10, (const int*)ts->total_words, g_na_len, g_na_syn, nl_na_len, nl_na_syn);
  CU_TRY(cudaGetLastError());
  // reference order: SA_length, SA_phrase, SA_syn, NA_length, NA_phrase, NA_syn
  const float* parts[6] = {nl_sa_len, nll_sa, nl_sa_syn, nl_na_len, nll_na, nl_na_syn};
  const int counts[6] = {slots, rows, slots, slots, rows, slots};
  for (int i = 0; i < 6; ++i) {
    launch_k(sum_kernel, 1, 1024, 0, s, parts[i], counts[i], sums + i, 1.0f);
    CU_TRY(cudaGetLastError());
  }
  launch_k(xe_finish_losses_kernel, 1, 32, 0, s, (const float*)sums, (const int*)ts->total_words, losses);
  CU_TRY(cudaGetLastError());
  return train_backward_impl<T>(e, s, ts, dz_sa, dz_na, ldz, g_sa_len, g_sa_syn, g_na_len, g_na_syn, sa_len, sa_syn, na_len, na_syn);
}

extern "C" int bofi_train_step_xe(bofi_handle_t e, void* stream, const float* att_feats, const int32_t* att_len, int32_t B, int32_t R,
                       int32_t seq_per_img, int32_t Lt, int32_t P, const int32_t* labels, const int32_t* phrase_num,
                       const int32_t* phrase_length, const int32_t* phrase_syn, const int32_t* ext_syn, const int32_t* ext_seq,
                       const int32_t* sa_vis, float* losses) {
  if (!e || !att_feats || !losses || !phrase_syn) return fail(BOFI_ERR_INVALID, "null argument");
  if (!e->flat_g) return fail(BOFI_ERR_STATE, "bofi_train_step_xe needs bofi_train_bind (gradient buffer)");
  CU_TRY(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  e->launches = 0;
  RC_TRY(train_begin(e, s, B, R, seq_per_img, Lt, P, labels, phrase_num, phrase_length, phrase_syn, ext_syn, ext_seq, sa_vis));
  TrainState* ts = train_state(e);
  int rc = e->bf16_mode ? train_step_impl<bf16>(e, s, ts, att_feats, att_len, losses) : train_step_impl<float>(e, s, ts, att_feats, att_len, losses);
  ts->valid = false;
  return rc;
}

// ---- self-critical training: sampling with a tape ----------------------------------------------
// loss_wrapper.py:194-214 calls model(..., opt={'sample_method': 'sample', 'sample_n': n, 'train_mode': 'SAIC' | 'NAIC'},
// mode='sample') in train() mode and back-propagates StructureLosses (losses.py:157-176) through the returned seq_logprobs.
// What is differentiable in the reference is exactly one decoder pass on fixed inputs: the boxes come from an argmax and the
// words from a multinomial draw.  So:
//   1. encoder with tape and dropout (t_encode_fwd), its memory handed to the decode path;
//   2. NAIC: the bounding loop (decode kernels) fixes the syn label of every slot and the fill window; the taped NA decoder
//      pass gives the log-probs the tokens are drawn from (Gumbel-max in the vocabulary epilogue);
//      SAIC: the incremental decode draws words and boxes phrase by phrase; the taped SA decoder pass then replays all
//      phrases at once on those words (hidden states of earlier phrases never change under the phrase-block-causal mask),
//      slots that were never committed return zeros like the reference's zero-initialised seq_logprobs;
//   3. bofi_sc_backward: d logits from d seq_logprobs, decoder, memory K/V projections, encoder, att_embed.
// Deviations, both on the non-differentiable side: the bounding layer and (SAIC) the word draws run with the eval()
// arithmetic (no dropout), the taped pass and the encoder with the train() dropout of bofi_train_set_dropout.
template <typename T>
static int sc_sample_impl(bofi_engine* e, cudaStream_t s, TrainState* ts, const float* att, const int* att_len, int mode, long long* seq,
                          float* logprobs, int* phrase_num, int* phrase_length, long long* phrase_syn) {
  const bofi_config_t& c = e->cfg;
  const int B = ts->B, R = ts->R, M = B * R, sn = ts->spi, N = ts->N, L = e->L, Lb = e->Lb;
  ts->site = 0;
  RC_TRY(t_encode_fwd<T>(e, s, ts, att, att_len));
  // the decode path reads the memory from the engine workspace (padded layout, eval arithmetic)
  RC_TRY(reserve_encode(e, B, R));
  CU_TRY(cudaMemcpyAsync(e->memT.p, ts->memT, (size_t)M * kD * sizeof(T), cudaMemcpyDeviceToDevice, s));
  e->have_len = (att_len != nullptr);
  if (att_len) CU_TRY(cudaMemcpyAsync(e->attlen.p, att_len, (size_t)B * 4, cudaMemcpyDeviceToDevice, s));
  e->enc_off = e->mem_off = nullptr;
  e->B = B;
  e->R = R;
  e->have_memory = true;
  RC_TRY(reserve_decode(e, B, R, sn));
  ts->sc_words = nullptr;
  ts->sc_vis = aalloc<int>(ts, (size_t)N * L); A_TRY(ts->sc_vis);
  ts->sc_total = aalloc<int>(ts, (size_t)N); A_TRY(ts->sc_total);
  ts->ext_syn = aalloc<int>(ts, (size_t)N * Lb); A_TRY(ts->ext_syn);
  const int* src_syn;
  if (mode == BOFI_MODE_NAIC) {
    RC_TRY(naic_bound_phase<T>(e, s, sn));
    src_syn = e->st.ext;
  } else {
    ts->sc_words = aalloc<int>(ts, (size_t)N * L); A_TRY(ts->sc_words);
    RC_TRY(decode_saic_incremental<T>(e, s, sn, 1, seq, nullptr, phrase_num, phrase_length, phrase_syn));
    CU_TRY(cudaMemcpy2DAsync(ts->sc_words, (size_t)L * 4, e->st.ext_word + 1, (size_t)Lb * 4, (size_t)L * 4, N, cudaMemcpyDeviceToDevice, s));
    src_syn = e->st.ext_syn;
  }
  CU_TRY(cudaMemcpyAsync(ts->ext_syn, src_syn, (size_t)N * Lb * 4, cudaMemcpyDeviceToDevice, s));
  CU_TRY(cudaMemcpyAsync(ts->sc_vis, e->st.vis_fill, (size_t)N * L * 4, cudaMemcpyDeviceToDevice, s));
  CU_TRY(cudaMemcpyAsync(ts->sc_total, e->st.last, (size_t)N * 4, cudaMemcpyDeviceToDevice, s));
  if (mode == BOFI_MODE_SAIC) {
    // a row without any phrase keeps an all-False phrase mask: the reference aborts the whole batch there ("phrase nan!",
    // :1956-1958).  Its slots return zeros and carry no gradient here either, but their activations must stay finite
    // (0 * NaN would poison the weight gradients), so such rows attend to their first slot.
    launch_k(clamp_min_i32_kernel, ceil_div(N * L, 256), 256, 0, s, ts->sc_vis, 1, N * L);
    CU_TRY(cudaGetLastError());
  }
  // the taped decoder pass: NAIC draws the tokens from it; SAIC only scores the words already drawn
  if (mode == BOFI_MODE_NAIC) {
    RC_TRY(t_dec_fwd<T>(e, s, ts, ts->na_d, nullptr, c.bos_idx, ts->sc_vis, L, 1, logprobs, false, e->sampler, seq, ts->sc_total));
    launch_k(export_boxes_kernel, ceil_div(N * L, 256), 256, 0, s, e->st, N, Lb, L, 0, phrase_num, phrase_length, phrase_syn);
    CU_TRY(cudaGetLastError());
  } else {
    RC_TRY(t_dec_fwd<T>(e, s, ts, ts->na_d, ts->sc_words, -1, ts->sc_vis, L, 1, logprobs, false));
    launch_k(zero_tail_rows_kernel, ceil_div((size_t)N * L, 8), 256, 0, s, logprobs, (const int*)ts->sc_total, -1, N * L, L, e->V);
    CU_TRY(cudaGetLastError());
  }
  ts->sc = true;
  ts->sc_mode = mode;
  return BOFI_OK;
}

extern "C" int bofi_sc_sample(bofi_handle_t e, void* stream, int32_t mode, int32_t sample_n, const float* att_feats, const int32_t* att_len,
                              int32_t B, int32_t R, int64_t* seq, float* logprobs, int32_t* phrase_num, int32_t* phrase_length,
                              int64_t* phrase_syn) {
  if (!e || !att_feats || !seq || !logprobs || !phrase_num || !phrase_length || !phrase_syn) return fail(BOFI_ERR_INVALID, "null argument");
  if (!e->finalized) return fail(BOFI_ERR_STATE, "weights not finalised");
  if (!e->flat_g) return fail(BOFI_ERR_STATE, "bofi_sc_sample needs bofi_train_bind (gradient buffer)");
  if (e->cfg.n_len < 1) return fail(BOFI_ERR_INVALID, "the training path is built for N_len >= 1");
  if (mode != BOFI_MODE_NAIC && mode != BOFI_MODE_SAIC) return fail(BOFI_ERR_INVALID, "mode %d (NAIC = 0, SAIC = 1)", mode);
  if (B <= 0 || R <= 0 || R > kMaxKeys || sample_n < 1) return fail(BOFI_ERR_INVALID, "bad batch B=%d R=%d sample_n=%d", B, R, sample_n);
  CU_TRY(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  e->launches = 0;
  TrainState* ts = train_state(e);
  ts->valid = false;
  ts->arena.reset();
  ts->B = B; ts->R = R; ts->spi = sample_n; ts->N = B * sample_n; ts->T = e->L; ts->Tb = e->Lb; ts->P = 1; ts->Mb = ts->N;
  RC_TRY(ts->zeros.reserve(16384 * 4));
  CU_TRY(cudaMemsetAsync(ts->zeros.p, 0, 16384 * 4, s));
  int rc = e->bf16_mode ? sc_sample_impl<bf16>(e, s, ts, att_feats, att_len, mode, (long long*)seq, logprobs, phrase_num, phrase_length, (long long*)phrase_syn)
                        : sc_sample_impl<float>(e, s, ts, att_feats, att_len, mode, (long long*)seq, logprobs, phrase_num, phrase_length, (long long*)phrase_syn);
  ts->valid = (rc == BOFI_OK);
  return rc;
}

template <typename T>
static int sc_backward_impl(bofi_engine* e, cudaStream_t s, TrainState* ts, const float* g_logp, const float* logp) {
  const bofi_config_t& c = e->cfg;
  const int N = ts->N, T_ = ts->T, rows = N * T_, M = ts->B * ts->R, ldz = round_up(e->V, 64);
  const int nb_layers = std::max(1, c.n_len);
  RC_TRY(e->logits.reserve((size_t)rows * ldz * sizeof(T)));
  T* dz = e->logits.as<T>();
  e->launches += 1;
  launch_k(logsoftmax_bwd_kernel<T>, rows, 512, 0, s, g_logp, logp, e->V, dz, ldz);
  CU_TRY(cudaGetLastError());
  if (ts->sc_mode == BOFI_MODE_SAIC) {      // the zero rows of never-committed slots are constants
    launch_k(zero_tail_rows_kernel_t<T>, ceil_div((size_t)rows, 8), 256, 0, s, dz, (const int*)ts->sc_total, -1, rows, T_, ldz);
    CU_TRY(cudaGetLastError());
  }
  const size_t big = std::max((size_t)rows, (size_t)M);
  RC_TRY(e->x.reserve(big * kD * 4));
  RC_TRY(e->y.reserve(big * kD * sizeof(T)));
  RC_TRY(e->q.reserve(big * kD * sizeof(T)));
  RC_TRY(e->ffh.reserve(big * std::max(c.d_ff, 3 * kD) * sizeof(T)));
  RC_TRY(ts->dkv.reserve((size_t)(nb_layers + c.n_dec) * M * 2 * kD * sizeof(T)));
  float* dx = e->x.as<float>();
  T* dxT = e->y.as<T>();
  T* g1 = e->q.as<T>();
  T* g2 = e->ffh.as<T>();
  const bool naic = ts->sc_mode == BOFI_MODE_NAIC;
  RC_TRY(t_dec_bwd<T>(e, s, ts, ts->na_d, dz, ldz, naic ? nullptr : ts->sc_words, naic, ts->sc_vis, T_, 1, dx, dxT, g1, g2, 1));
  return t_encode_bwd<T>(e, s, ts, dx, dxT, g1, g2, nb_layers);      // the bounding layer's memory K/V carry no gradient here
}

extern "C" int bofi_sc_backward(bofi_handle_t e, void* stream, const float* g_logprobs, const float* logprobs) {
  if (!e || !g_logprobs || !logprobs) return fail(BOFI_ERR_INVALID, "null argument");
  if (!e->flat_g) return fail(BOFI_ERR_STATE, "bofi_sc_backward needs bofi_train_bind (gradient buffer)");
  TrainState* ts = train_state(e);
  if (!ts->valid || !ts->sc) return fail(BOFI_ERR_STATE, "bofi_sc_backward needs a preceding bofi_sc_sample");
  CU_TRY(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  int rc = e->bf16_mode ? sc_backward_impl<bf16>(e, s, ts, g_logprobs, logprobs) : sc_backward_impl<float>(e, s, ts, g_logprobs, logprobs);
  ts->valid = false;
  return rc;
}

// The decoder inputs of the taped pass of the last bofi_sc_sample (tests / inspection): word ids (NAIC: bos everywhere),
// syn ids and visible-key counts per (row, slot), i32 [N, L] each, and the committed word count + 1 per row, i32 [N].
extern "C" int bofi_sc_inputs(bofi_handle_t e, void* stream, int32_t* word_ids, int32_t* syn_ids, int32_t* vis, int32_t* total) {
  if (!e || !word_ids || !syn_ids || !vis || !total) return fail(BOFI_ERR_INVALID, "null argument");
  TrainState* ts = train_state(e);
  if (!ts->sc) return fail(BOFI_ERR_STATE, "bofi_sc_inputs needs a preceding bofi_sc_sample");
  CU_TRY(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int N = ts->N, L = ts->T, Lb = ts->Tb;
  if (ts->sc_words) CU_TRY(cudaMemcpyAsync(word_ids, ts->sc_words, (size_t)N * L * 4, cudaMemcpyDeviceToDevice, s));
  else {
    launch_k(fill_i32_kernel, ceil_div(N * L, 256), 256, 0, s, word_ids, e->cfg.bos_idx, N * L);
    CU_TRY(cudaGetLastError());
  }
  CU_TRY(cudaMemcpy2DAsync(syn_ids, (size_t)L * 4, ts->ext_syn + 1, (size_t)Lb * 4, (size_t)L * 4, N, cudaMemcpyDeviceToDevice, s));
  CU_TRY(cudaMemcpyAsync(vis, ts->sc_vis, (size_t)N * L * 4, cudaMemcpyDeviceToDevice, s));
  CU_TRY(cudaMemcpyAsync(total, ts->sc_total, (size_t)N * 4, cudaMemcpyDeviceToDevice, s));
  return BOFI_OK;
}

extern "C" int bofi_train_set_dropout(bofi_handle_t e, float p, float p_att_embed, uint32_t seed) {
  if (!e) return fail(BOFI_ERR_INVALID, "null handle");
  if (p < 0.f || p >= 1.f || p_att_embed < 0.f || p_att_embed >= 1.f) return fail(BOFI_ERR_INVALID, "dropout probabilities must be in [0, 1)");
  TrainState* ts = train_state(e);
  ts->p_sub = p;
  ts->p_att = p_att_embed;
  ts->seed = seed;
  return BOFI_OK;
}

extern "C" int bofi_train_set_glat(bofi_handle_t e, float glat_p, uint32_t seed) {
  if (!e) return fail(BOFI_ERR_INVALID, "null handle");
  if (glat_p > 1.f) return fail(BOFI_ERR_INVALID, "glat_p %g (an unmasked-token ratio in [0, 1], < 0 = off)", glat_p);
  TrainState* ts = train_state(e);
  ts->glat_p = glat_p;
  ts->glat_seed = seed;
  return BOFI_OK;
}

extern "C" int bofi_train_set_grad_event(bofi_handle_t e, void* event) {
  if (!e) return fail(BOFI_ERR_INVALID, "null handle");
  train_state(e)->grad_event = (cudaEvent_t)event;
  return BOFI_OK;
}

extern "C" int bofi_train_set_layer_event(bofi_handle_t e, int32_t layer, void* event) {
  if (!e) return fail(BOFI_ERR_INVALID, "null handle");
  if (layer < 0 || layer >= e->cfg.n_enc) return fail(BOFI_ERR_INVALID, "encoder layer %d of %d", layer, e->cfg.n_enc);
  TrainState* ts = train_state(e);
  if (ts->layer_events.size() < (size_t)e->cfg.n_enc) ts->layer_events.assign((size_t)e->cfg.n_enc, nullptr);
  ts->layer_events[layer] = (cudaEvent_t)event;
  return BOFI_OK;
}

extern "C" int bofi_train_launches(bofi_handle_t e) { return e ? e->launches : -1; }
