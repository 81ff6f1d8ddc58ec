/*
 * bofi_b200.h -- C ABI of the B200-native BoFiCap hot path (libbofi_b200.so).
 *
 * The reference (ChangxinWang/BoFiCap) has no FFI / plugin interface: its boundary for this path is
 * the Python model class.  This header is the boundary a binding would target; every entry point
 * names the reference code it replaces (paths relative to /root/reference/captioning/models/).
 * The Python drop-in (boficap_b200/captioning/models) binds it with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all sizes explicit; int return codes (0 = BOFI_OK);
 *     no exception crosses the ABI; bofi_last_error() returns the text of the last failure
 *     on the calling thread.
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every *device*
 *     entry point is asynchronous on that stream and never synchronises; the *_host entry points
 *     copy H2D / D2H themselves and return after the results are in the host buffers.
 *   - the library owns only its packed weights and its workspace (grown on demand, freed by
 *     bofi_destroy); callers own every buffer they pass.
 *   - one handle per device; a handle is not thread-safe, distinct handles are independent.
 *   - there is no CPU fallback: every call fails with BOFI_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef BOFI_B200_H_
#define BOFI_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOFI_ABI_VERSION 2

enum {
  BOFI_OK = 0,
  BOFI_ERR_INVALID = 1,   /* bad argument / unsupported configuration */
  BOFI_ERR_CUDA = 2,      /* CUDA runtime / driver error              */
  BOFI_ERR_STATE = 3,     /* call order (weights missing, not finalised, ...) */
  BOFI_ERR_NOMEM = 4
};

enum { BOFI_PRECISION_FP32 = 0, BOFI_PRECISION_BF16 = 1 };
enum { BOFI_MODE_NAIC = 0, BOFI_MODE_SAIC = 1 };
/* element type of a region-feature buffer handed to the *_ex entry points (the loader's `att_feats`, dataloader.py:24-86):
 * fp32 as the reference's loader yields it, or the 2-byte forms of the pinned feeder (boficap_b200/data). */
enum { BOFI_FEAT_F32 = 0, BOFI_FEAT_BF16 = 1, BOFI_FEAT_F16 = 2 };

/* Hyper-parameters the reference reads from `opt`
 * (AttModel.py:56-75, TransformerModel.py:1631-1640, :404-411).  tgt_vocab = vocab_size + 4. */
typedef struct bofi_config {
  int32_t abi_version;    /* BOFI_ABI_VERSION */
  int32_t tgt_vocab;      /* V, 9491 for uic_sd */
  int32_t att_feat_size;  /* 2048 */
  int32_t n_enc, n_dec, n_len;
  int32_t d_model, d_ff, heads;
  int32_t seq_length;     /* L = 20; bounding slots Lb = L + 2 */
  int32_t pad_idx, bos_idx, eos_idx, len_idx;
  int32_t precision;      /* BOFI_PRECISION_* : arithmetic of the dense contractions */
} bofi_config_t;

typedef struct bofi_engine* bofi_handle_t;

/* Diagnostics of the last decode on a handle (device-resident counters read back on request). */
typedef struct bofi_decode_info {
  int32_t bounding_steps;   /* S: bounding iterations that did work (core_NAIC loop, :1833-1870) */
  int32_t fill_width;       /* w = last[B-1]-1, the stale-index key window (:1871-1873)         */
  int32_t kernel_launches;  /* kernels this library enqueued for the decode                     */
  int32_t nan_batch;        /* 1 when w == 0 made the whole batch NaN (AttModel.py:427-428)     */
} bofi_decode_info_t;

const char* bofi_last_error(void);
int bofi_abi_version(void);

/* make_model UIC branch + TransformerModel.__init__ (TransformerModel.py:1558-1568, :1626-1666):
 * allocates device storage for the 311-entry state_dict of `cfg` on CUDA device `device`. */
int bofi_create(const bofi_config_t* cfg, int device, bofi_handle_t* out);
int bofi_destroy(bofi_handle_t h);

/* load_state_dict (tools/eval.py:103): one call per state_dict entry, `name` is the reference key
 * (SURVEY.md Appendix B), `data` fp32 host memory with `numel` elements (shape is checked by count).
 * Keys that the decode path never reads (model.length_predictor.length_attn.* / .ff.* when
 * N_len >= 1) are accepted and ignored. */
int bofi_set_weight(bofi_handle_t h, const char* name, const float* host_data, int64_t numel);
/* Number of state_dict entries still missing (0 = complete). */
int bofi_missing_weights(bofi_handle_t h);
/* Packs weights for the device path (bf16 copies, fused QKV / KV matrices, (syn,pos) input tables). */
int bofi_finalize_weights(bofi_handle_t h, void* stream);

/* Workspace the library will hold for a batch of B images x R regions x sample_n (bytes). */
int64_t bofi_workspace_bytes(bofi_handle_t h, int32_t B, int32_t R, int32_t sample_n);

/* _prepare_feature (TransformerModel.py:1674-1711: clip_att, pack_wrapper(att_embed), encode):
 *   att_feats  dev f32 [B,R,att_feat_size]
 *   att_len    dev i32 [B] = number of valid (prefix) regions per image, or NULL when att_masks is None
 *   memory_out dev f32 [B,R,d_model] or NULL (memory stays in the workspace for bofi_decode). */
int bofi_encode(bofi_handle_t h, void* stream, const float* att_feats, const int32_t* att_len,
                int32_t B, int32_t R, float* memory_out);

/* The same with 2-byte features (feat_dtype = BOFI_FEAT_*; att_feats 16-byte aligned).  A bf16 engine reads bf16 features in
 * place (the att_embed GEMM's TMA loads them from the caller's buffer, no conversion pass); every other combination costs
 * one conversion kernel.  Halves the bytes of the input side (147 KB instead of 295 KB per 36-region image). */
int bofi_encode_ex(bofi_handle_t h, void* stream, const void* att_feats, int32_t feat_dtype, const int32_t* att_len,
                   int32_t B, int32_t R, float* memory_out);

/* att_masks -> att_len for bofi_encode (the reduction `att_masks.sum(1)` of a PREFIX mask, dataloader.py:333-338):
 *   att_masks dev f32 [B,R] (0 / non-zero), att_len dev i32 [B].
 * The kernel also checks that every row really is a prefix mask (the reference's masked_fill accepts any mask, this
 * library only counts); a violation raises a device-side flag that bofi_check_masks reads (and clears) -- *bad = 1 --
 * at a point where the caller synchronises anyway.  Neither call is needed when the caller already has the counts. */
int bofi_masks_to_len(bofi_handle_t h, void* stream, const float* att_masks, int32_t B, int32_t R, int32_t* att_len);
int bofi_check_masks(bofi_handle_t h, void* stream, int32_t* bad);

/* core_NAIC / core_SAIC + logit + log_softmax + greedy sample_next_word + tail padding
 * (TransformerModel.py:1823-1986, AttModel.py:203-210, :419-437, CaptionModel.py:383-390),
 * on the memory left in the workspace by the preceding bofi_encode.  Rows = B * sample_n.
 *   seq           dev i64 [rows, L]
 *   logprobs      dev f32 [rows, L, V] or NULL (skip materialising the 759 KB/row tensor)
 *   phrase_num    dev i32 [rows]; phrase_length dev i32 [rows, L]; phrase_syn dev i64 [rows, L]
 *   output_logsoftmax: 1 = log-softmax (default of the reference), 0 = raw logits. */
int bofi_decode(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, int32_t output_logsoftmax,
                int64_t* seq, float* logprobs, int32_t* phrase_num, int32_t* phrase_length,
                int64_t* phrase_syn);

/* bofi_decode with a PITCH for the log-prob rows: logprobs is dev f32 [rows, L, logprob_ld] of which the first V columns of
 * every row are written (logprob_ld = 0 or V: the dense tensor of bofi_decode).  NAIC only.  With a pitch that is a multiple
 * of four floats (e.g. 9492 for V = 9491) every row starts on a 16-byte boundary and the tcgen05 engine writes the tensor with
 * TMA stores instead of 4-byte row-segment stores -- what `seq_logprob[:, :, :V]` of a padded allocation is in PyTorch terms
 * (same values, same indexing, a strided view). */
int bofi_decode_ex(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, int32_t output_logsoftmax,
                   int64_t* seq, float* logprobs, int64_t logprob_ld, int32_t* phrase_num, int32_t* phrase_length,
                   int64_t* phrase_syn);

/* _prepare_feature on COMPACT features: att_compact holds only the valid regions, image after image -- [total_rows, F] with
 * total_rows = sum(att_len) -- which is what pack_wrapper (AttModel.py:33-51) hands to att_embed.
 * att_embed runs on those rows directly (no padded GEMM, no compaction pass) and a host caller moves sum(att_len) instead of B * R
 * rows over PCIe.  Same memory (bit for bit) as bofi_encode_ex on the padded tensor with the same att_len.  Needs att_len.
 * bofi_stage_compact copies a host (or device) compact batch into the library's staging buffer, bofi_encode_staged_compact encodes
 * it; follow with bofi_decode / bofi_decode_host_async. */
int bofi_encode_compact(bofi_handle_t h, void* stream, const void* att_compact, int32_t feat_dtype, const int32_t* att_len,
                        int32_t total_rows, int32_t B, int32_t R);
int bofi_stage_compact(bofi_handle_t h, void* stream, const void* att_compact, int32_t feat_dtype, const int32_t* att_len,
                       int32_t total_rows, int32_t B, int32_t R);
int bofi_encode_staged_compact(bofi_handle_t h, void* stream, int32_t feat_dtype, int32_t total_rows, int32_t B, int32_t R);
/* One batch of several that will share a call (bofi_set_shard): its rows_part compact rows go to rows [row0, ...) of the staging
 * buffer (the compact rows of the batches follow each other), its Bpart counts to images [image0, ...). */
int bofi_stage_compact_part(bofi_handle_t h, void* stream, const void* att_compact, int32_t feat_dtype, const int32_t* att_len,
                            int32_t rows_part, int32_t row0, int32_t image0, int32_t Bpart, int32_t Btotal, int32_t R);

/* Several batches in ONE call (what nn.DataParallel replicas / successive eval batches are to the reference: independent
 * `_sample` calls).  The only place where a row of `_sample` depends on the other rows of its batch is the fill window
 * w = last[B-1] - 1 of core_NAIC (TransformerModel.py:1871-1873).  With bofi_set_shard(h, n) the B images of the following
 * bofi_encode / bofi_decode calls are B / n batches of n images (the last may be shorter) and every batch keeps its own window,
 * so the result is bit for bit what B / n separate calls give -- while the bounding loop's ~190 latency-bound launches are paid
 * once.  n = 0: the whole call is one batch (default).  NAIC; bofi_decode_info reports the last batch's window.
 * bofi_stage_part copies one batch (host or device memory; features [Bpart, R, F] of feat_dtype, optional att_len [Bpart]) to rows
 * [row0, row0 + Bpart) of a library-owned device buffer for Btotal images, bofi_encode_staged runs bofi_encode_ex on that buffer,
 * bofi_sample_staged = bofi_encode_staged + bofi_decode_host_async. */
int bofi_set_shard(bofi_handle_t h, int32_t shard_images);
int bofi_stage_part(bofi_handle_t h, void* stream, const void* att_feats, int32_t feat_dtype, const int32_t* att_len,
                    int32_t row0, int32_t Bpart, int32_t Btotal, int32_t R);
int bofi_encode_staged(bofi_handle_t h, void* stream, int32_t feat_dtype, int32_t have_len, int32_t B, int32_t R);
int bofi_sample_staged(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, int32_t output_logsoftmax,
                       int32_t feat_dtype, int32_t have_len, int32_t B, int32_t R, int64_t* seq, float* logprobs,
                       int32_t* phrase_num, int32_t* phrase_length, int64_t* phrase_syn);
/* bofi_decode + the asynchronous device-to-host copies of the results (pinned host outputs), on the memory of the preceding
 * bofi_encode*.  The staging buffers of bofi_stage_part are free again once bofi_encode_staged has run: a caller that records an
 * event between bofi_encode_staged and this call can stage the next batch on a copy stream underneath this decode. */
int bofi_decode_host_async(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, int32_t output_logsoftmax,
                           int64_t* seq, float* logprobs, int32_t* phrase_num, int32_t* phrase_length, int64_t* phrase_syn);

/* eval_split's entropy / perplexity (captioning/utils/eval_utils.py:183-184) without materialising seq_logprob:
 * when set (dev f32 [rows, L] each; NULL, NULL switches it off), the following bofi_decode calls also write, per slot,
 *   slot_entropy = -sum_v p_v log p_v   and   slot_logp = log p of the token written to seq
 * (zero for slots SAIC never commits, as the reference's zero-initialised seq_logprobs give).  Needs output_logsoftmax = 1.
 *   entropy[b]    =  sum_t slot_entropy[b,t] / (count(seq[b] > 3) + 1)
 *   perplexity[b] = -sum_t slot_logp[b,t]    / (count(seq[b] > 3) + 1)          (boficap_b200/captioning/models: sample_stats) */
int bofi_set_decode_stats(bofi_handle_t h, float* slot_entropy, float* slot_logp);

/* sample_next_word (CaptionModel.py:383-431) for the following bofi_decode calls on this handle:
 *   method 0  greedy: first maximal index (default)
 *   method 1  sample_method='sample': Categorical(logits = logprobs / temperature, NaN -> -10), drawn as a Gumbel-max
 *             with the library's counter-based random stream seeded by `seed` (same distribution as the reference,
 *             not torch's stream).  This is the sampling half of self-critical training (loss_wrapper.py:194-209).
 * gumbel / top-k / nucleus variants (:391-421) are not built. */
int bofi_set_sampling(bofi_handle_t h, int32_t method, float temperature, uint32_t seed);

/* AttModel._sample (AttModel.py:307-334, :419-437) end to end with HOST buffers: H2D of att_feats /
 * att_len, encode, decode, D2H of the results, then a stream synchronise.  host pointers may be
 * pageable or pinned; logprobs may be NULL. */
int bofi_sample_host(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, int32_t output_logsoftmax,
                     const float* att_feats, const int32_t* att_len, int32_t B, int32_t R,
                     int64_t* seq, float* logprobs, int32_t* phrase_num, int32_t* phrase_length,
                     int64_t* phrase_syn);

/* Same work, enqueued only: nothing is synchronised, the results are in the host buffers once `stream` has
 * drained (cudaStreamSynchronize / an event recorded after the call).  Host buffers must be PINNED for the
 * copies to be asynchronous and must stay alive until then.  Two handles on two streams, fed alternately,
 * overlap the H2D copy and the latency-bound bounding loop of one batch with the dense encoder / filling
 * GEMMs of the other (boficap_b200/pipeline.py: the double-buffered feeder). */
int bofi_sample_host_async(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, int32_t output_logsoftmax,
                           const float* att_feats, const int32_t* att_len, int32_t B, int32_t R,
                           int64_t* seq, float* logprobs, int32_t* phrase_num, int32_t* phrase_length,
                           int64_t* phrase_syn);

/* bofi_sample_host_async with 2-byte features in (pinned) host memory: half the H2D bytes of the fp32 form, which is what
 * bounds the end-to-end rate once several GPUs share the host's PCIe / memory bandwidth. */
int bofi_sample_host_async_ex(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, int32_t output_logsoftmax,
                              const void* att_feats, int32_t feat_dtype, const int32_t* att_len, int32_t B, int32_t R,
                              int64_t* seq, float* logprobs, int32_t* phrase_num, int32_t* phrase_length,
                              int64_t* phrase_syn);

/* Counters of the last decode (synchronises `stream`). */
int bofi_get_decode_info(bofi_handle_t h, void* stream, bofi_decode_info_t* out);

/* Per-launch CUDA-event timing of the library's own kernels, for bench.py's roofline line.
 * Classes: 0 tcgen05 GEMM, 1 FFMA GEMM, 2 attention, 3 LayerNorm, 4 vocab epilogue, 5 other.
 * bofi_set_profiling(h, 1) clears the records and starts recording; bofi_get_profile synchronises
 * `stream` and sums launches / milliseconds / algorithmic flops / algorithmic bytes per class
 * (arrays of BOFI_PROFILE_CLASSES entries). */
#define BOFI_PROFILE_CLASSES 6
int bofi_set_profiling(bofi_handle_t h, int32_t enable);
int bofi_get_profile(bofi_handle_t h, void* stream, int32_t* launches, double* ms, double* flops, double* bytes);
/* The (M, N, K) of the tcgen05 GEMM shape with the largest summed duration among the recorded launches
 * (mnk[3]), with its launch count, milliseconds and 2*M*N*K flops summed over those launches. */
int bofi_get_profile_top_gemm(bofi_handle_t h, void* stream, int32_t* mnk, int32_t* launches, double* ms, double* flops);

/* ---- XE training --------------------------------------------------------------
 * TransformerModel._forward for train_mode UIC (TransformerModel.py:1713-1775 -> EncoderDecoder_UIC.forward :413-468
 * with glat_p < 0, ss_prob == 0) and its backward pass; LanguageModelCriterion_UIC (losses.py:315-369) fused in
 * bofi_train_step_xe.  Built for N_len == 1 (uic_sd.yml).  Dropout is not applied (the reference's eval() arithmetic).
 *
 * Parameters and gradients live in two caller-owned flat fp32 device buffers of bofi_param_numel() elements;
 * entry `name` of the state_dict occupies [offset, offset + numel) (bofi_param_offset).  The drop-in module makes
 * its nn.Parameters views of these buffers, so an optimiser step on them needs no copy:
 *   bofi_train_bind(h, stream, params, grads)    copy the loaded weights into `params`, adopt both buffers
 *   bofi_refresh_weights(h, stream)              after the parameters changed: bf16 mirror + derived tables
 * Gradients are ACCUMULATED into `grads` (zero them like optimizer.zero_grad()). */
int64_t bofi_param_numel(bofi_handle_t h);
int bofi_param_offset(bofi_handle_t h, const char* name, int64_t* offset, int64_t* numel);
int bofi_train_bind(bofi_handle_t h, void* stream, float* flat_params, float* flat_grads);
int bofi_refresh_weights(bofi_handle_t h, void* stream);

/* Teacher-forced forward.  All pointers are device pointers; integer inputs are int32 with N = B * seq_per_img
 * caption rows, L = loader seq_length (16), Lb = L + 2:
 *   att_feats f32 [B,R,att_feat_size]; att_len i32 [B] or NULL
 *   labels [N,Lb]; phrase_num [N]; phrase_length [N,Lb]; ext_syn = extend_phrase_syn_seq [N,Lb];
 *   ext_seq = extend_phrase_seq [N,L]; sa_vis [N,L] = row sums of extend_phrase_seq_mask (a prefix mask,
 *   dataloader.py:391); P = max(phrase_num) (bounding passes, :492).
 * Outputs (f32, caller-owned, the six tensors of :1774-1775):
 *   sa/na_len_logp [N,Lb-1,20], sa/na_syn_logp [N,Lb-1,10], sa/na_logp [N,L,V].
 * Activations stay in the handle until the matching bofi_train_backward. */
int bofi_train_forward(bofi_handle_t h, void* stream, const float* att_feats, const int32_t* att_len, int32_t B, int32_t R,
                       int32_t seq_per_img, int32_t L, int32_t P, const int32_t* labels, const int32_t* phrase_num,
                       const int32_t* phrase_length, const int32_t* ext_syn, const int32_t* ext_seq, const int32_t* sa_vis,
                       float* sa_len_logp, float* sa_syn_logp, float* sa_logp, float* na_len_logp, float* na_syn_logp,
                       float* na_logp);
/* Backward of the last bofi_train_forward: g_* are the gradients of a scalar loss w.r.t. the six outputs (same shapes),
 * the six outputs themselves are passed back (the handle does not keep them). */
int bofi_train_backward(bofi_handle_t h, void* stream, const float* g_sa_len, const float* g_sa_syn, const float* g_sa_logp,
                        const float* g_na_len, const float* g_na_syn, const float* g_na_logp, const float* sa_len_logp,
                        const float* sa_syn_logp, const float* sa_logp, const float* na_len_logp, const float* na_syn_logp,
                        const float* na_logp);
/* One XE step without materialising the [N,L,V] log-probs: forward, criterion (reduction='mean'), backward.
 * phrase_syn i32 [N,Lb] are the syn targets.  losses: device f32[7] = total, SA_length, SA_phrase, SA_syn, NA_length,
 * NA_phrase, NA_syn (the return order of losses.py:369). */
int bofi_train_step_xe(bofi_handle_t h, void* stream, const float* att_feats, const int32_t* att_len, int32_t B, int32_t R,
                       int32_t seq_per_img, int32_t L, int32_t P, const int32_t* labels, const int32_t* phrase_num,
                       const int32_t* phrase_length, const int32_t* phrase_syn, const int32_t* ext_syn, const int32_t* ext_seq,
                       const int32_t* sa_vis, float* losses);
/* ---- self-critical training -----------------------------------------------------------
 * AttModel._sample called in train() mode with sample_method='sample' (captioning/modules/loss_wrapper.py:194-214), whose
 * seq_logprobs StructureLosses (captioning/modules/losses.py:157-176) back-propagates through.  bofi_sc_sample returns the
 * reference's tuple for N = B * sample_n rows -- seq i64 [N, L], logprobs f32 [N, L, V] (log-softmax; SAIC: zeros where
 * no word was committed), phrase_num i32 [N], phrase_length i32 [N, L], phrase_syn i64 [N, L] -- and keeps the tape of the
 * differentiable part: the encoder and ONE decoder pass on the sampled boxes (NAIC) / the sampled words (SAIC).  Tokens
 * are drawn with bofi_set_sampling's method / temperature / seed; dropout follows bofi_train_set_dropout.
 * bofi_sc_backward(g_logprobs = d loss / d logprobs, logprobs as returned) accumulates the parameter gradients into the
 * buffer bound by bofi_train_bind.  The boxes and the draws carry no gradient (argmax / multinomial in the reference too).
 * att_feats dev f32 [B,R,att_feat_size]; att_len dev i32 [B] or NULL.  Needs N_len == 1 and bofi_train_bind. */
int bofi_sc_sample(bofi_handle_t h, void* stream, int32_t mode, int32_t sample_n, const float* att_feats, const int32_t* att_len,
                   int32_t B, int32_t R, int64_t* seq, float* logprobs, int32_t* phrase_num, int32_t* phrase_length,
                   int64_t* phrase_syn);
int bofi_sc_backward(bofi_handle_t h, void* stream, const float* g_logprobs, const float* logprobs);
/* Decoder inputs of the taped pass of the last bofi_sc_sample (dev i32): word_ids / syn_ids / visible-key counts [N, L] and
 * the committed word count + 1 per row [N] -- what a checker needs to restate the pass. */
int bofi_sc_inputs(bofi_handle_t h, void* stream, int32_t* word_ids, int32_t* syn_ids, int32_t* vis, int32_t* total);

/* Dropout of the reference's train() mode for the following training calls (0, 0 = off, the eval() arithmetic):
 *   p            opt.dropout: sub-layer outputs (TransformerModel.py:1363), attention probabilities (:1430), FFN hidden
 *                (:1478), positional encoding (:1506), bounding-head hidden (:376)
 *   p_att_embed  opt.drop_prob_lm after att_embed's ReLU (:1646)
 * Masks are counter-based (murmur3 finaliser of (seed, site, element index); common.cuh: Drop): the backward pass
 * regenerates them, and oracle/bofi_oracle.py:DropSim reproduces them bit for bit.  The positional-encoding mask of
 * the bounding input is drawn once per caption and shared by its bounding passes (the reference redraws it per pass). */
int bofi_train_set_dropout(bofi_handle_t h, float p, float p_att_embed, uint32_t seed);
/* Glancing training (EncoderDecoder_UIC.forward with glat_p >= 0, TransformerModel.py:437-464; configs/uic_glat_token.yaml) for
 * the following bofi_train_forward / bofi_train_step_xe calls: a no-grad NA decoder pass predicts the words, and every real word
 * slot of a caption takes its ground-truth word instead of bos as NA decoder input with probability
 * (mismatched predictions / words) * glat_p, drawn from the library's counter-based stream seeded by `seed`.  glat_p < 0: off. */
int bofi_train_set_glat(bofi_handle_t h, float glat_p, uint32_t seed);
/* Gradient all-reduce overlap (tools/train.py:99-101 runs the reference under nn.DataParallel; here: one process per GPU).
 * `event` (a cudaEvent_t, NULL switches it off) is recorded on the training stream inside every following backward pass at
 * the point where all gradients EXCEPT those of att_embed and model.encoder.* -- the flat entries before
 * bofi_param_offset("model.decoder.layers.0.self_attn.linears.0.weight") -- are final: the caller waits for it on a side
 * stream and reduces that part of the gradient buffer while the encoder's backward pass still runs. */
int bofi_train_set_grad_event(bofi_handle_t h, void* event);
/* The same for the encoder itself: `event` (NULL switches it off) is recorded once the backward pass of encoder layer `layer` has
 * run -- the flat entries of model.encoder.layers.<layer>.* (and, for the last layer, model.encoder.norm.*) are final, so their
 * all-reduce can run underneath the backward pass of the layers below.  Only att_embed is left for after the backward pass. */
int bofi_train_set_layer_event(bofi_handle_t h, int32_t layer, void* event);
/* Kernels enqueued by the last training call. */
int bofi_train_launches(bofi_handle_t h);

/* ---- unit entry points (used by the parity tests to pin individual kernels) ------------------- */
/* LayerNorm of TransformerModel.py:1338-1349 on rows x d_model fp32 (dev pointers). */
int bofi_layernorm_f32(bofi_handle_t h, void* stream, const float* x, const float* a2, const float* b2,
                       float* out, int32_t rows);
/* out[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual) through the handle's GEMM backend
 * (fp32: SIMT FFMA; bf16: tcgen05).  A, W, bias, residual, out are dev f32; conversion to the
 * backend's operand type happens inside.  relu and residual are optional (0 / NULL). */
int bofi_linear_f32(bofi_handle_t h, void* stream, const float* A, const float* W, const float* bias,
                    const float* residual, float* out, int32_t M, int32_t N, int32_t K, int32_t relu);
/* The residual GEMM with the following LayerNorm in its epilogue (bf16 / tcgen05 engines only; SublayerConnection,
 * TransformerModel.py:1351-1363, then the LayerNorm of :1338-1349):  x[M,512] += A[M,K] @ W[512,K]^T + bias in place (fp32) and
 * y[M,512] = a2 * (x - mean) / (std + 1e-6) + b2 (computed as bf16, returned widened to f32).  All pointers dev f32; A and W
 * are rounded to bf16 inside.  K % 128 == 0.  rows_dev (optional dev i32): only the first *rows_dev rows are computed. */
int bofi_linear_resid_ln(bofi_handle_t h, void* stream, const float* A, const float* W, const float* bias, float* x,
                         const float* a2, const float* b2, float* y_out, int32_t M, int32_t K, const int32_t* rows_dev);
/* softmax(QK^T/sqrt(dk), keys >= vis masked) V for `heads` heads, fp32 dev tensors
 * (TransformerModel.py:1421-1432): q [B,Tq,d], k/v [B,Tk,d], vis i32 [B,Tq] visible-key counts. */
int bofi_attention_f32(bofi_handle_t h, void* stream, const float* q, const float* k, const float* v,
                       const int32_t* vis, float* out, int32_t B, int32_t Tq, int32_t Tk);

#ifdef __cplusplus
}
#endif
#endif /* BOFI_B200_H_ */
