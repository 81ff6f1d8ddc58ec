"""CPU oracle for the BoFiCap greedy Bounding-and-Filling decode path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file.  The product path (boficap_b200/) never does.

A plain PyTorch-fp32 functional restatement of the reference algorithm, operating directly on
the reference's 311-key state_dict.  It follows the REFERENCE FORMULATION (all 22 bounding rows
recomputed every step, memory K/V re-projected every call) so that, timed on host cores, it is
an honest stand-in for the reference's own CPU path.  The per-row python loops of the reference
are restated as vectorised index arithmetic (same results; slightly *faster* than the
reference, i.e. conservative as a baseline).

Parity pinning: the reference ships no tests/golden vectors.  This oracle is
pinned against the UNMODIFIED reference model imported in the build container
(oracle/make_golden.py -> tests/golden/*.npz; tests/test_oracle_vs_reference.py runs the live
comparison whenever /root/reference is present).

Reference citations (relative to /root/reference/captioning/models/):
  prepare           TransformerModel.py:1674-1711, AttModel.py:33-51 (pack_wrapper), :113-120 (clip_att)
  layer_norm        TransformerModel.py:1338-1349
  mha / attention   TransformerModel.py:1421-1467
  ffn               TransformerModel.py:1469-1478
  encode            TransformerModel.py:1325-1336, :1365-1377, :470-474
  embed / pos       TransformerModel.py:1480-1507
  bounding_head     TransformerModel.py:357-383, :1016-1029
  decode_naic       TransformerModel.py:1823-1876, AttModel.py:203-210, :419-429
  decode_saic       TransformerModel.py:1878-1986, :515-530
  sample_next_word  CaptionModel.py:383-437
  forward_xe        TransformerModel.py:413-468 (glat_p < 0 branch), :476-565, :1713-1775 (train_mode UIC)
  loss_xe           captioning/modules/losses.py:315-369 (LanguageModelCriterion_UIC, reduction='mean')
"""
import math
import time

import torch
import torch.nn.functional as F

SYN_LOWER, SYN_UPPER = 4, 6      # TransformerModel.py:41-42,186-187,331-332
LENGTH_DIM, SYN_DIM = 20, 10     # TransformerModel.py:329-330


class OracleConfig:
    def __init__(self, **kw):
        self.N_enc = 6
        self.N_dec = 6
        self.N_len = 1
        self.d_model = 512
        self.d_ff = 2048
        self.h = 8
        self.seq_length = 20
        self.vocab_size = 9487
        self.pad_idx, self.bos_idx, self.eos_idx, self.len_idx = 0, 1, 2, 3
        self.decoder_input_mode = "add"
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def tgt_vocab(self):
        return self.vocab_size + 4


class BofiOracle:
    def __init__(self, state_dict, cfg=None, record=False, device="cpu"):
        # device="cuda": the same eager PyTorch restatement on the GPU (bench.py's `gpu_eager` bar); tests use the CPU
        self.device = torch.device(device)
        self.sd = {k: v.detach().float().to(self.device) for k, v in state_dict.items()}
        self.cfg = cfg or OracleConfig()
        self.record = record
        self.trace = {}

    # ---- primitives -----------------------------------------------------------------------
    def lin(self, p, x):
        return F.linear(x, self.sd[p + ".weight"], self.sd[p + ".bias"])

    def layer_norm(self, p, x, eps=1e-6):
        mean = x.mean(-1, keepdim=True)
        std = x.std(-1, keepdim=True)          # unbiased (N-1); eps added to std, not variance
        return self.sd[p + ".a_2"] * (x - mean) / (std + eps) + self.sd[p + ".b_2"]

    def mha(self, p, q_in, kv_in, mask):
        """mask: bool/float broadcastable to [B,1,Tq,Tk] after unsqueeze(1); 0 => -inf."""
        h, d = self.cfg.h, self.cfg.d_model
        dk = d // h
        nb = q_in.size(0)
        q = self.lin(p + ".linears.0", q_in).view(nb, -1, h, dk).transpose(1, 2)
        k = self.lin(p + ".linears.1", kv_in).view(nb, -1, h, dk).transpose(1, 2)
        v = self.lin(p + ".linears.2", kv_in).view(nb, -1, h, dk).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
        if mask is not None:
            scores = scores.masked_fill(mask.unsqueeze(1) == 0, float("-inf"))
        pattn = F.softmax(scores, dim=-1)      # all-masked row -> NaN (kept on purpose)
        x = torch.matmul(pattn, v).transpose(1, 2).contiguous().view(nb, -1, d)
        return self.lin(p + ".linears.3", x)

    def ffn(self, p, x):
        return self.lin(p + ".w_2", F.relu(self.lin(p + ".w_1", x)))

    def embed(self, table, ids):
        return F.embedding(ids, self.sd["model.%s.lut.weight" % table]) * math.sqrt(self.cfg.d_model)

    def pos(self, x):
        return x + self.sd["model.pos_embed.pe"][:, : x.size(1)]

    # ---- A1: feature preparation ----------------------------------------------------------
    def prepare(self, att_feats, att_masks):
        att_feats = att_feats.float()
        if att_masks is not None:
            lens = att_masks.long().sum(1)
            max_len = int(lens.max())
            att_feats = att_feats[:, :max_len].contiguous()
            att_masks = att_masks[:, :max_len].contiguous()
            x = F.relu(self.lin("att_embed.0", att_feats))
            # pack_wrapper: the module only sees the first `lens[b]` rows; padded rows come back 0
            keep = torch.arange(max_len, device=lens.device)[None, :] < lens[:, None]
            x = x * keep[:, :, None].to(x.dtype)
            mask = att_masks
        else:
            x = F.relu(self.lin("att_embed.0", att_feats))
            mask = torch.ones(att_feats.shape[:2], dtype=torch.bool, device=att_feats.device)
        return x, mask.unsqueeze(-2)

    # ---- A2: encoder ----------------------------------------------------------------------
    def encode(self, x, src_mask):
        for l in range(self.cfg.N_enc):
            p = "model.encoder.layers.%d" % l
            y = self.layer_norm(p + ".sublayer.0.norm", x)
            x = x + self.mha(p + ".self_attn", y, y, src_mask)
            x = x + self.ffn(p + ".feed_forward", self.layer_norm(p + ".sublayer.1.norm", x))
        return self.layer_norm("model.encoder.norm", x)

    def _dec_layer(self, p, ffname, x, memory, src_mask, tgt_mask):
        y = self.layer_norm(p + ".sublayer.0.norm", x)
        x = x + self.mha(p + ".self_attn", y, y, tgt_mask)
        x = x + self.mha(p + ".src_attn", self.layer_norm(p + ".sublayer.1.norm", x), memory, src_mask)
        return x + self.ffn(p + "." + ffname, self.layer_norm(p + ".sublayer.2.norm", x))

    # ---- A8: bounding head ----------------------------------------------------------------
    def bounding_head(self, x, memory, src_mask, tgt_mask):
        lp = "model.length_predictor"
        if self.cfg.N_len == 0:
            y = self.layer_norm(lp + ".LengthPredictor.norm", x)
            out = self.layer_norm(lp + ".norm", x + self.mha(lp + ".length_attn", y, memory, src_mask))
        else:
            for l in range(self.cfg.N_len):
                x = self._dec_layer(lp + ".LengthPredictor.%d" % l, "ff", x, memory, src_mask, tgt_mask)
            out = self.layer_norm(lp + ".norm", x)
        hrow = out[:, 0, :]
        len_logp = F.log_softmax(self.lin(lp + ".Length_classifier2",
                                          F.relu(self.lin(lp + ".Length_classifier1", hrow))), dim=-1)
        syn_logp = F.log_softmax(self.lin(lp + ".Syntactic_classifier2",
                                          F.relu(self.lin(lp + ".Syntactic_classifier1", hrow))), dim=-1)
        return len_logp.argmax(1).int(), len_logp, syn_logp.argmax(1).long(), syn_logp, hrow

    # ---- A10: decoder stack ---------------------------------------------------------------
    def decoder(self, x, memory, src_mask, tgt_mask):
        for l in range(self.cfg.N_dec):
            x = self._dec_layer("model.decoder.layers.%d" % l, "feed_forward", x, memory, src_mask, tgt_mask)
        return self.layer_norm("model.decoder.norm", x)

    def decoder_input(self, word_ids, syn_ids):
        mode = self.cfg.decoder_input_mode
        if mode == "add":
            return self.pos(self.embed("tgt_embed", word_ids) + self.embed("syn_embed", syn_ids))
        raise NotImplementedError("decoder_input_mode=%s (uic_sd.yml uses 'add')" % mode)

    def logit(self, hid):
        return self.lin("model.generator.proj", hid)

    # ---- box bookkeeping shared by NAIC / SAIC --------------------------------------------
    def _box_rule(self, len_n, syn_n, last, finished):
        """Returns (write[B] bool, new_len[B] int, finished'[B]).  TransformerModel.py:1843-1867."""
        L = self.cfg.seq_length
        active = ~finished
        eos = (len_n == 0) | (syn_n < SYN_LOWER) | (syn_n > SYN_UPPER)
        clip = (len_n + last) >= (L + 1)
        new_len = torch.where(clip, (L + 1) - last, len_n)
        write = active & ~eos
        finished = finished | (active & (eos | clip))
        return write, new_len, finished

    # ---- A9 + A13: NAIC -------------------------------------------------------------------
    def decode_naic(self, memory, src_mask, output_logsoftmax=1, sample_method="greedy",
                    temperature=1.0, generator=None):
        c = self.cfg
        B, Lb, L = memory.size(0), c.seq_length + 2, c.seq_length
        phrase_num = torch.zeros(B, dtype=torch.int32)
        phrase_length = torch.zeros(B, Lb, dtype=torch.int32)
        phrase_syn = torch.full((B, Lb), c.pad_idx, dtype=torch.long)
        ext = torch.full((B, Lb), c.pad_idx, dtype=torch.long)
        tgt_mask = torch.zeros(B, Lb, Lb, dtype=torch.bool)
        last = torch.zeros(B, dtype=torch.int32)
        finished = torch.zeros(B, dtype=torch.bool)
        ar = torch.arange(Lb)
        steps, step_len_logp, step_syn_logp = 0, [], []
        for i in range(L):
            if i == 0:
                ext[:, 0] = c.len_idx
                tgt_mask[:, :, 0] = True
                last[:] = 1
            len_n, len_logp, syn_n, syn_logp, _ = self.bounding_head(
                self.pos(self.embed("syn_embed", ext)), memory, src_mask, tgt_mask)
            steps += 1
            if self.record:
                step_len_logp.append(len_logp)
                step_syn_logp.append(syn_logp)
            write, new_len, finished = self._box_rule(len_n, syn_n, last, finished)
            wl = torch.where(write, new_len, torch.zeros_like(new_len))
            phrase_length[:, i] = wl
            phrase_syn[:, i] = torch.where(write, syn_n, torch.zeros_like(syn_n))
            phrase_num += write.int()
            span = (ar[None, :] >= last[:, None]) & (ar[None, :] < (last + wl)[:, None])
            ext = torch.where(span, syn_n[:, None].expand(-1, Lb), ext)
            newlast = last + wl
            # tgt_mask[j, last:, :last+len] = True ; tgt_mask[j, 0, :last'] = True
            rows = (ar[None, :, None] >= last[:, None, None]) | (ar[None, :, None] == 0)
            cols = ar[None, None, :] < newlast[:, None, None]
            tgt_mask = tgt_mask | (rows & cols & write[:, None, None])
            last = newlast
            if bool(finished.all()):
                break
        w = int(last[B - 1]) - 1   # stale loop index j == B-1 at TransformerModel.py:1871-1873
        fill_mask = torch.zeros(B, L, L, dtype=torch.bool)
        fill_mask[:, :, :max(w, 0)] = True
        word = torch.full((B, L), c.bos_idx, dtype=torch.long)
        hidden = self.decoder(self.decoder_input(word, ext[:, 1:-1]), memory, src_mask, fill_mask)
        logits = self.logit(hidden)
        logp = F.log_softmax(logits, dim=2) if output_logsoftmax else logits
        seq = self.sample_next_word(logp, sample_method, temperature, generator)
        total = phrase_length.sum(1)
        seq = torch.where(torch.arange(L)[None, :] >= total[:, None], torch.zeros_like(seq), seq)
        if self.record:
            self.trace.update(steps=steps, last=last.clone(), fill_width=w, ext=ext.clone(),
                              fill_hidden=hidden, logits=logits,
                              step_len_logp=step_len_logp, step_syn_logp=step_syn_logp)
        self.last_steps = steps
        return seq, logp, phrase_num, phrase_length[:, :-2], phrase_syn[:, :-2]

    # ---- A14: SAIC ------------------------------------------------------------------------
    def decode_saic(self, memory, src_mask, output_logsoftmax=1, sample_method="greedy",
                    temperature=1.0, generator=None):
        c = self.cfg
        B, Lb, L, V = memory.size(0), c.seq_length + 2, c.seq_length, c.tgt_vocab
        phrase_num = torch.zeros(B, dtype=torch.int32)
        phrase_length = torch.zeros(B, Lb, dtype=torch.int32)
        phrase_syn = torch.full((B, Lb), c.pad_idx, dtype=torch.long)
        seq = torch.full((B, Lb), c.pad_idx, dtype=torch.long)
        seq_logp = torch.zeros(B, Lb, V)
        ext_len = torch.full((B, Lb), c.pad_idx, dtype=torch.long)
        ext = torch.full((B, Lb), c.pad_idx, dtype=torch.long)
        ext_syn = torch.full((B, Lb), c.pad_idx, dtype=torch.long)
        len_mask = torch.zeros(B, Lb, Lb, dtype=torch.bool)
        phrase_mask = torch.zeros(B, Lb, Lb, dtype=torch.bool)
        finished = torch.zeros(B, dtype=torch.bool)
        seq_last = torch.zeros(B, dtype=torch.int32)
        phrase_last = torch.zeros(B, dtype=torch.int32)
        ar = torch.arange(Lb)
        steps = 0
        for i in range(1, L + 1):
            if i == 1:
                seq[:, 0] = c.bos_idx
                phrase_length[:, 0] = 1
                ext_len[:, 0] = c.len_idx
                len_mask[:, :, 0] = True
                phrase_last[:] = 1
            len_n, _, syn_n, _, _ = self.bounding_head(
                self.pos(self.embed("tgt_embed", ext_len)), memory, src_mask, len_mask)
            steps += 1
            write, new_len, finished = self._box_rule(len_n, syn_n, phrase_last, finished)
            n = torch.where(write, new_len, torch.zeros_like(new_len))
            phrase_length[:, i] = n
            phrase_syn[:, i] = torch.where(write, syn_n, torch.zeros_like(syn_n))
            phrase_num += write.int()
            m = phrase_length[:, i - 1]
            # position-wise copy of the previous phrase (TransformerModel.py:1928-1948)
            for j in range(B):
                nj, mj, p, s = int(n[j]), int(m[j]), int(phrase_last[j]), int(seq_last[j])
                if nj == 0:
                    continue
                ext_syn[j, p:p + nj] = phrase_syn[j, i]
                if nj <= mj:
                    ext[j, p:p + nj] = seq[j, s + (mj - nj): s + mj]
                else:
                    pre_less, ct = mj - (nj % mj), nj // mj
                    reps = torch.tensor([ct if k < pre_less else ct + 1 for k in range(mj)])
                    ext[j, p:p + nj] = torch.repeat_interleave(seq[j, s:s + mj], reps)
                phrase_mask[j, p:, :p + nj] = True
            hidden = self.decoder(self.decoder_input(ext[:, 1:-1], ext_syn[:, 1:-1]), memory, src_mask,
                                  phrase_mask[:, 1:-1, 1:-1])
            logits = self.logit(hidden)
            logp = F.log_softmax(logits, dim=2) if output_logsoftmax else logits
            if bool(torch.isnan(logp).any()):      # "phrase nan!" early return (:1956-1958)
                self.last_steps = steps
                return seq[:, 1:-1], seq_logp[:, 1:-1], phrase_num, phrase_length[:, 1:-1], phrase_syn[:, 1:-1]
            tok = self.sample_next_word(logp, sample_method, temperature, generator)
            for j in range(B):
                nj, p = int(n[j]), int(phrase_last[j])
                if nj == 0:
                    continue
                seq[j, p:p + nj] = tok[j, p - 1:p - 1 + nj]
                seq_logp[j, p:p + nj] = logp[j, p - 1:p - 1 + nj]
                ext_len[j, p:p + nj] = tok[j, p - 1:p - 1 + nj]
                len_mask[j, p:, :p + nj] = True
                phrase_last[j] += nj
                len_mask[j, 0, :int(phrase_last[j])] = True
                seq_last[j] += int(m[j])
            if bool(finished.all()):
                break
        self.last_steps = steps
        return seq[:, 1:-1], seq_logp[:, 1:-1], phrase_num, phrase_length[:, 1:-1], phrase_syn[:, 1:-1]

    # ---- A12 ------------------------------------------------------------------------------
    @staticmethod
    def sample_next_word(logp, sample_method="greedy", temperature=1.0, generator=None):
        if sample_method == "greedy":
            return logp.argmax(2)            # first maximal index, like torch.max(dim=2)
        lp = logp / temperature
        lp = lp.masked_fill(torch.isnan(lp), -10.0)
        probs = F.softmax(lp, dim=-1)
        flat = torch.multinomial(probs.view(-1, probs.size(-1)), 1, generator=generator)
        return flat.view(logp.shape[:2])

    # ---- A13: the `_sample` entry ---------------------------------------------------------
    @torch.no_grad()
    def sample(self, fc_feats, att_feats, att_masks=None, opt=None):
        with torch.device(self.device):          # index / mask tensors are created on the oracle's device
            return self._sample(fc_feats, att_feats, att_masks, opt)

    def _sample(self, fc_feats, att_feats, att_masks=None, opt=None):
        opt = opt or {}
        mode = opt.get("train_mode", "NAIC")
        sample_n = int(opt.get("sample_n", 1))
        x, src_mask = self.prepare(att_feats, att_masks)
        memory = self.encode(x, src_mask)
        if self.record:
            self.trace["memory"] = memory
        if sample_n > 1:
            memory = memory.repeat_interleave(sample_n, 0)
            src_mask = src_mask.repeat_interleave(sample_n, 0)
        t0 = time.time()
        fn = {"NAIC": self.decode_naic, "SAIC": self.decode_saic}[mode]
        out = fn(memory, src_mask, opt.get("output_logsoftmax", 1),
                 opt.get("sample_method", "greedy"), opt.get("temperature", 1.0))
        return tuple(out) + (time.time() - t0,)


    # ---- A16: XE-training forward (teacher forcing) ---------------------------------------
    def _teacher_bounding(self, x_in, memory, src_mask, phrase_num, phrase_length):
        """get_predict_phrase_length_syn_SA / _NA (TransformerModel.py:476-513, :532-565): one bounding pass per
        phrase index, every pass over all rows with the mask grown by the ground-truth boxes."""
        N, Lb = phrase_length.shape
        dev = memory.device                      # (cuda for bench.py's eager bar)
        tgt_mask = torch.zeros(N, Lb, Lb, dtype=torch.bool, device=dev)
        len_logp = torch.zeros(N, Lb, LENGTH_DIM, device=dev)
        syn_logp = torch.zeros(N, Lb, SYN_DIM, device=dev)
        last = torch.ones(N, dtype=torch.long, device=dev)
        tgt_mask[:, :, 0] = True
        _, ll, _, sl, _ = self.bounding_head(x_in, memory, src_mask, tgt_mask)
        len_logp[:, 1], syn_logp[:, 1] = ll.to(len_logp.dtype), sl.to(syn_logp.dtype)
        ar = torch.arange(Lb, device=dev)
        for i in range(1, int(phrase_num.max())):
            grow = phrase_num > i
            newlast = last + torch.where(grow, phrase_length[:, i], torch.zeros_like(last))
            rows = (ar[None, :, None] >= last[:, None, None]) | (ar[None, :, None] == 0)
            cols = ar[None, None, :] < newlast[:, None, None]
            tgt_mask = tgt_mask | (rows & cols & grow[:, None, None])
            last = newlast
            _, ll, _, sl, _ = self.bounding_head(x_in, memory, src_mask, tgt_mask)
            len_logp = torch.cat([len_logp[:, :i + 1], ll[:, None], len_logp[:, i + 2:]], 1)
            syn_logp = torch.cat([syn_logp[:, :i + 1], sl[:, None], syn_logp[:, i + 2:]], 1)
        return len_logp[:, 1:], syn_logp[:, 1:], last

    def forward_xe(self, att_feats, att_masks, labels, phrase_num, phrase_length, extend_phrase_syn_seq,
                   extend_phrase_seq, extend_phrase_seq_mask):
        """`_forward` for train_mode UIC with glat_p < 0 and ss_prob == 0, dropout off (eval()).  Inputs may carry the
        [B, seq_per_img, ...] layout of the loader.  Returns the six log-prob tensors of TransformerModel.py:1774-1775."""
        c = self.cfg
        if labels.dim() == 3:
            labels = labels.reshape(-1, labels.shape[2])
            phrase_num = phrase_num.reshape(-1)
            phrase_length = phrase_length.reshape(-1, phrase_length.shape[2])
            extend_phrase_syn_seq = extend_phrase_syn_seq.reshape(-1, extend_phrase_syn_seq.shape[2])
            extend_phrase_seq = extend_phrase_seq.reshape(-1, extend_phrase_seq.shape[2])
            L = extend_phrase_seq.shape[1]
            extend_phrase_seq_mask = extend_phrase_seq_mask.reshape(-1, L, L)
        spi = labels.shape[0] // att_feats.shape[0]
        x, src_mask = self.prepare(att_feats, att_masks)
        memory = self.encode(x, src_mask)                      # encoding once per image == encoding the repeats
        memory = memory.repeat_interleave(spi, 0)
        src_mask = src_mask.repeat_interleave(spi, 0)
        phrase_num, phrase_length = phrase_num.long(), phrase_length.long()
        word_seq = labels.clone().long()
        word_seq[:, 0] = c.len_idx
        sa_len, sa_syn, _ = self._teacher_bounding(self.pos(self.embed("tgt_embed", word_seq)), memory, src_mask,
                                                   phrase_num, phrase_length)
        sa_hidden = self.decoder(self.decoder_input(extend_phrase_seq, extend_phrase_syn_seq[:, 1:-1]), memory, src_mask,
                                 extend_phrase_seq_mask)
        na_len, na_syn, last = self._teacher_bounding(self.pos(self.embed("syn_embed", extend_phrase_syn_seq)), memory,
                                                      src_mask, phrase_num, phrase_length)
        L = extend_phrase_seq.shape[1]
        syn_mask = torch.arange(L, device=last.device)[None, None, :] < (last - 1)[:, None, None]
        syn_mask = syn_mask.expand(-1, L, -1)
        bos = torch.full_like(extend_phrase_seq, c.bos_idx)
        na_hidden = self.decoder(self.decoder_input(bos, extend_phrase_syn_seq[:, 1:-1]), memory, src_mask, syn_mask)
        sa_logp = F.log_softmax(self.logit(sa_hidden), dim=-1)   # Generator.forward (TransformerModel.py:1315-1323)
        na_logp = F.log_softmax(self.logit(na_hidden), dim=-1)
        return sa_len, sa_syn, sa_logp, na_len, na_syn, na_logp

    @staticmethod
    def loss_xe(outs, phrase_num, phrase_length, phrase_syn, labels):
        """LanguageModelCriterion_UIC, reduction='mean', self_dis=False (losses.py:315-369): six masked NLL sums, each
        divided by the number of word slots."""
        sa_len, sa_syn, sa_logp, na_len, na_syn, na_logp = outs
        if labels.dim() == 3:
            phrase_num = phrase_num.reshape(-1)
            phrase_length = phrase_length.reshape(-1, phrase_length.shape[2])
            phrase_syn = phrase_syn.reshape(-1, phrase_syn.shape[2])
            labels = labels.reshape(-1, labels.shape[2])
        dev = sa_logp.device
        phrase_num, phrase_length, phrase_syn, labels = [t.to(dev) for t in (phrase_num, phrase_length, phrase_syn, labels)]
        words = labels[:, 1:-1].long()
        T = words.shape[1]
        word_mask = (torch.arange(T, device=dev)[None, :] < (phrase_length.sum(1) - 1)[:, None]).to(sa_logp.dtype)
        box_mask = (torch.arange(phrase_length.shape[1] - 1, device=dev)[None, :] < phrase_num[:, None]).to(sa_logp.dtype)
        len_t, syn_t = phrase_length[:, 1:].long(), phrase_syn[:, 1:].long()
        nll = lambda lp, tgt, m: -(lp.gather(2, tgt.unsqueeze(2)).squeeze(2) * m).sum()
        denom = word_mask.sum()
        parts = [nll(sa_len, len_t, box_mask) / denom, nll(sa_logp, words, word_mask) / denom, nll(sa_syn, syn_t, box_mask) / denom,
                 nll(na_len, len_t, box_mask) / denom, nll(na_logp, words, word_mask) / denom, nll(na_syn, syn_t, box_mask) / denom]
        return sum(parts), parts


# ---- A16 with dropout: the CUDA path's formulation, masks reproduced bit for bit -------------------------------
class DropSim:
    """Counter-based dropout masks of the CUDA training path (boficap_b200/csrc/common.cuh: drop_hash / Drop,
    train.inl: TrainState::next_drop).  Sites are numbered in forward order from 0; key(site) = drop_hash(seed, site);
    element `idx` of a site is kept iff drop_hash(key, idx) >= p * 2^32 and then scaled by 1 / (1 - p)."""
    M = 0xFFFFFFFF

    def __init__(self, p_sub, p_att, seed):
        self.p_sub, self.p_att, self.seed, self.site = float(p_sub), float(p_att), int(seed) & self.M, 0

    @classmethod
    def _hash(cls, key, idx):
        h = ((idx * 0x9E3779B1) & cls.M) ^ key
        h = h ^ (h >> 16)
        h = (h * 0x85EBCA6B) & cls.M
        h = h ^ (h >> 13)
        h = (h * 0xC2B2AE35) & cls.M
        return h ^ (h >> 16)

    def _next(self, p, idx):
        site = self.site
        self.site += 1
        if p <= 0:
            return torch.ones(idx.shape)
        key = int(self._hash(self.seed, torch.tensor(site, dtype=torch.int64)))
        thresh = min(4294967295, int(p * 4294967296.0))
        scale = float(torch.tensor(1.0, dtype=torch.float32) / (torch.tensor(1.0, dtype=torch.float32) - torch.tensor(p, dtype=torch.float32)))
        return (self._hash(key, idx) >= thresh).float() * scale

    def elem(self, rows, width, p=None):
        """mask [rows, width] of an element-wise site (index = row * width + col)."""
        idx = torch.arange(rows * width, dtype=torch.int64).view(rows, width)
        return self._next(self.p_sub if p is None else p, idx)

    def att(self, nrows, Tk):
        """mask [nrows, heads=8, Tk] of an attention-probability site (index = ((row * 8 + head) * 128 + key))."""
        r = torch.arange(nrows, dtype=torch.int64)[:, None, None]
        h = torch.arange(8, dtype=torch.int64)[None, :, None]
        k = torch.arange(Tk, dtype=torch.int64)[None, None, :]
        return self._next(self.p_sub, (r * 8 + h) * 128 + k)


def _fused_methods():
    def mha_d(self, p, q_in, kv_in, mask, amask):
        """mha with a dropout mask amask [nb*Tq, h, Tk] on the probabilities (or None)."""
        h, d = self.cfg.h, self.cfg.d_model
        dk = d // h
        nb, Tq = q_in.shape[:2]
        q = self.lin(p + ".linears.0", q_in).view(nb, -1, h, dk).transpose(1, 2)
        k = self.lin(p + ".linears.1", kv_in).view(nb, -1, h, dk).transpose(1, 2)
        v = self.lin(p + ".linears.2", kv_in).view(nb, -1, h, dk).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(dk)
        if mask is not None:
            scores = scores.masked_fill(mask.unsqueeze(1) == 0, float("-inf"))
        pattn = F.softmax(scores, dim=-1)
        if amask is not None:
            pattn = pattn * amask.view(nb, Tq, h, -1).transpose(1, 2)
        x = torch.matmul(pattn, v).transpose(1, 2).contiguous().view(nb, -1, d)
        return self.lin(p + ".linears.3", x)

    def layer_d(self, p, ffname, x, memory, src_mask, tgt_mask, drop, cross):
        nb, T = x.shape[:2]
        flat = lambda m: m.view(nb, T, -1)
        a1, o1 = drop.att(nb * T, T), drop.elem(nb * T, self.cfg.d_model)
        x = x + flat(o1) * mha_d(self, p + ".self_attn", self.layer_norm(p + ".sublayer.0.norm", x), self.layer_norm(p + ".sublayer.0.norm", x), tgt_mask, a1)
        f = 1
        if cross:
            a2, o2 = drop.att(nb * T, memory.shape[1]), drop.elem(nb * T, self.cfg.d_model)
            x = x + flat(o2) * mha_d(self, p + ".src_attn", self.layer_norm(p + ".sublayer.1.norm", x), memory, src_mask, a2)
            f = 2
        hmask, o3 = drop.elem(nb * T, self.cfg.d_ff), drop.elem(nb * T, self.cfg.d_model)
        y = self.layer_norm(p + ".sublayer.%d.norm" % f, x)
        hid = F.relu(self.lin(p + "." + ffname + ".w_1", y)) * flat(hmask)
        return x + flat(o3) * self.lin(p + "." + ffname + ".w_2", hid)

    def bound_d(self, x_in, memory, src_mask, vis_b, drop):
        """All P bounding passes as one batch of (n, p) single-query rows (N_len == 1); x_in [N, Tb, 512] already carries its
        positional-encoding dropout; vis_b [N, P] visible-key counts of the [LEN] row."""
        N, Tb, d = x_in.shape
        P = vis_b.shape[1]
        h, dk = self.cfg.h, d // self.cfg.h
        lp = "model.length_predictor"
        p = lp + ".LengthPredictor.0"
        y0 = self.layer_norm(p + ".sublayer.0.norm", x_in)
        q0 = self.lin(p + ".self_attn.linears.0", y0[:, 0]).view(N, h, dk)
        k = self.lin(p + ".self_attn.linears.1", y0).view(N, Tb, h, dk)
        v = self.lin(p + ".self_attn.linears.2", y0).view(N, Tb, h, dk)
        a_self, o_self = drop.att(N * P, Tb), drop.elem(N * P, d)
        a_cross, o_cross = drop.att(N * P, memory.shape[1]), drop.elem(N * P, d)
        hmask, o_ffn, hidmask = drop.elem(N * P, self.cfg.d_ff), drop.elem(N * P, d), drop.elem(N * P, 200)
        scores = torch.einsum("nhd,nkhd->nhk", q0, k) / math.sqrt(dk)                       # [N, h, Tb]
        scores = scores[:, None].expand(N, P, h, Tb)
        visible = torch.arange(Tb)[None, None, None, :] < vis_b[:, :, None, None]
        pattn = F.softmax(scores.masked_fill(~visible, float("-inf")), dim=-1) * a_self.view(N, P, h, Tb)
        ao = torch.einsum("nphk,nkhd->nphd", pattn, v).reshape(N * P, d)
        x = x_in[:, 0].repeat_interleave(P, 0) + o_self * self.lin(p + ".self_attn.linears.3", ao)
        mem_rep = memory.repeat_interleave(P, 0)
        src_rep = src_mask.repeat_interleave(P, 0)
        x = x[:, None] + (o_cross * mha_d(self, p + ".src_attn", self.layer_norm(p + ".sublayer.1.norm", x)[:, None], mem_rep, src_rep,
                                          a_cross).squeeze(1))[:, None]
        x = x.squeeze(1)
        y = self.layer_norm(p + ".sublayer.2.norm", x)
        x = x + o_ffn * self.lin(p + ".ff.w_2", F.relu(self.lin(p + ".ff.w_1", y)) * hmask)
        hn = self.layer_norm(lp + ".norm", x)
        hid_len = F.relu(self.lin(lp + ".Length_classifier1", hn)) * hidmask[:, :100]
        hid_syn = F.relu(self.lin(lp + ".Syntactic_classifier1", hn)) * hidmask[:, 100:]
        ll = F.log_softmax(self.lin(lp + ".Length_classifier2", hid_len), dim=-1).view(N, P, -1)
        sl = F.log_softmax(self.lin(lp + ".Syntactic_classifier2", hid_syn), dim=-1).view(N, P, -1)
        pad = lambda t: torch.cat([t, torch.zeros(N, Tb - 1 - P, t.shape[2])], 1)
        return pad(ll), pad(sl)

    def forward_xe_fused(self, att_feats, att_masks, labels, phrase_num, phrase_length, extend_phrase_syn_seq, extend_phrase_seq,
                         extend_phrase_seq_mask, drop=None, glat_p=-1.0, glat_seed=0):
        """The training forward in the CUDA path's formulation (boficap_b200/csrc/train.inl): encoder once per image, memory
        shared by the captions of an image, all bounding passes as one batch of [LEN] rows.  With drop=None it equals
        forward_xe (tested); with a DropSim it reproduces the CUDA path's dropout masks element for element."""
        c = self.cfg
        drop = drop or DropSim(0.0, 0.0, 0)
        if labels.dim() == 3:
            labels = labels.reshape(-1, labels.shape[2])
            phrase_num = phrase_num.reshape(-1)
            phrase_length = phrase_length.reshape(-1, phrase_length.shape[2])
            extend_phrase_syn_seq = extend_phrase_syn_seq.reshape(-1, extend_phrase_syn_seq.shape[2])
            extend_phrase_seq = extend_phrase_seq.reshape(-1, extend_phrase_seq.shape[2])
            L = extend_phrase_seq.shape[1]
            extend_phrase_seq_mask = extend_phrase_seq_mask.reshape(-1, L, L)
        B, R = att_feats.shape[:2]
        N, Tb = labels.shape
        T, spi = Tb - 2, N // B
        drop.site = 0
        m_att = drop.elem(B * R, c.d_model, drop.p_att).view(B, R, -1)
        x = F.relu(self.lin("att_embed.0", att_feats.float())) * m_att
        if att_masks is not None:
            x = x * att_masks[:, :, None].to(x.dtype)
            src_mask = att_masks.unsqueeze(-2)
        else:
            src_mask = torch.ones(B, 1, R, dtype=torch.bool)
        for l in range(c.N_enc):
            x = layer_d(self, "model.encoder.layers.%d" % l, "feed_forward", x, None, None, src_mask, drop, False)
        memory = self.layer_norm("model.encoder.norm", x)
        memory_n, src_n = memory.repeat_interleave(spi, 0), src_mask.repeat_interleave(spi, 0)
        phrase_num, phrase_length = phrase_num.long(), phrase_length.long()
        P = int(phrase_num.max())
        steps = torch.arange(P)[None, :]
        grow = (steps >= 1) & (steps < phrase_num[:, None])
        vis_b = 1 + torch.cumsum(torch.where(grow, phrase_length[:, :P], torch.zeros_like(phrase_length[:, :P])), 1)
        word_seq = labels.clone().long()
        word_seq[:, 0] = c.len_idx

        def decode(word_ids, tgt_mask):
            xin = self.decoder_input(word_ids, extend_phrase_syn_seq[:, 1:-1]) * drop.elem(N * T, c.d_model).view(N, T, -1)
            for l in range(c.N_dec):
                xin = layer_d(self, "model.decoder.layers.%d" % l, "feed_forward", xin, memory_n, src_n, tgt_mask, drop, True)
            return F.log_softmax(self.logit(self.layer_norm("model.decoder.norm", xin)), dim=-1)

        sa_in = self.pos(self.embed("tgt_embed", word_seq)) * drop.elem(N * Tb, c.d_model).view(N, Tb, -1)
        sa_len, sa_syn = bound_d(self, sa_in, memory_n, src_n, vis_b, drop)
        sa_logp = decode(extend_phrase_seq, extend_phrase_seq_mask)
        na_in = self.pos(self.embed("syn_embed", extend_phrase_syn_seq)) * drop.elem(N * Tb, c.d_model).view(N, Tb, -1)
        na_len, na_syn = bound_d(self, na_in, memory_n, src_n, vis_b, drop)
        syn_mask = (torch.arange(T)[None, None, :] < (vis_b[:, -1] - 1)[:, None, None]).expand(-1, T, -1)
        na_words = torch.full_like(extend_phrase_seq, c.bos_idx)
        if glat_p >= 0:
            # glancing (EncoderDecoder_UIC.forward :437-464): no-grad NA pass -> predictions; a real word slot takes its ground-truth
            # word with probability (mismatches / words) * glat_p.  The uniform draws are the CUDA path's counter-based ones
            # (train_kernels.cuh: glat_input_kernel; the reference uses torch.rand), so the two inputs are identical.
            with torch.no_grad():
                pred = decode(na_words, syn_mask).argmax(-1)
            real = labels[:, 1:-1].long()
            length = phrase_length.sum(1) - 1
            pos = torch.arange(T)[None, :] < length[:, None]
            same = ((pred == real) & pos).sum(1)
            mismatch = torch.where(length > 0, (length - same).float() / length.clamp(min=1).float(), torch.zeros(N))
            keep_prob = (mismatch * torch.tensor(glat_p, dtype=torch.float32))[:, None] * pos.float()
            key = int(DropSim._hash(int(glat_seed) & DropSim.M, torch.tensor(0x474C4154, dtype=torch.int64)))
            u = (DropSim._hash(key, torch.arange(N * T, dtype=torch.int64)).float() + 0.5) * (1.0 / 4294967296.0)
            keep = u.view(N, T) < keep_prob
            na_words = torch.where(keep, real, na_words)
            self.trace["glat_words"] = na_words
        na_logp = decode(na_words, syn_mask)
        return sa_len, sa_syn, sa_logp, na_len, na_syn, na_logp

    def forward_sc(self, att_feats, att_masks, word_ids, syn_ids, vis, sample_n, drop=None):
        """The differentiable part of sampling in train() mode (AttModel._sample under loss_wrapper.py:194-214) in the CUDA
        path's formulation (boficap_b200/csrc/train_abi.inl: sc_sample_impl): encoder once per image with dropout, then ONE
        decoder pass on fixed inputs -- word_ids / syn_ids [N, L] (NAIC: bos everywhere + the syn label of every slot, the
        decode_NA input :570-587; SAIC: the position-wise copied words of :1928-1948) under the prefix masks `vis` [N, L]
        (visible keys per query slot) -- and the log-softmax of the generator.  N = B * sample_n rows, row n uses image
        n // sample_n.  Returns log-probs [N, L, V]; autograd gives the oracle gradients of any loss on them."""
        c = self.cfg
        drop = drop or DropSim(0.0, 0.0, 0)
        B, R = att_feats.shape[:2]
        N, T = syn_ids.shape
        drop.site = 0
        m_att = drop.elem(B * R, c.d_model, drop.p_att).view(B, R, -1)
        x = F.relu(self.lin("att_embed.0", att_feats.float())) * m_att
        if att_masks is not None:
            x = x * att_masks[:, :, None].to(x.dtype)
            src_mask = att_masks.unsqueeze(-2)
        else:
            src_mask = torch.ones(B, 1, R, dtype=torch.bool)
        for l in range(c.N_enc):
            x = layer_d(self, "model.encoder.layers.%d" % l, "feed_forward", x, None, None, src_mask, drop, False)
        memory = self.layer_norm("model.encoder.norm", x)
        memory_n, src_n = memory.repeat_interleave(sample_n, 0), src_mask.repeat_interleave(sample_n, 0)
        tgt_mask = torch.arange(T)[None, None, :] < vis.long()[:, :, None]
        xin = self.decoder_input(word_ids.long(), syn_ids.long()) * drop.elem(N * T, c.d_model).view(N, T, -1)
        for l in range(c.N_dec):
            xin = layer_d(self, "model.decoder.layers.%d" % l, "feed_forward", xin, memory_n, src_n, tgt_mask, drop, True)
        return F.log_softmax(self.logit(self.layer_norm("model.decoder.norm", xin)), dim=-1)

    BofiOracle.forward_xe_fused = forward_xe_fused
    BofiOracle.forward_sc = forward_sc


_fused_methods()
