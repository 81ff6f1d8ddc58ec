"""`TransformerModel` drop-in: a parameter tree with the reference's exact state_dict layout whose
sampling path runs on the hand-written sm_100a kernels behind libbofi_b200.so.

Reference behaviour mirrored here (paths relative to /root/reference/captioning/models/):
  forward(*args, mode=)        CaptionModel.py:42-46
  _sample option parsing       AttModel.py:307-334
  NAIC / SAIC return tuples    AttModel.py:419-437  (seq, seq_logprob, phrase_num, phrase_length, phrase_syn, elapsed)
  sample_n repeat              AttModel.py:331-334, utils.py:3-14
  state_dict layout            TransformerModel.py:1558-1568, :1626-1666
The module holds NO PyTorch math for this path: without the CUDA library it raises.
"""

import torch
import torch.nn as nn

from ...engine import BofiEngine
from ...layout import BofiConfig, state_spec
from ...synth import sinusoid_table


class _Node(nn.Module):
    """Anonymous container so that dotted reference keys map onto nested modules."""


def _attach(root, dotted, tensor, is_buffer):
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Node())
        mod = mod._modules[p]
    if is_buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor))


class _XEForward(torch.autograd.Function):
    """Autograd bridge of the teacher-forced forward: the six log-prob tensors come from bofi_train_forward; their
    gradients go to bofi_train_backward, which accumulates straight into the parameters' `.grad` storage (the flat
    gradient buffer), so no gradient is returned to autograd for the parameters themselves."""

    @staticmethod
    def forward(ctx, model, att_feats, att_len, batch, anchor):
        eng = model._engine
        outs = eng.train_forward(att_feats, att_len, batch)
        ctx.model = model
        ctx.save_for_backward(*outs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        outs = ctx.saved_tensors
        grads = [g if g is not None else torch.zeros_like(o) for g, o in zip(grads, outs)]
        ctx.model._engine.train_backward(grads, outs)
        return None, None, None, None, None


class _SCSample(torch.autograd.Function):
    """Autograd bridge of sampling in train() mode (loss_wrapper.py:194-214): bofi_sc_sample returns the sampled captions and
    their log-probs; the gradient of a structure loss w.r.t. those log-probs goes to bofi_sc_backward, which accumulates
    straight into the flat gradient buffer (no gradient is returned to autograd for the parameters themselves)."""

    @staticmethod
    def forward(ctx, model, att_feats, att_len, mode, sample_n, anchor):
        seq, logp, pnum, plen, psyn = model._engine.sc_sample(att_feats, att_len, mode, sample_n)
        ctx.model = model
        ctx.save_for_backward(logp)
        ctx.mark_non_differentiable(seq, pnum, plen, psyn)
        return seq, logp, pnum, plen, psyn

    @staticmethod
    def backward(ctx, g_seq, g_logp, g_pnum, g_plen, g_psyn):
        (logp,) = ctx.saved_tensors
        if g_logp is not None:
            ctx.model._engine.sc_backward(g_logp, logp)
        return None, None, None, None, None, None


class TransformerModel(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.cfg = BofiConfig.from_opt(opt)
        c = self.cfg
        self.vocab_size, self.tgt_vocab = c.vocab_size, c.tgt_vocab
        self.seq_length = c.seq_length
        self.train_mode = c.train_mode
        self.pad_idx, self.bos_idx, self.eos_idx, self.len_idx = c.pad_idx, c.bos_idx, c.eos_idx, c.len_idx
        self.N_enc, self.N_dec, self.N_len = c.N_enc, c.N_dec, c.N_len
        self.d_model, self.d_ff, self.h = c.d_model, c.d_ff, c.h
        self.vocab = getattr(opt, "vocab", None)
        self.ss_prob = 0.0
        self.precision = getattr(opt, "bofi_precision", "fp32")
        for name, (shape, kind) in state_spec(c).items():
            if kind == "pe":
                _attach(self, name, sinusoid_table(shape[1], shape[2]), True)
                continue
            t = torch.empty(shape)
            if kind in ("matrix", "embedding"):
                nn.init.xavier_uniform_(t)               # make_model, TransformerModel.py:1621-1623
            elif kind == "bias":
                fan_in = state_spec(c)[name[:-4] + "weight"][0][1]
                nn.init.uniform_(t, -fan_in ** -0.5, fan_in ** -0.5)
            elif kind == "ones":
                t.fill_(1.0)
            else:
                t.zero_()
            _attach(self, name, t, False)
        self._engine = None
        self._engine_key = None
        self._bound = None          # (flat_w, flat_g, version stamp) once the parameters are views of the training buffers
        self.bofi_dropout_seed = 0  # base seed of the counter-based dropout masks; the step counter is added to it
        self._train_steps = 0

    # ---- training: parameters as views of the library's flat buffers ---------------------------
    def train_bind(self, device=None, precision=None):
        """Moves the parameters into the flat parameter buffer shared with libbofi_b200.so and their `.grad`s into the
        flat gradient buffer (boficap_b200/engine.py:train_bind).  Call after the model is on its device; optimisers
        must be created afterwards (they hold the re-pointed parameters)."""
        eng = self.engine(device, precision)
        flat_w, flat_g = eng.train_bind()
        layout = eng.param_layout()
        for name, p in list(self.named_parameters()) + list(self.named_buffers()):
            off, n = layout[name]
            p.data = flat_w[off:off + n].view(p.shape)
            if isinstance(p, nn.Parameter):
                p.grad = flat_g[off:off + n].view(p.shape)
        self._bound = [flat_w, flat_g, self._version_stamp(), layout]
        self._engine_key = None
        return self

    def _version_stamp(self):
        return sum(p._version for p in self.parameters())

    def _reattach_grads(self):
        """`optimizer.zero_grad()` defaults to set_to_none=True (tools/train.py:210 calls it every step): it drops the
        `.grad` views, after which `optimizer.step()` would skip every parameter while the library kept accumulating
        into the flat buffer.  Before every training forward the views are put back; when they were dropped the flat
        buffer is zeroed, which is what zero_grad meant."""
        flat_g = self._bound[1]
        lo, hi = flat_g.data_ptr(), flat_g.data_ptr() + flat_g.numel() * 4
        missing = [(n, p) for n, p in self.named_parameters()
                   if p.grad is None or not (lo <= p.grad.data_ptr() < hi)]
        if not missing:
            return
        if any(p.grad is not None for _, p in missing):
            raise RuntimeError("a parameter's .grad was replaced by a tensor outside the flat gradient buffer; "
                               "use optimizer.zero_grad() / model.zero_grad() instead of assigning .grad")
        flat_g.zero_()
        layout = self._bound[3]
        for n, p in missing:
            off, k = layout[n]
            p.grad = flat_g[off:off + k].view(p.shape)

    def flat_grads(self):
        return self._bound[1]

    def flat_params(self):
        return self._bound[0]

    def zero_grad(self, set_to_none=False):          # the gradients are views of one buffer: always zero in place
        if self._bound is not None:
            self._bound[1].zero_()
            return
        super().zero_grad(set_to_none=set_to_none)

    # ---- engine management -----------------------------------------------------------------
    def _weights_key(self):
        # version stamp only: load_state_dict / optimiser steps / in-place edits bump the parameters' version counters
        # (one integer sum per call instead of a 311-tuple of pointers and versions)
        return (self._version_stamp(), len(self._parameters) + len(self._modules))

    def engine(self, device=None, precision=None):
        precision = precision or self.precision
        device = torch.device(device if device is not None else "cuda")
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._bound is not None and self._engine is not None and self._engine.precision == precision:
            stamp = self._version_stamp()               # optimiser steps write the shared buffer in place
            if stamp != self._bound[2]:
                self._engine.refresh_weights()
                self._bound[2] = stamp
            return self._engine
        key = (index, precision, self._weights_key())
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            self._engine = BofiEngine(self.cfg, index, precision).load_state_dict(self.state_dict())
            self._engine_key = key
        return self._engine

    # ---- reference-facing API ------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        mode = kwargs.pop("mode", "forward")
        return getattr(self, "_" + mode)(*args, **kwargs)

    @staticmethod
    def _xe_batch(att_feats, seq, phrase_num, phrase_length, phrase_syn, extend_phrase_syn_seq, extend_phrase_seq,
                  extend_phrase_seq_mask):
        """The loader's [B, seq_per_img, ...] tensors as 2-D caption rows on the feature device (the reshape of
        TransformerModel.py:1714-1722); the boolean phrase-block mask becomes its visible-key counts."""
        dev = att_feats.device
        r2 = lambda t: t.reshape(-1, t.shape[-1]).to(dev) if t is not None else None
        labels = r2(seq)
        ext_seq = r2(extend_phrase_seq)
        L = ext_seq.shape[1]
        mask = extend_phrase_seq_mask.reshape(-1, L, L).to(dev)
        sa_vis = mask.sum(-1)
        pn = phrase_num.reshape(-1)
        return dict(labels=labels, phrase_num=pn.to(dev), phrase_length=r2(phrase_length), phrase_syn=r2(phrase_syn),
                    extend_phrase_syn_seq=r2(extend_phrase_syn_seq), extend_phrase_seq=ext_seq, sa_vis=sa_vis, P=int(pn.max()))

    def _train_engine(self, att_feats):
        if not att_feats.is_cuda:
            raise RuntimeError("boficap_b200 runs on CUDA tensors only (no CPU fallback)")
        if self._bound is None:
            self.train_bind(att_feats.device)
        self._reattach_grads()
        eng = self.engine(att_feats.device)
        # nn.Module.train() / eval() decide, as for the reference's nn.Dropout modules
        self._step_seed = self.bofi_dropout_seed + self._train_steps      # also seeds the glancing draws of this step
        if self.training:
            eng.train_set_dropout(self.cfg.dropout, self.cfg.drop_prob_lm, self._step_seed)
            self._train_steps += 1
        else:
            eng.train_set_dropout(0.0, 0.0, 0)
        return eng

    def _forward(self, fc_feats, att_feats, seq, att_masks=None, phrase_num=None, phrase_length=None, phrase_syn=None,
                 extend_phrase_syn_seq=None, extend_phrase_seq=None, extend_phrase_seq_mask=None, glat_p=-1.0):
        """TransformerModel._forward, train_mode UIC (TransformerModel.py:1713-1775): the six log-prob tensors
        (SA length / syn / word, NA length / syn / word), differentiable w.r.t. the parameters."""
        if self.ss_prob > 0:
            raise NotImplementedError("scheduled sampling (ss_prob > 0, TransformerModel.py:1759-1766) is not built")
        eng = self._train_engine(att_feats)
        # glancing (TransformerModel.py:437-464): the keep draws use the library's counter-based stream, seeded like dropout
        eng.train_set_glat(glat_p, self._step_seed)
        batch = self._xe_batch(att_feats, seq, phrase_num, phrase_length, phrase_syn, extend_phrase_syn_seq, extend_phrase_seq,
                               extend_phrase_seq_mask)
        att_len = self._train_att_len(att_masks)
        anchor = next(self.parameters())                  # makes the outputs require grad
        return _XEForward.apply(self, att_feats.float(), att_len, batch, anchor)

    def xe_step(self, fc_feats, att_feats, seq, att_masks=None, phrase_num=None, phrase_length=None, phrase_syn=None,
                extend_phrase_syn_seq=None, extend_phrase_seq=None, extend_phrase_seq_mask=None, glat_p=-1.0):
        """One XE training step's forward + LanguageModelCriterion_UIC(reduction='mean') + backward in a single library
        call (bofi_train_step_xe): gradients are accumulated into the parameters' `.grad`; returns the criterion's seven
        values (total, SA_length, SA_phrase, SA_syn, NA_length, NA_phrase, NA_syn) as a device tensor."""
        eng = self._train_engine(att_feats)
        eng.train_set_glat(glat_p, self._step_seed)
        batch = self._xe_batch(att_feats, seq, phrase_num, phrase_length, phrase_syn, extend_phrase_syn_seq, extend_phrase_seq,
                               extend_phrase_seq_mask)
        att_len = self._train_att_len(att_masks)
        return eng.train_step_xe(att_feats.float(), att_len, batch)

    def _train_att_len(self, att_masks):
        if att_masks is None:
            return None
        return self._engine.masks_to_len(att_masks.data)    # a non-prefix mask is reported by the next check_masks()

    def _sample_with_tape(self, att_feats, att_masks, train_mode, sample_n, temperature, output_logsoftmax):
        """`_sample` as self-critical training calls it (train() mode, sample_method='sample', gradients enabled,
        loss_wrapper.py:194-214): the returned seq_logprobs are differentiable w.r.t. the parameters (bofi_sc_sample /
        bofi_sc_backward); dropout is on like in every other training forward."""
        if not output_logsoftmax:
            raise NotImplementedError("sampling with a tape returns log-softmax outputs (output_logsoftmax=1, what nscl uses)")
        eng = self._train_engine(att_feats)
        att_len = self._train_att_len(att_masks)
        eng.set_sampling("sample", temperature, int(torch.randint(0, 2 ** 31 - 1, (1,))))
        stream = torch.cuda.current_stream(att_feats.device)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        anchor = next(self.parameters())
        seq, logp, pnum, plen, psyn = _SCSample.apply(self, att_feats.float(), att_len, train_mode, sample_n, anchor)
        eng.set_sampling("greedy")
        ev1.record(stream)
        ev1.synchronize()
        return seq, logp, pnum, plen, psyn, ev0.elapsed_time(ev1) * 1e-3

    def sample_stats(self, fc_feats, att_feats, att_masks=None, opt={}):
        """`_sample` for evaluation loops that only need captions, entropy and perplexity (eval_utils.py:176-184): the
        [B, L, V] log-prob tensor is never materialised, the two statistics come out of the vocab epilogue.
        Returns (seq, entropy [B], perplexity [B], phrase_num, phrase_length, phrase_syn, elapsed)."""
        rows = att_feats.shape[0] * int(opt.get("sample_n", 1))
        ent = torch.empty(rows, self.seq_length, device=att_feats.device)
        lp = torch.empty(rows, self.seq_length, device=att_feats.device)
        seq, _, pnum, plen, psyn, dt = self._sample(fc_feats, att_feats, att_masks, dict(opt, output_logsoftmax=1), _stats=(ent, lp))
        denom = (seq > 3).to(ent.dtype).sum(1) + 1          # VOCAB_LOWER = 3 (eval_utils.py:144)
        return seq, ent.sum(1) / denom, -lp.sum(1) / denom, pnum, plen, psyn, dt

    def _sample(self, fc_feats, att_feats, att_masks=None, opt={}, _stats=None):
        sample_method = opt.get("sample_method", "greedy")
        beam_size = opt.get("beam_size", 1)
        temperature = opt.get("temperature", 1.0)
        sample_n = int(opt.get("sample_n", 1))
        group_size = opt.get("group_size", 1)
        output_logsoftmax = opt.get("output_logsoftmax", 1)
        train_mode = opt.get("train_mode", "AIC")
        if (beam_size > 1 and sample_method in ("greedy", "beam_search")) or group_size > 1:
            raise NotImplementedError("beam / diverse decoding is undefined for the UIC model in the reference "
                                      "(EncoderDecoder_UIC has no decode()); use beam_size=1, group_size=1")
        if train_mode not in ("NAIC", "SAIC"):
            raise NotImplementedError("opt['train_mode'] must be 'NAIC' or 'SAIC' for the BoFi model, got %r" % train_mode)
        if not att_feats.is_cuda:
            raise RuntimeError("boficap_b200 runs on CUDA tensors only (no CPU fallback)")
        if self.training and torch.is_grad_enabled() and sample_method == "sample":
            return self._sample_with_tape(att_feats, att_masks, train_mode, sample_n, temperature, output_logsoftmax)
        eng = self.engine(att_feats.device, opt.get("bofi_precision"))
        att_len = None
        if att_masks is not None:
            att_len = eng.masks_to_len(att_masks.data)      # counts + device-side check that the masks are prefix masks
        if sample_method == "greedy":
            eng.set_sampling("greedy")
        elif sample_method == "sample":
            # the device draws the tokens (Gumbel-max, library stream); torch's CPU generator supplies the seed so that
            # torch.manual_seed makes a run reproducible
            eng.set_sampling("sample", temperature, int(torch.randint(0, 2 ** 31 - 1, (1,))))
        else:
            # CaptionModel.py:391-421: the 'gumbel' and 'top<k|p>' branches reduce over dim=1, which for the BoFi model's
            # [B, L, V] log-probs is the SLOT axis (they were written for the 2-D AIC log-probs): undefined for UIC in the reference
            raise NotImplementedError("sample_method=%r: only 'greedy' and 'sample' are defined for the BoFi model (the reference's "
                                      "gumbel / top-k / nucleus branches reduce over the slot axis of [B, L, V])" % sample_method)
        if att_feats.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            att_feats = att_feats.float()
        # `elapsed` (AttModel.py:337, :425: wall clock between two device synchronisations around the core) comes from two
        # events on the stream, so the call synchronises once, at the end, like any caller that reads the captions would
        stream = torch.cuda.current_stream(att_feats.device)
        eng.encode(att_feats, att_len)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        if _stats is not None:
            eng.set_decode_stats(*_stats)
        want_logprobs = _stats is None and bool(opt.get("bofi_logprobs", True))
        # NAIC log-probs: rows padded to a 16-byte pitch and returned as the [:, :, :V] view (TMA stores; engine.py:decode) unless
        # opt['bofi_dense_logprobs'] asks for the reference's contiguous tensor
        seq, logp, pnum, plen, psyn = eng.decode(train_mode, sample_n, output_logsoftmax, want_logprobs,
                                                 dense_logprobs=bool(opt.get("bofi_dense_logprobs", False)))
        eng.set_decode_stats(None, None)
        eng.set_sampling("greedy")
        ev1.record(stream)
        if att_masks is not None:
            eng.check_masks()                                # synchronises the stream; raises on a non-prefix mask
        else:
            ev1.synchronize()
        return seq, logp, pnum, plen, psyn, ev0.elapsed_time(ev1) * 1e-3
