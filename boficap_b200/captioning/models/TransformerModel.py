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
import time

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

    # ---- engine management -----------------------------------------------------------------
    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, device=None, precision=None):
        precision = precision or self.precision
        device = torch.device(device if device is not None else "cuda")
        index = device.index if device.index is not None else torch.cuda.current_device()
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

    def _forward(self, *args, **kwargs):
        raise NotImplementedError("the XE-training forward (TransformerModel._forward) is outside this round's "
                                  "hot path; use the reference model for training")

    def _sample(self, fc_feats, att_feats, att_masks=None, opt={}):
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
        eng = self.engine(att_feats.device, opt.get("bofi_precision"))
        att_len = None
        if att_masks is not None:
            att_len = att_masks.data.long().sum(1).to(torch.int32)
        eng.encode(att_feats.float(), att_len)
        torch.cuda.synchronize(att_feats.device)
        start = time.time()
        seq, logp, pnum, plen, psyn = eng.decode(train_mode, sample_n, output_logsoftmax, True)
        if sample_method != "greedy":
            if train_mode != "NAIC":
                raise NotImplementedError("multinomial sampling is only wired for NAIC in this round")
            lp = logp / temperature
            lp = lp.masked_fill(lp.isnan(), -10.0)
            seq = torch.distributions.Categorical(logits=lp).sample()
            total = plen.sum(1, keepdim=True)
            seq = seq.masked_fill(torch.arange(self.seq_length, device=seq.device)[None, :] >= total, self.pad_idx)
        torch.cuda.synchronize(att_feats.device)
        return seq, logp, pnum, plen, psyn, time.time() - start
