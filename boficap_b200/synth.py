"""Seeded synthetic checkpoints and inputs (no trained weights ship with the reference:
model_example.pth / infos_example.pkl are git-LFS pointer stubs).

`synth_state_dict` draws a reference-layout state_dict (layout.state_spec) from a seeded CPU
generator: Xavier-uniform for every >1-D tensor (what make_model does, TransformerModel.py:1621-1623),
small uniform biases, mildly perturbed LayerNorm gains/offsets (so the affine part is exercised),
the closed-form sinusoid table for `pe` (TransformerModel.py:1496-1503).

A randomly initialised bounding head predicts EOS for every image.  `apply_calibration` rescales
the two classifier heads with constants that were fitted ONCE against the unmodified reference in
the build container (oracle/make_golden.py, SURVEY.md Appendix A recipe) and are committed in
boficap_b200/data/synth_calib.json, so that any box regenerates bit-identical weights without
running any model code.

`synth_inputs` follows SURVEY.md section 8(d): att_feats = rand(B,R,2048) (bottom-up features are
post-ReLU, non-negative), fc_feats = att_feats.mean(1); adaptive: n_i ~ randint(10,R+1), row 0 forced to
R, float32 prefix masks, features zeroed beyond n_i.
"""
import json
import math
import os
from collections import OrderedDict

import torch

from .layout import BofiConfig, state_spec

_CALIB_PATH = os.path.join(os.path.dirname(__file__), "data", "synth_calib.json")


def sinusoid_table(max_len, d):
    pe = torch.zeros(max_len, d)
    position = torch.arange(0, max_len).unsqueeze(1).float()
    div_term = torch.exp(torch.arange(0, d, 2).float() * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0)


def synth_state_dict(cfg=None, seed=0, calib=None):
    cfg = cfg or BofiConfig()
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, (shape, kind) in state_spec(cfg).items():
        if kind in ("matrix", "embedding"):
            fan_out, fan_in = shape
            a = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif kind == "bias":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "ones":
            t = 1.0 + 0.1 * (torch.rand(shape, generator=g) * 2 - 1)
        elif kind == "zeros":
            t = 0.05 * (torch.rand(shape, generator=g) * 2 - 1)
        elif kind == "pe":
            t = sinusoid_table(shape[1], shape[2])
        else:
            raise ValueError(kind)
        sd[name] = t.contiguous()
    if calib is not None:
        apply_calibration(sd, cfg, seed, calib)
    return sd


def calibration_key(cfg, seed, calib):
    return "%s/seed%d/Nenc%d_Ndec%d_Nlen%d_V%d" % (calib, seed, cfg.N_enc, cfg.N_dec, cfg.N_len, cfg.vocab_size)


def load_calibration_table():
    with open(_CALIB_PATH) as f:
        return json.load(f)


def apply_calibration(sd, cfg, seed, calib):
    table = load_calibration_table()
    key = calibration_key(cfg, seed, calib)
    if key not in table:
        raise KeyError("no committed calibration for %s (have: %s)" % (key, sorted(table)))
    entry = table[key]
    lp = "model.length_predictor."
    for head, name in (("len", "Length_classifier2"), ("syn", "Syntactic_classifier2")):
        sd[lp + name + ".weight"] = sd[lp + name + ".weight"] * float(entry[head + "_gain"])
        sd[lp + name + ".bias"] = torch.tensor(entry[head + "_bias"], dtype=torch.float32)
    # guard against RNG drift between torch builds: the constants only fit these exact weights
    probe = float(sd["model.length_predictor.Length_classifier1.weight"].double().abs().sum())
    if abs(probe - entry["probe"]) > 1e-6 * max(1.0, abs(entry["probe"])):
        raise RuntimeError("synthetic weights differ from the calibrated ones (probe %r vs %r)" % (probe, entry["probe"]))
    return sd


def synth_inputs(B, R=36, seed=1, adaptive=False, feat=2048, min_regions=10):
    g = torch.Generator().manual_seed(seed)
    att = torch.rand(B, R, feat, generator=g)
    masks = None
    if adaptive:
        n = torch.randint(min_regions, R + 1, (B,), generator=g)
        n[0] = R
        masks = (torch.arange(R)[None, :] < n[:, None]).float()
        att = att * masks[:, :, None]
    fc = att.mean(1)
    return fc, att, masks


def make_infos(cfg, opt_namespace=None):
    """`infos.pkl` payload in the reference's format (tools/train.py:55-69): the drop-in loads
    `infos['opt']` / `infos['vocab']` exactly as tools/eval.py:49-63 does."""
    import argparse
    vocab = {str(i): "w%d" % i for i in range(4, cfg.vocab_size + 4)}
    if opt_namespace is None:
        opt_namespace = argparse.Namespace(
            caption_model="transformer", vocab_size=cfg.vocab_size, input_encoding_size=cfg.d_model,
            rnn_size=cfg.d_ff, num_layers=cfg.N_enc, drop_prob_lm=cfg.drop_prob_lm, seq_length=cfg.seq_length,
            max_length=cfg.seq_length, fc_feat_size=2048, att_feat_size=cfg.att_feat_size, att_hid_size=512,
            use_bn=0, logit_layers=1, train_mode=cfg.train_mode, decoder_input_mode=cfg.decoder_input_mode,
            N_enc=cfg.N_enc, N_dec=cfg.N_dec, N_len=cfg.N_len, d_model=cfg.d_model, d_ff=cfg.d_ff,
            num_att_heads=cfg.h, dropout=cfg.dropout)
    return {"iter": 0, "epoch": 0, "vocab": vocab, "opt": opt_namespace, "best_val_score": None}


def synth_xe_batch(B, seq_per_img=5, seq_length=16, seed=3, vocab_size=9487, len_idx=3, bos_idx=1, eos_idx=2,
                   max_phrases=8):
    """Synthetic teacher-forcing batch with the tensors the reference's collate_func builds for train_mode UIC
    (captioning/data/dataloader.py:343-428): random phrase segmentations (syn labels 4..6, lengths 1..5, at most
    `seq_length` words), random words, and the derived decoder inputs:
      labels                [B, spi, L+2] i64   0, words..., 0 padding
      phrase_num            [B, spi]      i64   number of phrases + 1 (the BOS pseudo-phrase)
      phrase_length/_syn    [B, spi, L+2] i64   slot 0 = (1, bos); syn gets eos after the last phrase
      extend_phrase_syn_seq [B, spi, L+2] i64   slot 0 = len_idx, then the syn label of every word slot
      extend_phrase_seq     [B, spi, L]   i64   position-wise copy of the previous phrase's words
      extend_phrase_seq_mask[B, spi, L*L] bool  phrase-block-causal mask (prefix form)
    """
    import numpy as np
    rng = np.random.RandomState(seed)
    N, L = B * seq_per_img, seq_length
    labels = np.zeros((N, L + 2), dtype=np.int64)
    pnum = np.zeros(N, dtype=np.int64)
    plen = np.zeros((N, L + 2), dtype=np.int64)
    psyn = np.zeros((N, L + 2), dtype=np.int64)
    ext_syn = np.zeros((N, L + 2), dtype=np.int64)
    ext_seq = np.zeros((N, L), dtype=np.int64)
    ext_mask = np.zeros((N, L, L), dtype=bool)
    plen[:, 0] = 1
    psyn[:, 0] = bos_idx
    ext_syn[:, 0] = len_idx
    for n in range(N):
        lens, total = [], 0
        for _ in range(int(rng.randint(1, max_phrases + 1))):
            ln = int(rng.randint(1, 6))
            if total + ln > L:
                break
            lens.append(ln)
            total += ln
        k = len(lens)
        syns = rng.randint(4, 7, size=k)
        labels[n, 1:1 + total] = rng.randint(4, vocab_size + 4, size=total)
        pnum[n] = k + 1
        plen[n, 1:k + 1] = lens
        psyn[n, 1:k + 1] = syns
        psyn[n, k + 1] = eos_idx
        pos = 1
        for j in range(k):
            ext_syn[n, pos:pos + lens[j]] = syns[j]
            pos += lens[j]
        src, dst = 0, 0                      # start of the previous phrase in labels / of this phrase in the decoder input
        for j in range(1, k + 1):
            cur, prev = int(plen[n, j]), int(plen[n, j - 1])
            if cur <= prev:                  # keep the last `cur` words of the previous phrase
                ext_seq[n, dst:dst + cur] = labels[n, src + prev - cur:src + prev]
            else:                            # stretch: the first words ct times, the rest ct + 1 times
                few, ct = prev - cur % prev, cur // prev
                reps = [ct if q < few else ct + 1 for q in range(prev)]
                ext_seq[n, dst:dst + cur] = np.repeat(labels[n, src:src + prev], reps)
            ext_mask[n, dst:, :dst + cur] = True
            src += prev
            dst += cur
    t = lambda a, *shape: torch.from_numpy(a).reshape(B, seq_per_img, *shape)
    return dict(labels=t(labels, L + 2), phrase_num=t(pnum), phrase_length=t(plen, L + 2), phrase_syn=t(psyn, L + 2),
                extend_phrase_syn_seq=t(ext_syn, L + 2), extend_phrase_seq=t(ext_seq, L),
                extend_phrase_seq_mask=t(ext_mask, L * L))


def trained_like_state_dict(cfg=None, seed=0, images=64, regions=36, max_steps=1500, lr=5e-4, target_word_nll=0.05,
                            precision="bf16", device=0, log=None):
    """A checkpoint whose argmax margins look like a TRAINED model's (no trained weights ship with the reference, and the
    Xavier-initialised synthetic checkpoints decide most tokens by margins of ~1e-2, which says little about bf16).

    Recipe (needs a CUDA device; deterministic up to the fp32 atomics of the word-embedding gradient): start from
    `synth_state_dict(cfg, seed)`, and run the library's own fused XE step (`TransformerModel.xe_step`, eval() arithmetic:
    dropout off) with Adam(lr, betas=(0.9, 0.98), eps=1e-9, 20 warm-up steps) on ONE fixed synthetic batch --
    `synth_inputs(images, regions, seed + 101)` with one caption per image from `synth_xe_batch(images, 1, seed + 103)`,
    so that every image has a single caption to memorise -- until the mean word NLL of both decoder passes is below
    `target_word_nll` (or `max_steps`).  Returns (state_dict on the CPU, info) where info holds the batch, the step count
    and the loss curve; the weights are NOT committed, only this recipe."""
    from .captioning import models
    cfg = cfg or BofiConfig()
    infos = make_infos(cfg)
    opt = infos["opt"]
    opt.vocab = infos["vocab"]
    opt.bofi_precision = precision
    model = models.setup(opt)
    model.load_state_dict(synth_state_dict(cfg, seed, None))
    dev = torch.device("cuda", device)
    model = model.to(dev).eval()
    model.train_bind(dev)
    flat_w, flat_g = model.flat_params(), model.flat_grads()
    flat_p = torch.nn.Parameter(flat_w)
    flat_p.grad = flat_g
    optim = torch.optim.Adam([flat_p], lr=lr, betas=(0.9, 0.98), eps=1e-9)
    fc, att, masks = synth_inputs(images, regions, seed=seed + 101)
    bt = synth_xe_batch(images, seq_per_img=1, seed=seed + 103, vocab_size=cfg.vocab_size)
    d = lambda t: t.to(dev) if t is not None else None
    args = (d(fc), d(att), d(bt["labels"]), d(masks), d(bt["phrase_num"]), d(bt["phrase_length"]), d(bt["phrase_syn"]),
            d(bt["extend_phrase_syn_seq"]), d(bt["extend_phrase_seq"]), d(bt["extend_phrase_seq_mask"]))
    curve, steps = [], 0
    for it in range(max_steps):
        for g in optim.param_groups:
            g["lr"] = lr * min(1.0, (it + 1) / 20.0)
        flat_g.zero_()
        losses = model.xe_step(*args)
        optim.step()
        model._engine.refresh_weights()
        steps = it + 1
        if it % 25 == 0 or it == max_steps - 1:
            l = [float(v) for v in losses.cpu()]
            curve.append((it, l))
            if log:
                log("trained-like step %d: total %.3f SA(len %.3f word %.3f syn %.3f) NA(len %.3f word %.3f syn %.3f)" % ((it,) + tuple(l)))
            if max(l[2], l[5]) < target_word_nll and max(l[1], l[3], l[4], l[6]) < 4 * target_word_nll:
                break
    sd = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    model._engine.close()
    return sd, dict(att=att, masks=masks, batch=bt, steps=steps, curve=curve)
