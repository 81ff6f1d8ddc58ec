"""CPU: checkpoint layout, drop-in module tree, synthetic checkpoint determinism and the C ABI surface."""
import argparse
import os
import re

import pytest
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig, state_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _opt(**kw):
    o = argparse.Namespace(caption_model="transformer", vocab_size=9487, input_encoding_size=512, rnn_size=2048, num_layers=6,
                           drop_prob_lm=0.5, seq_length=20, max_length=20, fc_feat_size=2048, att_feat_size=2048,
                           att_hid_size=512, use_bn=0, logit_layers=1, train_mode="UIC", decoder_input_mode="add",
                           N_enc=6, N_dec=6, N_len=1, d_model=512, d_ff=2048, num_att_heads=8, dropout=0.1,
                           vocab={str(i): "w%d" % i for i in range(4, 9491)})
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def test_state_spec_counts_match_survey_appendix_b():
    spec = state_spec(BofiConfig())
    assert len(spec) == 311
    n_param = sum(int(torch.Size(s).numel()) for n, (s, k) in spec.items() if k != "pe")
    assert n_param == 62384049
    assert spec["model.pos_embed.pe"][0] == (1, 5000, 512)
    assert len(state_spec(BofiConfig(N_len=0))) == 287 and len(state_spec(BofiConfig(N_len=2))) == 337


def test_dropin_module_has_reference_state_dict_layout():
    from boficap_b200.captioning import models
    m = models.setup(_opt())
    sd = m.state_dict()
    spec = state_spec(BofiConfig())
    assert list(sd.keys()) == list(spec.keys())
    assert all(tuple(sd[k].shape) == spec[k][0] for k in sd)
    assert [n for n, _ in m.named_buffers()] == ["model.pos_embed.pe"]
    # a reference-format checkpoint round-trips through torch.save / load_state_dict
    ck = synth.synth_state_dict(BofiConfig(), 0, "s_real")
    assert m.load_state_dict(ck).missing_keys == []
    with pytest.raises(Exception):
        models.setup(_opt(caption_model="fc"))
    with pytest.raises(NotImplementedError):
        models.setup(_opt(train_mode="AIC"))


def test_dropin_sample_fails_loudly_without_cuda():
    from boficap_b200.captioning import models
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = models.setup(_opt())
    fc, att, _ = synth.synth_inputs(2, 36)
    with pytest.raises(RuntimeError):
        m(fc, att, None, opt={"train_mode": "NAIC"}, mode="sample")


def test_synth_checkpoint_is_deterministic_and_calibrated():
    a = synth.synth_state_dict(BofiConfig(), 0, "s_real")
    b = synth.synth_state_dict(BofiConfig(), 0, "s_real")
    assert all(torch.equal(a[k], b[k]) for k in a)
    raw = synth.synth_state_dict(BofiConfig(), 0)
    k = "model.length_predictor.Length_classifier2.weight"
    assert not torch.equal(raw[k], a[k])
    with pytest.raises(KeyError):
        synth.synth_state_dict(BofiConfig(N_enc=3), 0, "s_real")


def test_synth_inputs_shapes():
    fc, att, m = synth.synth_inputs(5, 50, seed=3, adaptive=True)
    assert att.shape == (5, 50, 2048) and fc.shape == (5, 2048) and m.shape == (5, 50)
    n = m.sum(1).long()
    assert n[0] == 50 and (n >= 10).all()
    assert (att[1, n[1]:] == 0).all()
    assert (att >= 0).all()


def test_shared_library_exports_every_symbol_of_the_header():
    from boficap_b200 import _lib
    header = open(os.path.join(ROOT, "include", "bofi_b200.h")).read()
    declared = set(re.findall(r"\b(bofi_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = _lib.load()              # raises if the .so is missing or lacks a symbol
    assert lib.bofi_abi_version() == _lib.ABI_VERSION
    for name in declared:
        assert hasattr(lib, name)


def test_create_without_gpu_reports_cuda_error():
    import ctypes as C
    from boficap_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    lib = _lib.load()
    cfg = BofiConfig()
    c = _lib.BofiConfigC(abi_version=_lib.ABI_VERSION, tgt_vocab=cfg.tgt_vocab, att_feat_size=2048, n_enc=6, n_dec=6, n_len=1, d_model=512,
                         d_ff=2048, heads=8, seq_length=20, pad_idx=0, bos_idx=1, eos_idx=2, len_idx=3, precision=0)
    h = C.c_void_p()
    rc = lib.bofi_create(C.byref(c), 0, C.byref(h))
    assert rc == _lib.ERR_CUDA and b"no CPU fallback" in lib.bofi_last_error()
