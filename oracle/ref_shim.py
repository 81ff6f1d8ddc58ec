"""Import shim for the UNMODIFIED reference model class (container only).

TEST INFRASTRUCTURE -- never imported by the product path.  /root/reference does not exist
on the GPU box; this module is only used by oracle/make_golden.py (run here, outputs committed
under tests/golden/) and by CPU tests that are skipped when the reference tree is absent.

What is shimmed:
  * `turtle`, `thop`: unused top-level imports of captioning/models/TransformerModel.py:15,22
  * torch.cuda.synchronize: called unconditionally in AttModel.py:337,425 -> no-op without CUDA
"""
import argparse
import os
import sys
import types

import torch
import yaml

REF_ROOT = os.environ.get("BOFI_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "captioning", "models"))


def import_reference_models():
    for name, attrs in {"turtle": {"Turtle": object},
                        "thop": {"profile": lambda *a, **k: (0, 0)}}.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None
    import captioning.models as ref_models  # noqa: E402  (the reference package)
    return ref_models


def reference_opt(cfg_name="uic_sd.yml", vocab_size=9487, **overrides):
    """opts.py defaults the model reads + YAML overrides (opts.py:272-275 precedence)."""
    cfg = yaml.safe_load(open(os.path.join(REF_ROOT, "configs", cfg_name)))
    opt = argparse.Namespace(vocab_size=vocab_size, input_encoding_size=512, rnn_size=2048,
                             num_layers=6, drop_prob_lm=0.5, seq_length=20, fc_feat_size=2048,
                             att_feat_size=2048, att_hid_size=512, max_length=20, use_bn=0,
                             logit_layers=1)
    for k, v in cfg.items():
        setattr(opt, k, v)
    for k, v in overrides.items():
        setattr(opt, k, v)
    opt.vocab = {str(i): "w%d" % i for i in range(4, vocab_size + 4)}
    return opt


def build_reference_model(state_dict=None, cfg_name="uic_sd.yml", vocab_size=9487, **overrides):
    ref_models = import_reference_models()
    opt = reference_opt(cfg_name, vocab_size, **overrides)
    model = ref_models.setup(opt).eval()
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model, opt
