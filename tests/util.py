import json
import os

import numpy as np
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith(("naic", "saic")))   # xe_*: test_oracle_golden_xe.py, collate_*: test_data_pipeline.py
_SD_CACHE = {}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    fix = {k: z[k] for k in z.files}
    cfg = BofiConfig(**json.loads(str(fix["cfg"])))
    return fix, cfg


def checkpoint(cfg, calib, seed=0):
    key = (json.dumps(cfg.to_dict(), sort_keys=True), calib, seed)
    if key not in _SD_CACHE:
        if len(_SD_CACHE) > 2:
            _SD_CACHE.clear()
        _SD_CACHE[key] = synth.synth_state_dict(cfg, seed, calib)
    return _SD_CACHE[key]


def golden_inputs(fix):
    return synth.synth_inputs(int(fix["B"]), int(fix["R"]), seed=int(fix["input_seed"]), adaptive=bool(fix["adaptive"]))


def oracle_for(cfg, sd, record=False):
    from oracle.bofi_oracle import BofiOracle, OracleConfig
    return BofiOracle(sd, OracleConfig(**cfg.to_dict()), record=record)


def assert_close_nan(a, b, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert (na == nb).all(), "%s: NaN pattern differs (%d vs %d NaNs)" % (what, na.sum(), nb.sum())
    if (~na).any():
        err = np.abs(a[~na] - b[~na]).max()
        assert err <= atol, "%s: max abs err %.3e > %.1e" % (what, err, atol)


def check_against_golden(fix, seq, logp, pnum, plen, psyn, logits=None, atol=1e-4, tokens_exact=True):
    """seq/logp/... are torch CPU tensors produced by the implementation under test."""
    assert np.array_equal(pnum.numpy(), fix["phrase_num"]), "phrase_num"
    assert np.array_equal(plen.numpy(), fix["phrase_length"]), "phrase_length"
    assert np.array_equal(psyn.numpy(), fix["phrase_syn"]), "phrase_syn"
    if tokens_exact:
        assert np.array_equal(seq.numpy(), fix["seq"]), "seq"
    k = fix["logp_head"].shape[2]
    assert_close_nan(logp[:, :, :k].numpy(), fix["logp_head"], atol, "logp_head")
    assert_close_nan(logp.max(2).values.numpy(), fix["logp_max"], atol, "logp_max")
    gseq = torch.from_numpy(fix["seq"])
    assert_close_nan(logp.gather(2, gseq.unsqueeze(2)).squeeze(2).numpy(), fix["logp_at_seq"], atol, "logp_at_seq")
    if logits is not None:
        assert_close_nan(logits[:, :, :k].numpy(), fix["logits_head"], atol, "logits_head")
        assert_close_nan(logits.max(2).values.numpy(), fix["logits_max"], atol, "logits_max")
        assert_close_nan(torch.logsumexp(logits, 2).numpy(), fix["logits_lse"], atol, "logits_lse")
