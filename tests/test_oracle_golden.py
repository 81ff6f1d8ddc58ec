"""CPU: the oracle (oracle/bofi_oracle.py) against the golden vectors that oracle/make_golden.py
recorded from the UNMODIFIED reference model (tests/golden/*.npz).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from util import GOLDEN_CASES, assert_close_nan, check_against_golden, checkpoint, golden_inputs, load_golden, oracle_for


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    fix, cfg = load_golden(name)
    sd = checkpoint(cfg, str(fix["calib"]))
    fc, att, masks = golden_inputs(fix)
    o = oracle_for(cfg, sd, record=True)
    mode = str(fix["mode"])
    seq, logp, pnum, plen, psyn, _ = o.sample(fc, att, masks, {"train_mode": mode})
    raw = oracle_for(cfg, sd).sample(fc, att, masks, {"train_mode": mode, "output_logsoftmax": 0})
    # the oracle is the same fp32 arithmetic as the reference on the same CPU kernels: 1e-6
    check_against_golden(fix, seq, logp, pnum, plen, psyn, logits=raw[1], atol=1e-6)
    mem = o.trace["memory"]
    assert_close_nan(mem[:, :, :16].numpy(), fix["memory_head"], 1e-6, "memory_head")
    assert_close_nan(mem.sum(2).numpy(), fix["memory_rowsum"], 1e-4, "memory_rowsum")
    assert bool(torch.isnan(logp).any()) == bool(fix["nan"])


def test_nan_batch_fixture_is_all_nan():
    fix, _ = load_golden("naic_b2_r36_nanbatch")
    assert bool(fix["nan"]) and np.isnan(fix["logp_head"]).all()
    assert (fix["seq"] == 0).all()


def test_fixtures_cover_the_edge_cases():
    kinds = {}
    for name in GOLDEN_CASES:
        fix, cfg = load_golden(name)
        kinds[name] = fix
    plen = kinds["naic_b16_r36"]["phrase_length"]
    tot = plen.sum(1)
    assert (tot == 0).any() and (tot == 20).any() and ((tot > 0) & (tot < 20)).any()   # EOS at step 0, clipped, mid
    assert bool(kinds["naic_b12_r50_adaptive"]["adaptive"])
