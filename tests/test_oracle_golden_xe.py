"""The oracle's XE-training forward / loss / autograd gradients against the golden vectors recorded from the
unmodified reference (oracle/make_golden_xe.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig
from oracle.bofi_oracle import BofiOracle, OracleConfig

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = ["xe_b2_r12", "xe_b3_r20_adaptive", "xe_b2_r12_nlen2"]      # (the last: two bounding layers, configs/uic_sd_N2.yml)
GLAT_CASES = ["xe_b3_r20_glat", "xe_b2_r12_glat1"]      # glancing training: forward_xe_fused(glat_p, glat_seed) vs the reference with the same uniforms


def run_oracle_xe(fix, requires_grad=True):
    cfg = BofiConfig(**(json.loads(str(fix["cfg_kw"])) if "cfg_kw" in fix else {}))
    sd = synth.synth_state_dict(cfg, 0, "s_real")
    if requires_grad:
        for k, v in sd.items():
            if k != "model.pos_embed.pe":
                v.requires_grad_(True)
    o = BofiOracle.__new__(BofiOracle)
    o.sd, o.cfg, o.record, o.trace = sd, OracleConfig(**cfg.to_dict()), False, {}
    B, R = int(fix["B"]), int(fix["R"])
    fc, att, masks = synth.synth_inputs(B, R, seed=int(fix["input_seed"]), adaptive=bool(fix["adaptive"]))
    bt = synth.synth_xe_batch(B, seed=int(fix["batch_seed"]), vocab_size=cfg.vocab_size)
    if "glat_p" in fix and float(fix["glat_p"]) >= 0:
        outs = o.forward_xe_fused(att, masks, bt["labels"], bt["phrase_num"], bt["phrase_length"], bt["extend_phrase_syn_seq"],
                                  bt["extend_phrase_seq"], bt["extend_phrase_seq_mask"], glat_p=float(fix["glat_p"]),
                                  glat_seed=int(fix["glat_seed"]))
    else:
        outs = o.forward_xe(att, masks, bt["labels"], bt["phrase_num"], bt["phrase_length"], bt["extend_phrase_syn_seq"],
                            bt["extend_phrase_seq"], bt["extend_phrase_seq_mask"])
    loss, parts = o.loss_xe(outs, bt["phrase_num"], bt["phrase_length"], bt["phrase_syn"], bt["labels"])
    return sd, bt, outs, loss, parts


@pytest.mark.parametrize("name", CASES + GLAT_CASES)
def test_oracle_xe_matches_reference(name):
    fix = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd, bt, outs, loss, parts = run_oracle_xe(fix)
    sa_len, sa_syn, sa_logp, na_len, na_syn, na_logp = [t.detach() for t in outs]
    for got, key in ((sa_len, "sa_len"), (sa_syn, "sa_syn"), (na_len, "na_len"), (na_syn, "na_syn")):
        np.testing.assert_allclose(got.numpy(), fix[key], atol=2e-5, rtol=0)
    words = bt["labels"].reshape(-1, bt["labels"].shape[2])[:, 1:-1]
    for tag, lp in (("sa", sa_logp), ("na", na_logp)):
        np.testing.assert_allclose(lp[:, :, :fix[tag + "_logp_head"].shape[2]].numpy(), fix[tag + "_logp_head"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(lp.max(2).values.numpy(), fix[tag + "_logp_max"], atol=2e-5, rtol=0)
        np.testing.assert_allclose(lp.gather(2, words.unsqueeze(2)).squeeze(2).numpy(), fix[tag + "_logp_at_label"], atol=2e-5, rtol=0)
    # reference order: total, SA_length, SA_phrase, SA_syn, NA_length, NA_phrase, NA_syn
    np.testing.assert_allclose([float(loss)] + [float(p) for p in parts], fix["losses"], rtol=1e-5)
    loss.backward()
    names = json.loads(str(fix["grad_names"]))
    for pname, ref_norm in zip(names, fix["grad_norms"]):
        g = sd[pname].grad
        g = torch.zeros_like(sd[pname]) if g is None else g
        ntol = 5e-4 if name in GLAT_CASES else 1e-4      # (glancing cases: the isolated ReLU flip described below)
        assert abs(float(g.double().norm()) - ref_norm) <= ntol * max(ref_norm, 1e-3) + 1e-7, pname
        np.testing.assert_allclose(g.reshape(-1)[:64].numpy(), fix["ghead/" + pname], atol=1e-5 + 1e-4 * float(np.abs(fix["ghead/" + pname]).max()), rtol=0, err_msg=pname)
        if "gfull/" + pname in fix:
            ref = fix["gfull/" + pname]
            bad = int((np.abs(g.numpy() - ref) > 1e-5 + 1e-4 * float(np.abs(ref).max())).sum())
            # the glancing cases go through forward_xe_fused (encoder once per image, batched bounding passes): a pre-activation
            # within rounding of zero may land on the other side of a ReLU -- one isolated entry of an FFN bias gradient
            assert bad <= (1 if name in GLAT_CASES else 0), (pname, bad)


def test_oracle_fused_formulation_equals_reference_formulation():
    """forward_xe_fused (the CUDA path's formulation; carries the reproducible dropout masks) == forward_xe at p = 0."""
    fix = np.load(os.path.join(GOLDEN, CASES[1] + ".npz"))
    cfg = BofiConfig()
    sd = synth.synth_state_dict(cfg, 0, "s_real")
    o = BofiOracle(sd, OracleConfig(**cfg.to_dict()))
    B, R = int(fix["B"]), int(fix["R"])
    fc, att, masks = synth.synth_inputs(B, R, seed=int(fix["input_seed"]), adaptive=bool(fix["adaptive"]))
    bt = synth.synth_xe_batch(B, seed=int(fix["batch_seed"]), vocab_size=cfg.vocab_size)
    a = (att, masks, bt["labels"], bt["phrase_num"], bt["phrase_length"], bt["extend_phrase_syn_seq"], bt["extend_phrase_seq"],
         bt["extend_phrase_seq_mask"])
    with torch.no_grad():
        for x, y in zip(o.forward_xe(*a), o.forward_xe_fused(*a)):
            assert float((x - y).abs().max()) < 5e-5
