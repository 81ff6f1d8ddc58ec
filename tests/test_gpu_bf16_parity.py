"""GPU: bf16 (tcgen05) engine against the ORACLE on the configurations bench.py runs, and on a trained-like checkpoint.

BASELINE.json north_star: bf16-mode logits within 2e-2 max-abs of the reference, agreement of the greedy captions
REPORTED.  A logit row is only comparable when the decoder saw the same input, i.e. when the image's boxes (phrase
lengths / syn labels) and the batch's fill window agree; box and token agreement are printed for every case.
  (a) NAIC, B = 1024, R = 36, calibration s_real      -- BASELINE config 2, what `bench.py` times
  (b) NAIC, B = 512, adaptive 10..100 regions          -- BASELINE config 3 per GPU (varlen encoder)
  (c) SAIC, calibration s_cap                          -- the eval default
  (d) the golden fixtures (reference-generated) in bf16
  (e) a trained-like checkpoint (boficap_b200/synth.py: trained_like_state_dict): margins, bf16-vs-fp32 token identity
"""
import numpy as np
import pytest
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig
from util import GOLDEN_CASES, checkpoint, golden_inputs, load_golden, oracle_for

pytestmark = pytest.mark.gpu
TOL = 2e-2          # BASELINE.json north_star: bf16-mode logits within 2e-2 max-abs


def _engine(cfg, sd, precision):
    from boficap_b200.engine import BofiEngine
    return BofiEngine(cfg, 0, precision).load_state_dict(sd)


def _run(eng, att, masks, mode="NAIC", logsoftmax=0, dtype=None):
    att_len = masks.long().sum(1).int().cuda() if masks is not None else None
    x = att.cuda()
    if dtype is not None:
        x = x.to(dtype)
    eng.encode(x, att_len)
    out = eng.decode(mode, 1, logsoftmax, True)
    torch.cuda.synchronize()
    return [t.cpu() for t in out]


def _robust_last(e16, e32, att, masks):
    """Boxes do not depend on the batch composition, the fill window does (last[B-1], TransformerModel.py:1871-1873): put an
    image last whose boxes the two precisions agree on and that has at least one phrase, so the windows coincide."""
    a, b = _run(e16, att, masks), _run(e32, att, masks)
    same = (a[3] == b[3]).all(1) & (a[4] == b[4]).all(1) & (b[3].sum(1) > 0)
    j = int(same.nonzero()[-1])
    B = att.shape[0]
    if j != B - 1:
        att = att.clone()
        att[[j, B - 1]] = att[[B - 1, j]]
        if masks is not None:
            masks = masks.clone()
            masks[[j, B - 1]] = masks[[B - 1, j]]
    return att, masks


def _report(tag, got, ref, capsys, slots=None):
    """got / ref: (seq, logits, pnum, plen, psyn).  Returns (box agreement, token identity, max |dlogit| on comparable rows)."""
    same_boxes = (got[3] == ref[3]).all(1) & (got[4] == ref[4]).all(1)
    same_window = bool(got[3][-1].sum() == ref[3][-1].sum())
    tok_rows = (got[0] == ref[0]).all(1).float().mean().item()
    tok_slots = (got[0] == ref[0]).float().mean().item()
    err = float("nan")
    if same_window and same_boxes.any():
        d = (got[1][same_boxes] - ref[1][same_boxes]).abs()
        if slots is not None:
            d = d[slots[same_boxes]]
        d = d[~torch.isnan(d)]
        err = d.max().item() if d.numel() else 0.0
    with capsys.disabled():
        print("\n[bf16 vs oracle | %s] boxes agree on %.1f%% of images, captions identical on %.1f%% (tokens %.2f%%), same fill window: %s, "
              "max |dlogit| on box-agreeing images %.4f" % (tag, 100 * same_boxes.float().mean().item(), 100 * tok_rows, 100 * tok_slots,
                                                            same_window, err))
    return same_boxes.float().mean().item(), tok_rows, err, same_window


def test_bf16_bench_config_b1024_r36_s_real_vs_oracle(capsys):
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_real")
    e16, e32 = _engine(cfg, sd, "bf16"), _engine(cfg, sd, "fp32")
    fc, att, _ = synth.synth_inputs(1024, 36, seed=1)
    att, _ = _robust_last(e16, e32, att, None)
    got = _run(e16, att, None)
    got_bf16_feats = _run(e16, att, None, dtype=torch.bfloat16)       # the feeder's format: read in place by the att_embed GEMM
    for x, y in zip(got, got_bf16_feats):
        assert torch.equal(torch.nan_to_num(x), torch.nan_to_num(y)), "bf16 features must give what fp32 features cast on the device give"
    ref = oracle_for(cfg, sd).sample(fc, att, None, {"train_mode": "NAIC", "output_logsoftmax": 0})
    f32 = _run(e32, att, None)
    assert torch.equal(f32[0], ref[0]) and torch.equal(f32[3], ref[3]) and torch.equal(f32[4], ref[4])     # fp32 engine == oracle at B = 1024
    boxes, toks, err, window = _report("B=1024 R=36 s_real NAIC", got, ref[:5], capsys)
    # random-weight checkpoints decide many boxes by margins of ~1e-2 (measured: 82.6 % of the images keep their boxes in bf16,
    # 96 % of the tokens; the trained-like checkpoint below keeps 100 %): the gate is on the logits, agreement is reported
    assert window and boxes > 0.7
    assert err < TOL, err
    e16.close()
    e32.close()


def test_bf16_adaptive_r100_b512_vs_oracle(capsys):
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_real")
    e16, e32 = _engine(cfg, sd, "bf16"), _engine(cfg, sd, "fp32")
    fc, att, masks = synth.synth_inputs(512, 100, seed=1, adaptive=True)
    att, masks = _robust_last(e16, e32, att, masks)
    got = _run(e16, att, masks)
    ref = oracle_for(cfg, sd).sample(fc, att, masks, {"train_mode": "NAIC", "output_logsoftmax": 0})
    f32 = _run(e32, att, masks)
    assert torch.equal(f32[0], ref[0]) and torch.equal(f32[3], ref[3]) and torch.equal(f32[4], ref[4])
    assert (f32[1] - ref[1]).abs().max().item() < 1e-4                 # varlen encoder, fp32: the reference's values
    boxes, toks, err, window = _report("B=512 adaptive 10..100 s_real NAIC", got, ref[:5], capsys)
    assert window and boxes > 0.7                                       # measured 77.9 % (random-weight margins, see above)
    assert err < TOL, err
    e16.close()
    e32.close()


def test_bf16_saic_s_cap_vs_oracle(capsys):
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_cap")
    e16 = _engine(cfg, sd, "bf16")
    fc, att, _ = synth.synth_inputs(64, 36, seed=7)
    got = _run(e16, att, None, mode="SAIC")
    ref = oracle_for(cfg, sd).sample(fc, att, None, {"train_mode": "SAIC", "output_logsoftmax": 0})
    # SAIC is autoregressive over phrases: a slot's logits are comparable while every earlier token agrees, i.e. up to and
    # including the phrase of the first token mismatch -- conservatively: the slots before the first mismatch and that slot
    L = got[0].shape[1]
    neq = (got[0] != ref[0])
    first = torch.where(neq.any(1), neq.float().argmax(1), torch.full((neq.shape[0],), L - 1))
    slots = torch.arange(L)[None, :] <= first[:, None]
    slots &= (ref[1].abs().sum(2) != 0) & (got[1].abs().sum(2) != 0)    # uncommitted slots stay zero in both
    boxes, toks, err, _ = _report("B=64 R=36 s_cap SAIC", got, ref[:5], capsys, slots=slots)
    assert boxes > 0.5
    assert err < TOL, err
    e16.close()


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if n.startswith("naic") and "nan" not in n])
def test_bf16_against_golden_fixtures(name, capsys):
    """The reference-generated fixtures (oracle/make_golden.py) hold the first columns of the raw logits: bf16 engine within
    2e-2 on the images whose boxes match the reference's."""
    fix, cfg = load_golden(name)
    sd = checkpoint(cfg, str(fix["calib"]))
    e16 = _engine(cfg, sd, "bf16")
    fc, att, masks = golden_inputs(fix)
    got = _run(e16, att, masks)
    same = torch.from_numpy((got[3].numpy() == fix["phrase_length"]).all(1) & (got[4].numpy() == fix["phrase_syn"]).all(1))
    window = bool(got[3][-1].sum().item() == fix["phrase_length"][-1].sum())
    k = fix["logits_head"].shape[2]
    with capsys.disabled():
        print("\n[bf16 vs golden | %s] boxes agree on %d / %d images, same fill window: %s" % (name, int(same.sum()), same.numel(), window))
    if window and same.any():
        d = np.abs(got[1][:, :, :k].numpy()[same.numpy()] - fix["logits_head"][same.numpy()])
        err = float(np.nanmax(d))
        with capsys.disabled():
            print("[bf16 vs golden | %s] max |dlogit| %.4f" % (name, err))
        assert err < TOL, err
    e16.close()


def test_trained_like_checkpoint_margins_and_bf16_token_identity(capsys):
    """Random-weight argmax margins are ~1e-2, so token agreement on the synthetic checkpoints says little about bf16.
    This builds a checkpoint with trained-like margins by over-fitting the library's own XE step on a fixed batch
    (recipe + seed committed in boficap_b200/synth.py, not the weights) and reports, on the memorised images:
    fp32 engine == oracle, the top-1 / top-2 margin quantiles, and bf16-vs-fp32 caption identity."""
    cfg = BofiConfig()
    logs = []
    sd, info = synth.trained_like_state_dict(cfg, seed=0, images=64, regions=36, log=logs.append)
    att, bt = info["att"], info["batch"]
    # longest caption last: its fill window covers every row (TransformerModel.py:1871-1873)
    total = bt["phrase_length"][:, 0, 1:].sum(1)
    order = torch.argsort(total, stable=True)
    att, labels, total = att[order], bt["labels"][order, 0], total[order]
    e16, e32 = _engine(cfg, sd, "bf16"), _engine(cfg, sd, "fp32")
    a, b = _run(e16, att, None, logsoftmax=1), _run(e32, att, None, logsoftmax=1)
    ref = oracle_for(cfg, sd).sample(None, att, None, {"train_mode": "NAIC"})
    assert torch.equal(b[0], ref[0]) and torch.equal(b[3], ref[3]) and torch.equal(b[4], ref[4]), "fp32 engine != oracle on the trained-like checkpoint"
    assert (b[1] - ref[1]).abs().max().item() < 1e-3
    L = b[0].shape[1]
    filled = torch.arange(L)[None, :] < b[3].sum(1)[:, None]
    top2 = b[1].topk(2, dim=2).values
    margin = (top2[..., 0] - top2[..., 1])[filled]
    q = torch.quantile(margin, torch.tensor([0.01, 0.05, 0.25, 0.5, 0.75]))
    hist = torch.histc(margin.clamp(max=20.0), bins=8, min=0.0, max=20.0).int().tolist()
    Lt = labels.shape[1] - 2
    want = torch.zeros(att.shape[0], L, dtype=torch.long)
    want[:, :Lt] = labels[:, 1:1 + Lt]
    memorised = ((b[0] == want) | ~filled).all(1).float().mean().item()
    boxes = ((a[3] == b[3]).all(1) & (a[4] == b[4]).all(1)).float().mean().item()
    ident = (a[0] == b[0]).all(1).float().mean().item()
    with capsys.disabled():
        print("\n[trained-like] %d XE steps, last losses %s" % (info["steps"], ["%.3f" % v for v in info["curve"][-1][1]]))
        print("[trained-like] captions reproduced from memory (fp32): %.1f%% of images; filled slots: %d" % (100 * memorised, int(filled.sum())))
        print("[trained-like] top-1 - top-2 log-prob margin quantiles 1%%/5%%/25%%/50%%/75%%: %s ; histogram 0..20 in 8 bins: %s"
              % (["%.2f" % v for v in q.tolist()], hist))
        print("[trained-like] bf16 vs fp32: boxes agree on %.1f%% of images, captions identical on %.1f%%, tokens identical on %.2f%%"
              % (100 * boxes, 100 * ident, 100 * (a[0] == b[0]).float().mean().item()))
    assert float(q[3]) > 1.0, "the checkpoint did not reach trained-like margins"
    assert ident >= 0.9, ident
    e16.close()
    e32.close()
