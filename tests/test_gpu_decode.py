"""GPU: the CUDA decode path (through the C ABI and through the drop-in model class) against the oracle
and the reference-generated golden vectors.  fp32 mode: boxes and greedy tokens identical, log-probs /
logits within 1e-4 max-abs (BASELINE.json north_star); bf16 mode: logits within 2e-2 where the boxes agree."""
import argparse

import numpy as np
import pytest
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig
from util import GOLDEN_CASES, assert_close_nan, check_against_golden, checkpoint, golden_inputs, load_golden, oracle_for

pytestmark = pytest.mark.gpu
NAIC_CASES = [n for n in GOLDEN_CASES if n.startswith("naic")]
_ENGINES = {}


def engine_for(cfg, calib, precision):
    from boficap_b200.engine import BofiEngine
    key = (str(cfg.to_dict()), calib, precision)
    if key not in _ENGINES:
        if len(_ENGINES) >= 3:
            for e in _ENGINES.values():
                e.close()
            _ENGINES.clear()
        _ENGINES[key] = BofiEngine(cfg, 0, precision).load_state_dict(checkpoint(cfg, calib))
    return _ENGINES[key]


def run_cuda(eng, att, masks, mode="NAIC", sample_n=1, output_logsoftmax=1):
    att_len = masks.long().sum(1).int().cuda() if masks is not None else None
    eng.encode(att.cuda(), att_len)
    out = eng.decode(mode, sample_n, output_logsoftmax, True)
    torch.cuda.synchronize()
    return [t.cpu() for t in out]


@pytest.mark.parametrize("name", NAIC_CASES)
def test_fp32_decode_matches_golden_and_oracle(name):
    fix, cfg = load_golden(name)
    calib = str(fix["calib"])
    eng = engine_for(cfg, calib, "fp32")
    fc, att, masks = golden_inputs(fix)
    seq, logp, pnum, plen, psyn = run_cuda(eng, att, masks)
    logits = run_cuda(eng, att, masks, output_logsoftmax=0)[1]
    check_against_golden(fix, seq, logp, pnum, plen, psyn, logits=logits, atol=1e-4)
    o = oracle_for(cfg, checkpoint(cfg, calib), record=True)
    ref = o.sample(fc, att, masks, {"train_mode": "NAIC"})
    assert torch.equal(seq, ref[0])
    assert_close_nan(logp.numpy(), ref[1].numpy(), 1e-4, "full log-prob tensor")
    info = eng.decode_info()
    assert info["bounding_steps"] == o.trace["steps"]
    assert info["fill_width"] == o.trace["fill_width"]
    assert info["nan_batch"] == int(bool(fix["nan"]))
    assert info["kernel_launches"] > 0


def test_fp32_saic_decode_matches_golden_and_oracle():
    fix, cfg = load_golden("saic_b8_r36")
    eng = engine_for(cfg, "s_cap", "fp32")
    fc, att, masks = golden_inputs(fix)
    seq, logp, pnum, plen, psyn = run_cuda(eng, att, masks, mode="SAIC")
    logits = run_cuda(eng, att, masks, mode="SAIC", output_logsoftmax=0)[1]
    check_against_golden(fix, seq, logp, pnum, plen, psyn, logits=None, atol=1e-4)
    o = oracle_for(cfg, checkpoint(cfg, "s_cap"))
    ref = o.sample(fc, att, masks, {"train_mode": "SAIC"})
    raw = oracle_for(cfg, checkpoint(cfg, "s_cap")).sample(fc, att, masks, {"train_mode": "SAIC", "output_logsoftmax": 0})
    assert torch.equal(seq, ref[0]) and torch.equal(pnum, ref[2]) and torch.equal(plen, ref[3]) and torch.equal(psyn, ref[4])
    assert_close_nan(logp.numpy(), ref[1].numpy(), 1e-4, "SAIC log-probs (zeros where unfilled)")
    assert_close_nan(logits.numpy(), raw[1].numpy(), 1e-4, "SAIC raw logits")
    assert eng.decode_info()["bounding_steps"] == o.last_steps


def test_saic_adaptive_regions_and_sample_n():
    cfg = BofiConfig()
    eng = engine_for(cfg, "s_cap", "fp32")
    fc, att, masks = synth.synth_inputs(5, 44, seed=13, adaptive=True)
    out = run_cuda(eng, att, masks, mode="SAIC", sample_n=2)
    ref = oracle_for(cfg, checkpoint(cfg, "s_cap")).sample(fc, att, masks, {"train_mode": "SAIC", "sample_n": 2})
    assert torch.equal(out[0], ref[0]) and torch.equal(out[3], ref[3]) and torch.equal(out[4], ref[4])
    assert_close_nan(out[1].numpy(), ref[1].numpy(), 1e-4, "logp")


def test_saic_phrase_nan_abort_matches_reference_behaviour():
    """A row that predicts EOS at the first step has an all-False phrase mask -> NaN -> the reference prints
    'phrase nan!' and returns the state as it is (TransformerModel.py:1956-1958)."""
    cfg = BofiConfig()
    eng = engine_for(cfg, "s_real", "fp32")          # s_real: several images stop at step 0
    fc, att, _ = synth.synth_inputs(16, 36, seed=7)
    out = run_cuda(eng, att, None, mode="SAIC")
    ref = oracle_for(cfg, checkpoint(cfg, "s_real")).sample(fc, att, None, {"train_mode": "SAIC"})
    assert (ref[0] == 0).all()                        # the oracle (== reference) aborted before any word
    for a, b in zip(out, ref[:5]):
        assert torch.equal(torch.nan_to_num(a.float()), torch.nan_to_num(b.float()))
    assert eng.decode_info()["nan_batch"] == 1


def test_encoder_memory_matches_oracle():
    cfg = BofiConfig()
    eng = engine_for(cfg, "s_real", "fp32")
    fc, att, masks = synth.synth_inputs(5, 50, seed=3, adaptive=True)
    mem = eng.encode(att.cuda(), masks.long().sum(1).int().cuda(), want_memory=True).cpu()
    o = oracle_for(cfg, checkpoint(cfg, "s_real"))
    x, m = o.prepare(att, masks)
    ref = o.encode(x, m)
    valid = masks.bool()
    assert (mem[valid] - ref[valid]).abs().max().item() < 1e-4


def test_dropin_model_sample_signature_and_parity():
    from boficap_b200.captioning import models
    cfg = BofiConfig()
    opt = synth.make_infos(cfg)["opt"]
    opt.vocab = synth.make_infos(cfg)["vocab"]
    model = models.setup(opt)
    model.load_state_dict(checkpoint(cfg, "s_real"))
    model = model.cuda().eval()
    fc, att, masks = synth.synth_inputs(6, 36, seed=7)
    out = model(fc.cuda(), att.cuda(), None, opt={"sample_method": "greedy", "beam_size": 1, "sample_n": 1, "train_mode": "NAIC"},
                mode="sample")
    assert len(out) == 6 and isinstance(out[5], float)
    seq, logp, pnum, plen, psyn = [t.cpu() for t in out[:5]]
    assert seq.dtype == torch.int64 and seq.shape == (6, 20)
    assert logp.dtype == torch.float32 and logp.shape == (6, 20, cfg.tgt_vocab)
    assert pnum.dtype == torch.int32 and plen.dtype == torch.int32 and psyn.dtype == torch.int64
    ref = oracle_for(cfg, checkpoint(cfg, "s_real")).sample(fc, att, None, {"train_mode": "NAIC"})
    assert torch.equal(seq, ref[0]) and torch.equal(pnum, ref[2]) and torch.equal(plen, ref[3]) and torch.equal(psyn, ref[4])
    assert_close_nan(logp.numpy(), ref[1].numpy(), 1e-4, "logp")
    with pytest.raises(NotImplementedError):
        model(fc.cuda(), att.cuda(), None, opt={"beam_size": 3, "train_mode": "NAIC"}, mode="sample")


def test_sample_n_repeats_rows():
    cfg = BofiConfig()
    eng = engine_for(cfg, "s_real", "fp32")
    fc, att, _ = synth.synth_inputs(4, 36, seed=7)
    one = run_cuda(eng, att, None)
    rep = run_cuda(eng, att, None, sample_n=3)
    ref = oracle_for(cfg, checkpoint(cfg, "s_real")).sample(fc, att, None, {"train_mode": "NAIC", "sample_n": 3})
    assert rep[0].shape == (12, 20)
    assert torch.equal(rep[0], ref[0]) and torch.equal(rep[3], ref[3])
    assert torch.equal(rep[2], one[2].repeat_interleave(3))


def test_host_entry_point_equals_device_path():
    cfg = BofiConfig()
    eng = engine_for(cfg, "s_real", "fp32")
    fc, att, masks = synth.synth_inputs(7, 40, seed=9, adaptive=True)
    dev = run_cuda(eng, att, masks)
    host = eng.sample_host(att.pin_memory(), masks.long().sum(1).int(), want_logprobs=True)
    assert torch.equal(host["seq"], dev[0]) and torch.equal(host["pnum"], dev[2])
    assert torch.equal(host["plen"], dev[3]) and torch.equal(host["psyn"], dev[4])
    assert torch.equal(torch.nan_to_num(host["logp"]), torch.nan_to_num(dev[1]))


def test_shard_invariance_at_bench_size():
    """Size-independent property at B=1024: images are independent except for the fill window taken from
    the LAST row (TransformerModel.py:1871-1873), so any shard that keeps the same last row reproduces
    the full batch's rows bit for bit."""
    cfg = BofiConfig()
    eng = engine_for(cfg, "s_real", "fp32")
    fc, att, _ = synth.synth_inputs(1024, 36, seed=1)
    full = run_cuda(eng, att, None)
    if int(full[3][-1].sum()) == 0:
        # an empty last row makes the whole batch NaN (w == 0); boxes do not depend on the batch, so move
        # an image that does produce phrases to the end and decode again
        j = int((full[3].sum(1) > 0).nonzero()[-1])
        att = att.clone()
        att[[j, 1023]] = att[[1023, j]]
        full = run_cuda(eng, att, None)
    assert not torch.isnan(full[1]).any()
    idx = torch.cat([torch.arange(100, 164), torch.tensor([1023])])
    part = run_cuda(eng, att[idx], None)
    for a, b in zip(part, full):
        assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b[idx]))
    # boxes never depend on the batch composition at all
    alone = run_cuda(eng, att[200:232], None)
    assert torch.equal(alone[3], full[3][200:232]) and torch.equal(alone[4], full[4][200:232])
    # sanity of the outputs at full size: token ids in range, zero beyond the caption, boxes consistent
    seq, logp, pnum, plen, psyn = full
    tot = plen.sum(1)
    assert ((seq >= 0) & (seq < cfg.tgt_vocab)).all() and (tot <= 20).all()
    pos = torch.arange(20)[None, :]
    assert (seq[pos >= tot[:, None]] == 0).all()
    assert ((plen > 0).sum(1) == pnum).all() and (((psyn >= 4) & (psyn <= 6)) == (plen > 0)).all()
    lse = torch.logsumexp(logp[:64], 2)
    assert lse.abs().max().item() < 1e-3


@pytest.mark.parametrize("gemm", ["tcgen05"])
def test_bf16_mode_logits_within_2e2_and_agreement_reported(gemm, capsys):
    cfg = BofiConfig()
    e16 = engine_for(cfg, "s_cap", "bf16")
    e32 = engine_for(cfg, "s_cap", "fp32")
    fc, att, _ = synth.synth_inputs(64, 36, seed=7)
    a = run_cuda(e16, att, None, output_logsoftmax=0)
    b = run_cuda(e32, att, None, output_logsoftmax=0)
    same_boxes = (a[3] == b[3]).all(1) & (a[4] == b[4]).all(1)
    same_window = bool(a[3][-1].sum() == b[3][-1].sum())
    tok_agree = (a[0] == b[0]).all(1).float().mean().item()
    with capsys.disabled():
        print("\n[bf16 vs fp32] boxes agree on %.1f%% of images, tokens identical on %.1f%%, same fill window: %s"
              % (100 * same_boxes.float().mean().item(), 100 * tok_agree, same_window))
    assert same_boxes.float().mean().item() > 0.5
    if same_window and same_boxes.any():
        err = (a[1][same_boxes] - b[1][same_boxes]).abs().max().item()
        with capsys.disabled():
            print("[bf16 vs fp32] max |dlogit| on box-agreeing images: %.4f" % err)
        assert err < 2e-2


def test_fast_len_row_bounding_equals_full_layer_formulation(monkeypatch):
    """N_len == 1: the [LEN]-row-only bounding step (tabulated self-attention K/V, M = rows GEMMs) must give
    bit-identical boxes/tokens/log-probs to the reference formulation that recomputes all 22 rows."""
    from boficap_b200.engine import BofiEngine
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_real")
    fc, att, masks = synth.synth_inputs(48, 40, seed=21, adaptive=True)
    for precision in ("fp32", "bf16"):
        fast = BofiEngine(cfg, 0, precision).load_state_dict(sd)
        monkeypatch.setenv("BOFI_BOUND", "generic")
        full = BofiEngine(cfg, 0, precision).load_state_dict(sd)
        monkeypatch.delenv("BOFI_BOUND")
        a = run_cuda(fast, att, masks)
        la = fast.decode_info()["kernel_launches"]
        b = run_cuda(full, att, masks)
        lb = full.decode_info()["kernel_launches"]
        if precision == "fp32":
            # every kernel of the fast path mirrors the arithmetic order of the full formulation: bit-identical
            for x, y in zip(a, b):
                assert torch.equal(torch.nan_to_num(x), torch.nan_to_num(y)), precision
        else:
            # bf16: the full formulation runs its 22-query attention on mma.sync, the [LEN]-row path on FFMA
            same = ((a[3] == b[3]).all(1) & (a[4] == b[4]).all(1)).float().mean().item()
            assert same >= 0.9, same
        assert la < lb
        fast.close()
        full.close()


def test_cluster_bounding_loop_matches_launch_chain(monkeypatch, capsys):
    """bf16 engine: the bounding loop as ONE cluster kernel (bound_loop.cuh: 8 CTAs per 64 rows, mma.sync GEMMs) against the
    chain of ~190 small launches it replaces (BOFI_BOUND_CLUSTER=0: tcgen05 GEMMs, separate LayerNorm / attention kernels).
    Same arithmetic up to the accumulation order: boxes agree on almost every image; where they do, tokens and logits agree
    like two bf16 runs do.  Ragged cluster (rows % 64 != 0), sample_n > 1 and padded regions are covered."""
    from boficap_b200.engine import BofiEngine
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_cap")
    monkeypatch.setenv("BOFI_BOUND_CLUSTER", "1")              # opt-in: measured slower than the launch chain (bound_loop.cuh)
    new = BofiEngine(cfg, 0, "bf16").load_state_dict(sd)
    monkeypatch.delenv("BOFI_BOUND_CLUSTER")
    old = BofiEngine(cfg, 0, "bf16").load_state_dict(sd)
    ref32 = engine_for(cfg, "s_cap", "fp32")
    for (B, R, adaptive, sn) in ((70, 36, False, 2), (200, 60, True, 1), (3, 36, False, 1)):
        fc, att, masks = synth.synth_inputs(B, R, seed=40 + B, adaptive=adaptive)
        a = run_cuda(new, att, masks, sample_n=sn, output_logsoftmax=0)
        la = new.decode_info()
        b = run_cuda(old, att, masks, sample_n=sn, output_logsoftmax=0)
        lb = old.decode_info()
        f = run_cuda(ref32, att, masks, sample_n=sn, output_logsoftmax=0)
        same = ((a[3] == b[3]).all(1) & (a[4] == b[4]).all(1))
        vs32_new = ((a[3] == f[3]).all(1) & (a[4] == f[4]).all(1)).float().mean().item()
        vs32_old = ((b[3] == f[3]).all(1) & (b[4] == f[4]).all(1)).float().mean().item()
        with capsys.disabled():
            print("\n[cluster bounding loop] B=%d R=%d sample_n=%d: boxes equal to the launch chain on %.1f%% of rows; vs fp32: cluster %.1f%%, chain %.1f%%; "
                  "launches %d vs %d, S %d vs %d" % (B, R, sn, 100 * same.float().mean().item(), 100 * vs32_new, 100 * vs32_old, la["kernel_launches"],
                                                     lb["kernel_launches"], la["bounding_steps"], lb["bounding_steps"]))
        assert same.float().mean().item() >= 0.9
        assert vs32_new >= vs32_old - 0.05                 # as close to the fp32 boxes as the chain is
        assert la["kernel_launches"] < lb["kernel_launches"] - 100
        if bool(a[3][-1].sum() == b[3][-1].sum()) and same.any():
            err = torch.nan_to_num(a[1][same] - b[1][same]).abs().max().item()
            assert err < 2e-2, err
    new.close()
    old.close()


def test_layernorm_epilogue_path_matches_two_launch_path(monkeypatch, capsys):
    """bf16 engine with BOFI_LNEPI=1: every residual GEMM on >= 2048 rows carries the following LayerNorm in its epilogue
    (gemm_tc2_ln.cuh).  The residual stream is bit-identical to the default path's (same fp32 sum); the normalised operand
    differs only where a bf16 rounding boundary is crossed, so boxes, tokens and logits agree like two bf16 runs of the same
    path.  Dense (B*R and B*20 rows) and varlen (device-side row count) encoders."""
    from boficap_b200.engine import BofiEngine
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_cap")
    monkeypatch.setenv("BOFI_LNEPI", "1")                      # opt-in: parity-green, measured slower than two launches
    new = BofiEngine(cfg, 0, "bf16").load_state_dict(sd)
    monkeypatch.delenv("BOFI_LNEPI")
    old = BofiEngine(cfg, 0, "bf16").load_state_dict(sd)
    for (B, R, adaptive) in ((128, 36, False), (160, 50, True)):
        fc, att, masks = synth.synth_inputs(B, R, seed=60 + B, adaptive=adaptive)
        a = run_cuda(new, att, masks, output_logsoftmax=0)
        la = new.decode_info()
        b = run_cuda(old, att, masks, output_logsoftmax=0)
        lb = old.decode_info()
        same = ((a[3] == b[3]).all(1) & (a[4] == b[4]).all(1))
        with capsys.disabled():
            print("\n[LayerNorm epilogue] B=%d R=%d: boxes equal on %.1f%% of rows, tokens on %.1f%%; launches %d vs %d"
                  % (B, R, 100 * same.float().mean().item(), 100 * (a[0] == b[0]).float().mean().item(), la["kernel_launches"], lb["kernel_launches"]))
        assert same.float().mean().item() >= 0.9
        assert la["kernel_launches"] < lb["kernel_launches"]
        if bool(a[3][-1].sum() == b[3][-1].sum()) and same.any():
            err = torch.nan_to_num(a[1][same] - b[1][same]).abs().max().item()
            assert err < 2e-2, err
    new.close()
    old.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_padded_logprob_pitch_equals_dense_tensor(precision):
    """bofi_decode_ex: log-prob rows on a 16-byte pitch (V = 9491 -> 9492 floats; the tcgen05 engine then writes them with TMA
    stores) hold exactly the values of the dense [rows, L, V] tensor; what comes back is the [:, :, :V] view."""
    cfg = BofiConfig()
    eng = engine_for(cfg, "s_cap", precision)
    for (B, R, sn, lsm) in ((130, 36, 1, 1), (24, 36, 2, 0)):
        fc, att, masks = synth.synth_inputs(B, R, seed=80 + B)
        eng.encode(att.cuda(), None)
        dense = eng.decode("NAIC", sn, lsm, True, dense_logprobs=True)
        eng.encode(att.cuda(), None)
        padded = eng.decode("NAIC", sn, lsm, True)
        torch.cuda.synchronize()
        assert dense[1].is_contiguous() and not padded[1].is_contiguous()
        assert padded[1].shape == dense[1].shape and padded[1].stride(1) == (cfg.tgt_vocab + 3) // 4 * 4
        assert torch.equal(padded[0], dense[0])
        assert torch.equal(torch.nan_to_num(padded[1]), torch.nan_to_num(dense[1]))
    # SAIC keeps the dense tensor
    eng.encode(att.cuda(), None)
    assert eng.decode("SAIC", 1, 1, True)[1].is_contiguous()


def _same(a, b):
    return torch.equal(torch.nan_to_num(a.float(), nan=-7.0), torch.nan_to_num(b.float(), nan=-7.0))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sharded_call_equals_separate_batches(precision):
    """bofi_set_shard: k batches decoded by ONE call, each with its own fill window w = last[end of its batch] - 1, give bit for
    bit what k separate `_sample` calls give -- incl. a batch whose window is empty (the reference's NaN batch,
    AttModel.py:427-428) next to healthy ones, sample_n > 1, a ragged last batch and adaptive regions."""
    fix, cfg = load_golden("naic_b2_r36_nanbatch")          # its last image predicts no phrase: w = 0 -> NaN log-probs
    eng = engine_for(cfg, str(fix["calib"]), precision)
    _, att_nan, _ = golden_inputs(fix)
    # (sizes at which a batch alone and the merged call take the same kernels: the library switches GEMM / vocabulary kernels at
    #  256 decode rows and 2048 GEMM rows, and those differ in the last bits of the log-probs; at B = 1024 both are far above)
    for (n, R, adaptive, sn) in ((2, 36, False, 1), (16, 36, False, 2), (14, 48, True, 1)):
        parts = []
        for k in range(3):
            _, att, masks = synth.synth_inputs(n if k < 2 else max(1, n - 1), R, seed=90 + 7 * k + n, adaptive=adaptive)
            parts.append((att, masks))
        if R == 36 and not adaptive and n == 2:
            parts[1] = (att_nan, None)                        # the NaN batch in the middle
        alone = [run_cuda(eng, a, m, sample_n=sn) for a, m in parts]
        nan_flags = [bool(torch.isnan(o[1]).any()) for o in alone]
        att_all = torch.cat([a for a, _ in parts]).cuda()
        len_all = torch.cat([m.long().sum(1) for _, m in parts]).int().cuda() if adaptive else None
        eng.set_shard(n)
        eng.encode(att_all, len_all)
        merged = [t.cpu() for t in eng.decode("NAIC", sn, 1, True)]
        info = eng.decode_info()
        eng.set_shard(0)
        lo = 0
        for (a, _), o in zip(parts, alone):
            hi = lo + a.shape[0] * sn
            for x, y in zip(o, merged):
                assert _same(x, y[lo:hi]), (precision, n, R, sn)
            lo = hi
        assert info["nan_batch"] == int(any(nan_flags))
        if n == 2 and R == 36:
            assert nan_flags[1]                               # the empty window is there, and (equality above) stays inside its batch


def test_grouped_pipeline_tickets_equal_standalone_decodes():
    """BofiPipeline(group=3): consecutive submissions share one library call (bofi_stage_part + bofi_set_shard); every ticket's
    slice equals the stand-alone decode of its batch -- device and pinned-host submissions, a group that is flushed incomplete,
    a change of batch shape in the middle of a group."""
    from boficap_b200.pipeline import BofiPipeline
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_cap")
    ref = engine_for(cfg, "s_cap", "bf16")
    pipe = BofiPipeline(cfg, sd, 0, "bf16", depth=2, group=3)
    batches = []
    for k, (B, R, adaptive) in enumerate(((40, 36, False),) * 4 + ((24, 36, False),) + ((32, 50, True),) * 2):
        _, att, masks = synth.synth_inputs(B, R, seed=120 + k, adaptive=adaptive)
        batches.append((att.to(torch.bfloat16), masks))
    want = []
    for att, masks in batches:
        ln = masks.long().sum(1).int().cuda() if masks is not None else None
        ref.encode(att.cuda(), ln)
        want.append([t.cpu() for t in ref.decode("NAIC", 1, 1, True)])
    torch.cuda.synchronize()
    # device submissions
    tickets = [pipe.submit_device(att.cuda(), masks.long().sum(1).int().cuda() if masks is not None else None, want_logprobs=True)
               for att, masks in batches]
    assert [t.launched for t in tickets] == [True, True, True, True, True, False, False]     # 3 | 1 (shape change) | 1 (shape change) | 2 still open
    for t, w in zip(tickets, want):
        got = t.wait()
        for x, y in zip(w, got):
            assert _same(x, y.cpu())
    # pinned-host submissions (results land in the slot's pinned buffers: consume a ticket before its slot is reused)
    for i in range(0, len(batches), 3):
        chunk = batches[i:i + 3]
        ts = [pipe.submit_host(att.pin_memory(), masks.long().sum(1).int() if masks is not None else None, want_logprobs=True)
              for att, masks in chunk]
        pipe.flush()
        for t, w in zip(ts, want[i:i + 3]):
            got = t.wait()
            for x, y in zip(w, (got["seq"], got["logp"], got["pnum"], got["plen"], got["psyn"])):
                assert _same(x, y)
    pipe.close()
