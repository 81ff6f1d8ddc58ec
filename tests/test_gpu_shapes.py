"""GPU: batch / region-count sweep (BASELINE.json configs 3 and 4: B in {1, 32, 512}, R up to 100, padded masks)
of the fp32 CUDA path against the oracle."""
import pytest
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig
from util import assert_close_nan, checkpoint, oracle_for

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from boficap_b200.engine import BofiEngine
    cfg = BofiConfig()
    e = BofiEngine(cfg, 0, "fp32").load_state_dict(checkpoint(cfg, "s_cap"))
    yield e
    e.close()


@pytest.mark.parametrize("B,R,adaptive", [(1, 36, False), (1, 100, False), (3, 100, True), (32, 36, False), (32, 100, True),
                                           (5, 10, False), (130, 37, True)])
def test_shape_sweep_fp32_vs_oracle(eng, B, R, adaptive):
    cfg = BofiConfig()
    fc, att, masks = synth.synth_inputs(B, R, seed=100 + B + R, adaptive=adaptive)
    att_len = masks.long().sum(1).int().cuda() if masks is not None else None
    eng.encode(att.cuda(), att_len)
    out = [t.cpu() for t in eng.decode("NAIC", 1, 1, True)]
    ref = oracle_for(cfg, checkpoint(cfg, "s_cap")).sample(fc, att, masks, {"train_mode": "NAIC"})
    assert torch.equal(out[2], ref[2]) and torch.equal(out[3], ref[3]) and torch.equal(out[4], ref[4])
    assert torch.equal(out[0], ref[0])
    assert_close_nan(out[1].numpy(), ref[1].numpy(), 1e-4, "logp")


def test_batch_512_matches_its_own_shards(eng):
    cfg = BofiConfig()
    fc, att, masks = synth.synth_inputs(512, 100, seed=5, adaptive=True)
    att_len = masks.long().sum(1).int().cuda()
    eng.encode(att.cuda(), att_len)
    full = [t.cpu() for t in eng.decode("NAIC", 1, 1, False)[:1] + eng.decode("NAIC", 1, 1, False)[2:]]
    idx = torch.cat([torch.arange(17, 49), torch.tensor([511])])
    eng.encode(att[idx].cuda(), att_len[idx.cuda()])
    o = eng.decode("NAIC", 1, 1, False)
    part = [t.cpu() for t in o[:1] + o[2:]]
    for a, b in zip(part, full):
        assert torch.equal(a, b[idx])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_varlen_encoder_equals_padded_formulation(precision, monkeypatch):
    """the compact (varlen) encoder must return what the padded one returns -- bit for bit in fp32 (a row's
    arithmetic does not depend on where the row sits), same boxes / tokens and 1e-2-close logits in bf16."""
    from boficap_b200.engine import BofiEngine
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_real")
    fc, att, masks = synth.synth_inputs(40, 100, seed=17, adaptive=True)
    att_len = masks.long().sum(1).int().cuda()
    var = BofiEngine(cfg, 0, precision).load_state_dict(sd)
    monkeypatch.setenv("BOFI_VARLEN", "0")
    pad = BofiEngine(cfg, 0, precision).load_state_dict(sd)
    monkeypatch.delenv("BOFI_VARLEN")
    res = []
    for eng in (var, pad):
        mem = eng.encode(att.cuda(), att_len, want_memory=True)
        out = eng.decode("NAIC", 1, 0, True)
        torch.cuda.synchronize()
        res.append((mem.cpu(), [t.cpu() for t in out], eng.decode_info()["kernel_launches"]))
    (mem_v, out_v, _), (mem_p, out_p, _) = res
    valid = masks.bool()
    if precision == "fp32":
        assert torch.equal(mem_v[valid], mem_p[valid])
        for a, b in zip(out_v, out_p):
            assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
    else:
        assert (mem_v[valid] - mem_p[valid]).abs().max().item() < 5e-2
        assert torch.equal(out_v[3], out_p[3]) and torch.equal(out_v[4], out_p[4])
    assert (mem_v[~valid] == 0).all()                      # padded rows of the returned memory are zero-filled
    # SAIC through the compact memory as well
    for mode_eng in (var, pad):
        mode_eng.encode(att.cuda(), att_len)
    a = var.decode("SAIC", 1, 1, True)
    b = pad.decode("SAIC", 1, 1, True)
    torch.cuda.synchronize()
    if precision == "fp32":
        for x, y in zip(a, b):
            assert torch.equal(torch.nan_to_num(x.float()), torch.nan_to_num(y.float()))
    var.close()
    pad.close()


def test_non_prefix_mask_is_reported():
    """The kernels keep only the count of valid regions; the reference's masked_fill accepts any mask.  A mask with a hole
    must raise instead of silently decoding something else (device-side check, read at the call's own synchronisation)."""
    from boficap_b200.captioning import models
    cfg = BofiConfig()
    infos = synth.make_infos(cfg)
    opt = infos["opt"]
    opt.vocab = infos["vocab"]
    model = models.setup(opt)
    model.load_state_dict(checkpoint(cfg, "s_real"))
    model = model.cuda().eval()
    fc, att, masks = synth.synth_inputs(6, 40, seed=3, adaptive=True)
    kw = {"sample_method": "greedy", "train_mode": "NAIC"}
    model(fc.cuda(), att.cuda(), masks.cuda(), opt=kw, mode="sample")            # prefix masks: fine
    bad = masks.clone()
    bad[2, 3] = 0.0                                                              # a hole inside the valid prefix
    with pytest.raises(ValueError, match="prefix mask"):
        model(fc.cuda(), att.cuda(), bad.cuda(), opt=kw, mode="sample")
    model(fc.cuda(), att.cuda(), masks.cuda(), opt=kw, mode="sample")            # the flag was cleared


def test_host_slot_buffers_follow_the_batch_shape():
    """pipeline.submit_host reuses a slot's pinned outputs: a short last batch followed by a full one (or another
    want_logprobs) must reallocate instead of writing past the old buffers."""
    from boficap_b200.pipeline import BofiPipeline
    cfg = BofiConfig()
    pipe = BofiPipeline(cfg, checkpoint(cfg, "s_real"), 0, "fp32", depth=1)
    _, small, _ = synth.synth_inputs(3, 36, seed=5)
    _, big, _ = synth.synth_inputs(9, 36, seed=6)
    a = pipe.submit_host(small.pin_memory()).wait()
    assert a["seq"].shape[0] == 3
    b = pipe.submit_host(big.pin_memory()).wait()
    assert b["seq"].shape[0] == 9 and b.get("logp") is None
    c = pipe.submit_host(big.pin_memory(), want_logprobs=True).wait()
    assert c["logp"].shape == (9, cfg.seq_length, cfg.tgt_vocab)
    ref = pipe.engines[0]
    ref.encode(big.cuda())
    want = ref.decode("NAIC", 1, 1, False)[0].cpu()
    assert torch.equal(b["seq"], want) and torch.equal(c["seq"], want)
    # the feeder's 2-byte features through the host entry point (fp32 engine: one widening kernel)
    d = pipe.submit_host(big.to(torch.bfloat16).pin_memory()).wait()
    ref.encode(big.to(torch.bfloat16).cuda())
    assert torch.equal(d["seq"], ref.decode("NAIC", 1, 1, False)[0].cpu())
    pipe.close()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_compact_features_equal_padded_features(precision):
    """bofi_encode_compact / BofiPipeline.submit_host_compact: only the valid regions, image after image (sum(att_len) rows instead of
    B * R) give bit for bit the memory and the decode of the padded tensor with the same att_len -- device entry (fp32 and bf16
    features) and the pinned-host pipeline entry."""
    from boficap_b200.engine import BofiEngine
    from boficap_b200.pipeline import BofiPipeline
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_cap")
    eng = BofiEngine(cfg, 0, precision).load_state_dict(sd)
    pipe = BofiPipeline(cfg, sd, 0, precision, depth=2, group=2)
    for (B, R, fdt) in ((70, 60, torch.float32), (200, 100, torch.bfloat16)):
        _, att, masks = synth.synth_inputs(B, R, seed=300 + B, adaptive=True)
        att = att.to(fdt)
        lens = masks.long().sum(1).int()
        compact = torch.cat([att[b, :int(lens[b])] for b in range(B)]).contiguous()
        assert compact.shape[0] == int(lens.sum()) < B * R
        eng.encode(att.cuda(), lens.cuda())
        want = [t.cpu() for t in eng.decode("NAIC", 1, 1, True)]
        eng.encode_compact(compact.cuda(), lens.cuda(), R)
        got = [t.cpu() for t in eng.decode("NAIC", 1, 1, True)]
        for x, y in zip(want, got):
            assert torch.equal(torch.nan_to_num(x.float()), torch.nan_to_num(y.float())), (precision, B, R)
        if fdt == torch.bfloat16:
            t1 = pipe.submit_host_compact(compact.pin_memory(), lens.pin_memory(), R, want_logprobs=True)
            t2 = pipe.submit_host(att.pin_memory(), lens.pin_memory(), want_logprobs=True)       # (a group of one, flushed by wait)
            for t in (t1, t2):
                out = t.wait()
                for x, y in zip(want, (out["seq"], out["logp"], out["pnum"], out["plen"], out["psyn"])):
                    assert torch.equal(torch.nan_to_num(x.float()), torch.nan_to_num(y.float()))
    # two compact batches of the same shape share one library call (group = 2): each ticket equals its stand-alone decode
    if precision == "bf16":
        batches, wants = [], []
        for k in range(2):
            _, att, masks = synth.synth_inputs(96, 80, seed=400 + k, adaptive=True)
            att = att.to(torch.bfloat16)
            lens = masks.long().sum(1).int()
            eng.encode(att.cuda(), lens.cuda())
            wants.append([t.cpu() for t in eng.decode("NAIC", 1, 1, True)])
            batches.append((torch.cat([att[b, :int(lens[b])] for b in range(96)]).contiguous().pin_memory(), lens.pin_memory()))
        tickets = [pipe.submit_host_compact(c, l, 80, want_logprobs=True) for c, l in batches]
        assert all(t.launched for t in tickets)
        for t, w in zip(tickets, wants):
            out = t.wait()
            for x, y in zip(w, (out["seq"], out["logp"], out["pnum"], out["plen"], out["psyn"])):
                assert torch.equal(torch.nan_to_num(x.float()), torch.nan_to_num(y.float()))
    pipe.close()
    eng.close()
