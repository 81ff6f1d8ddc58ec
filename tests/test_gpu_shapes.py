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
