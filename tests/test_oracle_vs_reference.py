"""CPU, build container only: live comparison of the oracle with the unmodified reference
(skipped on boxes without /root/reference)."""
import pytest
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig
from oracle import ref_shim
from util import checkpoint, oracle_for

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


@pytest.mark.parametrize("mode,calib,B", [("NAIC", "s_real", 6), ("SAIC", "s_cap", 3)])
def test_live_reference_equals_oracle(mode, calib, B):
    cfg = BofiConfig()
    sd = checkpoint(cfg, calib)
    ref, _ = ref_shim.build_reference_model(sd)
    fc, att, masks = synth.synth_inputs(B, 40, seed=11, adaptive=True)
    with torch.no_grad():
        r = ref(fc, att, masks, opt={"sample_method": "greedy", "train_mode": mode}, mode="sample")
    o = oracle_for(cfg, sd).sample(fc, att, masks, {"train_mode": mode})
    assert torch.equal(r[0], o[0])
    assert torch.allclose(r[1].detach(), o[1], atol=1e-6, equal_nan=True)
    for i in (2, 3, 4):
        assert torch.equal(r[i], o[i])
