"""CPU: BofiOracle.forward_sc (the checker of the self-critical tape pass) against the oracle's own decode, which is pinned
to the unmodified reference by the golden fixtures: with dropout off, scoring the boxes a NAIC decode produced gives that
decode's log-probs; scoring the words a SAIC decode produced gives its committed log-probs."""
import torch

from boficap_b200 import synth
from boficap_b200.layout import BofiConfig
from util import checkpoint, oracle_for


def test_forward_sc_reproduces_naic_and_saic_logprobs():
    cfg = BofiConfig()
    sd = checkpoint(cfg, "s_cap")
    fc, att, masks = synth.synth_inputs(3, 14, seed=9, adaptive=True)
    L = cfg.seq_length
    o = oracle_for(cfg, sd, record=True)
    seq, logp, pnum, plen, psyn, _ = o.sample(fc, att, masks, {"train_mode": "NAIC"})
    ext, w = o.trace["ext"], o.trace["fill_width"]
    words = torch.full((3, L), cfg.bos_idx)
    vis = torch.full((3, L), max(w, 0))
    with torch.no_grad():
        got = o.forward_sc(att, masks, words, ext[:, 1:-1], vis, 1)
    assert (got - logp).abs().max().item() < 1e-5
    # SAIC: rebuild the decoder inputs from the returned boxes and words (position-wise copy, :1928-1948)
    o2 = oracle_for(cfg, sd)
    seq, logp, pnum, plen, psyn, _ = o2.sample(fc, att, masks, {"train_mode": "SAIC"})
    words, syns, vis = torch.zeros(3, L, dtype=torch.long), torch.zeros(3, L, dtype=torch.long), torch.ones(3, L, dtype=torch.long)
    for b in range(3):
        full = torch.cat([torch.tensor([cfg.bos_idx]), seq[b]])                  # bos + generated words
        lens = [1] + [int(v) for v in plen[b] if int(v) > 0]
        src, dst = 0, 0
        for j in range(1, len(lens)):
            cur, prev = lens[j], lens[j - 1]
            if cur <= prev:
                words[b, dst:dst + cur] = full[src + prev - cur:src + prev]
            else:
                few, ct = prev - cur % prev, cur // prev
                reps = torch.tensor([ct if q < few else ct + 1 for q in range(prev)])
                words[b, dst:dst + cur] = torch.repeat_interleave(full[src:src + prev], reps)
            syns[b, dst:dst + cur] = int(psyn[b][j - 1])
            vis[b, dst:] = dst + cur
            src += prev
            dst += cur
    with torch.no_grad():
        got = o2.forward_sc(att, masks, words, syns, vis, 1)
    committed = torch.arange(L)[None, :] < plen.sum(1)[:, None]
    assert committed.any()
    assert (got[committed] - logp[committed]).abs().max().item() < 1e-5
    assert (logp[~committed] == 0).all()
