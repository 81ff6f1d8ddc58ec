"""Golden vectors of the XE-training forward / backward from the UNMODIFIED reference
model and criterion imported from /root/reference (build container only; outputs committed under tests/golden/).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden_xe.py [case ...]

The reference `_forward` (TransformerModel.py:1713-1775, train_mode UIC, glat_p = -1) is run in eval() mode
(dropout is identity, autograd still records), its six outputs go through LanguageModelCriterion_UIC
(losses.py:315-369) and `.backward()`.  Stored per case: the bounding log-probs in full, the word log-probs at the
labels / their leading columns / row maxima, the seven loss values, and for every parameter the gradient norm,
its leading 64 entries and (tensors of <= 4096 elements) the whole gradient.
"""
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from boficap_b200.layout import BofiConfig  # noqa: E402
from boficap_b200 import synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

CASES = [
    # name, B, R, adaptive, batch seed, glat_p, glat_seed
    ("xe_b2_r12", 2, 12, False, 3, -1.0, 0),
    ("xe_b3_r20_adaptive", 3, 20, True, 5, -1.0, 0),
    # glancing training (EncoderDecoder_UIC.forward :437-464, configs/uic_glat_token.yaml): the reference's torch.rand is
    # replaced by the CUDA path's counter-based uniforms so that the glanced inputs are reproducible
    ("xe_b3_r20_glat", 3, 20, True, 5, 0.5, 11),
    ("xe_b2_r12_glat1", 2, 12, False, 3, 1.0, 4),
    # two bounding layers (configs/uic_sd_N2.yml)
    ("xe_b2_r12_nlen2", 2, 12, False, 3, -1.0, 0, {"N_len": 2}),
]
LOGP_COLS = 32
GRAD_HEAD = 64
GRAD_FULL_MAX = 4096


def reference_criterion():
    # losses.py imports the CIDEr reward helpers (pycocoevalcap etc. are absent here); the XE criterion never calls them
    name = "captioning.utils.rewards"
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.get_scores = m.get_self_cider_scores = lambda *a, **k: None
        m.init_scorer = lambda *a, **k: None
        m.get_self_critical_reward = lambda *a, **k: None
        sys.modules[name] = m
    from captioning.modules import losses
    return losses.LanguageModelCriterion_UIC()


def glat_uniforms(shape, glat_seed):
    """u[n, t] of train_kernels.cuh: glat_input_kernel (== oracle DropSim hash): key = hash(seed, 'GLAT'), index n * T + t."""
    from oracle.bofi_oracle import DropSim
    n = int(np.prod(shape))
    key = int(DropSim._hash(int(glat_seed) & DropSim.M, torch.tensor(0x474C4154, dtype=torch.int64)))
    u = (DropSim._hash(key, torch.arange(n, dtype=torch.int64)).float() + 0.5) * (1.0 / 4294967296.0)
    return u.view(*shape)


def run_case(name, B, R, adaptive, seed, glat_p=-1.0, glat_seed=0, cfg_kw=None):
    cfg = BofiConfig(**(cfg_kw or {}))
    sd = synth.synth_state_dict(cfg, 0, "s_real")
    model, _ = ref_shim.build_reference_model(sd, vocab_size=cfg.vocab_size, N_enc=cfg.N_enc, N_dec=cfg.N_dec, N_len=cfg.N_len)
    # TransformerModel.py:481-482 allocates requires_grad leaves and then writes into them in place (an error on
    # torch >= 2): drop the flag, the tensors pick up the graph from the values assigned into them
    nz = torch.Tensor.new_zeros
    torch.Tensor.new_zeros = lambda self, *a, **k: nz(self, *a, **{kk: v for kk, v in k.items() if kk != "requires_grad"})
    crit = reference_criterion()
    fc, att, masks = synth.synth_inputs(B, R, seed=7, adaptive=adaptive)
    bt = synth.synth_xe_batch(B, seed=seed, vocab_size=cfg.vocab_size)
    model.eval()
    model.zero_grad()
    rand = torch.rand
    if glat_p >= 0:
        torch.rand = lambda shape, **k: glat_uniforms(tuple(shape), glat_seed)
    outs = model(fc, att, bt["labels"], masks, bt["phrase_num"], bt["phrase_length"], bt["phrase_syn"],
                 bt["extend_phrase_syn_seq"], bt["extend_phrase_seq"], bt["extend_phrase_seq_mask"], glat_p)
    torch.rand = rand
    losses = crit(*outs, bt["phrase_num"], bt["phrase_length"], bt["phrase_syn"], bt["labels"], reduction="mean")
    losses[0].backward()
    torch.Tensor.new_zeros = nz
    sa_len, sa_syn, sa_logp, na_len, na_syn, na_logp = [o.detach() for o in outs]
    words = bt["labels"].reshape(-1, bt["labels"].shape[2])[:, 1:-1]
    fix = dict(B=np.array(B), R=np.array(R), adaptive=np.array(adaptive), input_seed=np.array(7), batch_seed=np.array(seed),
               sa_len=sa_len.numpy(), sa_syn=sa_syn.numpy(), na_len=na_len.numpy(), na_syn=na_syn.numpy(),
               losses=np.array([float(v) for v in losses], dtype=np.float64), glat_p=np.array(glat_p), glat_seed=np.array(glat_seed),
               cfg_kw=np.array(json.dumps(cfg_kw or {})))
    for tag, lp in (("sa", sa_logp), ("na", na_logp)):
        fix[tag + "_logp_head"] = lp[:, :, :LOGP_COLS].numpy()
        fix[tag + "_logp_max"] = lp.max(2).values.numpy()
        fix[tag + "_logp_at_label"] = lp.gather(2, words.unsqueeze(2)).squeeze(2).numpy()
    names, norms = [], []
    for pname, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        names.append(pname)
        norms.append(float(g.double().norm()))
        fix["ghead/" + pname] = g.reshape(-1)[:GRAD_HEAD].numpy()
        if g.numel() <= GRAD_FULL_MAX:
            fix["gfull/" + pname] = g.numpy()
    fix["grad_names"] = np.array(json.dumps(names))
    fix["grad_norms"] = np.array(norms, dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **fix)
    print("golden", name, "loss %.5f" % float(losses[0]), "phrase_num", bt["phrase_num"].reshape(-1).tolist(),
          "nonzero grads %d/%d" % (sum(n > 0 for n in norms), len(norms)), "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]                           # optional case names (default: all)
    for case in CASES:
        if not only or case[0] in only:
            run_case(*case)
