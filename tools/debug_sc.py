"""Debug aid (GPU): gradient norms of the self-critical tape path per precision / mode / size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from boficap_b200 import synth  # noqa: E402
from test_gpu_selfcritical import sc_loss  # noqa: E402
from test_gpu_train import build_model  # noqa: E402

for precision in ("fp32", "bf16"):
    for (B, n) in ((4, 2), (64, 5), (256, 5)):
        if precision == "fp32" and B > 64:
            continue
        model, cfg = build_model(precision)
        model.train()
        model.train_bind()
        fc, att, _ = synth.synth_inputs(B, 36, seed=3)
        fc, att = fc.cuda(), att.cuda()
        for mode in ("NAIC", "SAIC"):
            for rep in range(2):
                torch.manual_seed(11)
                model.bofi_dropout_seed, model._train_steps = 21, 0
                model.zero_grad()
                seq, logp, pnum, plen, psyn, _ = model(fc, att, None, opt={"sample_method": "sample", "sample_n": n, "train_mode": mode}, mode="sample")
                reward = torch.randn(seq.shape[0], device="cuda", generator=torch.Generator("cuda").manual_seed(1))
                loss = sc_loss(logp, seq, reward)
                loss.backward()
                torch.cuda.synchronize()
                g = model.flat_grads()
                info = model._engine.decode_info()
                print("%s B=%d n=%d %s rep%d: loss %.5f logp finite %.3f nan %d tokens/row %.2f w=%d |g| %.6g nonzero %.4f nan_g %d requires_grad %s"
                      % (precision, B, n, mode, rep, float(loss), float(torch.isfinite(logp).float().mean()), int(torch.isnan(logp).sum()),
                         float((seq > 0).float().sum(1).mean()), info["fill_width"], float(g.double().norm()), float((g != 0).float().mean()),
                         int(torch.isnan(g).sum()), logp.requires_grad), flush=True)
        model._engine.close()
