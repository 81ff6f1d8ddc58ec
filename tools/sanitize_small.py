"""Every product path once at a tiny size, for `compute-sanitizer --tool memcheck` (out-of-bounds / misaligned accesses in the
kernels): NAIC and SAIC decode in both precisions with padded regions (varlen encoder, fused vocabulary projection), the
host entry point with bf16 features, sample_stats, the XE step and sampling with a tape.  Not a benchmark."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from boficap_b200 import synth  # noqa: E402
from boficap_b200.captioning import models  # noqa: E402
from boficap_b200.layout import BofiConfig  # noqa: E402


def main():
    cfg = BofiConfig()
    infos = synth.make_infos(cfg)
    for precision in ("bf16", "fp32"):
        opt = infos["opt"]
        opt.vocab = infos["vocab"]
        opt.bofi_precision = precision
        model = models.setup(opt)
        model.load_state_dict(synth.synth_state_dict(cfg, 0, "s_cap"))
        model = model.cuda().eval()
        fc, att, masks = synth.synth_inputs(14, 23, seed=5, adaptive=True)
        fc, att, masks = fc.cuda(), att.cuda(), masks.cuda()
        for mode in ("NAIC", "SAIC"):
            for kw in ({"sample_method": "greedy"}, {"sample_method": "sample", "sample_n": 2, "temperature": 0.9}):
                out = model(fc, att, masks, opt=dict(kw, train_mode=mode), mode="sample")
                assert out[0].shape[1] == cfg.seq_length
            model(fc, att.to(torch.bfloat16), None, opt={"train_mode": mode, "output_logsoftmax": 0}, mode="sample")
        model.sample_stats(fc, att, masks, opt={"train_mode": "NAIC"})
        eng = model._engine
        host = att.cpu().to(torch.bfloat16).pin_memory()
        eng.sample_host(host, masks.sum(1).int().cpu(), "NAIC", 1, 1, want_logprobs=True)
        # decode at a size that takes the pair GEMM and the fused vocabulary path (rows * L >= 256, M >= 2048)
        fc2, att2, m2 = synth.synth_inputs(70, 36, seed=6, adaptive=True)
        model(fc2.cuda(), att2.cuda(), m2.cuda(), opt={"train_mode": "NAIC"}, mode="sample")
        # training paths
        model.train()
        bt = synth.synth_xe_batch(4, seed=3, vocab_size=cfg.vocab_size)
        fc3, att3, m3 = synth.synth_inputs(4, 20, seed=7, adaptive=True)
        args = (fc3.cuda(), att3.cuda(), bt["labels"].cuda(), m3.cuda(), bt["phrase_num"].cuda(), bt["phrase_length"].cuda(), bt["phrase_syn"].cuda(),
                bt["extend_phrase_syn_seq"].cuda(), bt["extend_phrase_seq"].cuda(), bt["extend_phrase_seq_mask"].cuda())
        model.xe_step(*args)
        outs = model(*args)
        sum(o.sum() for o in outs).backward()
        for mode in ("NAIC", "SAIC"):
            seq, logp, *_ = model(fc3.cuda(), att3.cuda(), m3.cuda(), opt={"sample_method": "sample", "sample_n": 2, "train_mode": mode}, mode="sample")
            logp.gather(2, seq.unsqueeze(2)).sum().backward()
        torch.cuda.synchronize()
        print("sanitize_small: %s paths ran" % precision, flush=True)
        model._engine.close()


if __name__ == "__main__":
    main()
