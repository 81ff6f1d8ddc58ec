"""torchrun --nproc-per-node N tools/check_grad_reduce.py : the bucketed, overlapped gradient all-reduce (OverlappedGradReduce: decoder-side
buckets under the encoder's backward pass, encoder layers one by one under the layers below) must leave in the flat gradient buffer exactly what
ONE all-reduce (AVG) of the whole buffer after the backward pass leaves."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from boficap_b200 import synth
    from boficap_b200.captioning import models
    from boficap_b200.layout import BofiConfig
    from boficap_b200.parallel import OverlappedGradReduce
    cfg = BofiConfig()
    infos = synth.make_infos(cfg)
    opt = infos["opt"]
    opt.vocab = infos["vocab"]
    opt.bofi_precision = "bf16"
    model = models.setup(opt)
    model.load_state_dict(synth.synth_state_dict(cfg, 0, "s_real"))
    model = model.cuda().train()
    model.train_bind()
    B = 64
    fc, att, masks = synth.synth_inputs(B, 36, seed=1 + rank)
    bt = synth.synth_xe_batch(B, seq_per_img=5, seed=11 + rank, vocab_size=cfg.vocab_size)
    dev = lambda t: t.cuda() if t is not None else None
    args = (dev(fc), dev(att), dev(bt["labels"]), dev(masks), dev(bt["phrase_num"]), dev(bt["phrase_length"]), dev(bt["phrase_syn"]),
            dev(bt["extend_phrase_syn_seq"]), dev(bt["extend_phrase_seq"]), dev(bt["extend_phrase_seq_mask"]))
    flat = model.flat_grads()
    worst = 0.0
    for layer_buckets in (True, False):
        red = OverlappedGradReduce(model, buckets=4, layer_buckets=layer_buckets)
        for it in range(3):
            model.bofi_dropout_seed, model._train_steps = 100 + rank, it
            flat.zero_()
            model.xe_step(*args)
            torch.cuda.synchronize()
            want = flat.clone()
            dist.all_reduce(want, op=dist.ReduceOp.AVG)
            model.bofi_dropout_seed, model._train_steps = 100 + rank, it
            flat.zero_()
            model.xe_step(*args)
            red.reduce()
            torch.cuda.synchronize()
            err = float((flat - want).abs().max()) / max(float(want.abs().max()), 1e-12)
            worst = max(worst, err)
        red.close()
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("overlapped gradient all-reduce vs one all-reduce: worst relative difference %.3e over %d ranks" % (float(t), dist.get_world_size()))
        assert float(t) < 1e-5, float(t)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
