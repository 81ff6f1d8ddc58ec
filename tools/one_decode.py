"""One decode step inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off` captures (profiles/):
two warm-up decodes, then ONE profiled decode of the bench configuration (B=1024, R=36, bf16 engine, bf16 features).
BOFI_GRAPH=0 keeps every launch of the bounding loop a separate kernel node.  Not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--mode", default="NAIC")
    ap.add_argument("--calib", default="s_real")
    ap.add_argument("--adaptive", action="store_true")
    ap.add_argument("--no-logprobs", action="store_true")
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from boficap_b200 import synth
    from boficap_b200.engine import BofiEngine
    from boficap_b200.layout import BofiConfig
    cfg = BofiConfig()
    eng = BofiEngine(cfg, 0, a.precision).load_state_dict(synth.synth_state_dict(cfg, 0, a.calib))
    _, att, masks = synth.synth_inputs(a.batch, a.regions, seed=1, adaptive=a.adaptive)
    att = att.to(torch.bfloat16).cuda() if a.precision == "bf16" else att.cuda()
    att_len = masks.long().sum(1).int().cuda() if masks is not None else None
    for i in range(3):
        if i == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        eng.encode(att, att_len)
        out = eng.decode(a.mode, 1, 1, not a.no_logprobs)
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    info = eng.decode_info()
    print("one decode: %d launches, S=%d, w=%d, nan=%d, mean tokens %.2f" % (info["kernel_launches"], info["bounding_steps"], info["fill_width"],
                                                                          info["nan_batch"], float(out[3].sum(1).float().mean())))


if __name__ == "__main__":
    main()
