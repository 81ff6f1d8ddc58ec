"""Where the tcgen05 GEMM's cycles go, per shape class, inside the real decode step.

Needs the stall-counter build (`python -m boficap_b200.build --prof` -> lib/libbofi_b200_prof.so, kernels compiled
with -DBOFI_GEMM_PROF: every role adds its clock64() waits to a device table, see gemm_tc.cuh).  Run on the GPU box:

    BOFI_LIB_PATH=boficap_b200/lib/libbofi_b200_prof.so python tools/gemm_stalls.py [--batch 1024] [--mode NAIC]

The counters cost a few clock reads per k-block, so the numbers printed here are for attribution, never a bench value.
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

NAMES = ["FFN1 K512 N2048", "FFN2 K2048", "O-proj K512 N512 +resid", "proj K512 N512", "QKV/KV K512", "vocab K512 N9504", "other"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--mode", default="NAIC")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    from boficap_b200 import _lib, synth
    from boficap_b200.engine import BofiEngine
    from boficap_b200.layout import BofiConfig
    raw = _lib.load()
    if not hasattr(raw, "bofi_debug_gemm_prof"):
        raise SystemExit("%s has no stall counters: build with --prof and set BOFI_LIB_PATH" % _lib.LIB_PATH)
    raw.bofi_debug_gemm_prof.argtypes = [C.POINTER(C.c_uint64), C.c_int]
    cfg = BofiConfig()
    eng = BofiEngine(cfg, 0, "bf16").load_state_dict(synth.synth_state_dict(cfg, 0, "s_real"))
    _, att, _ = synth.synth_inputs(a.batch, a.regions, seed=1)
    att = att.cuda()

    def step():
        eng.encode(att, None)
        return eng.decode(a.mode, 1, 1, True)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    buf = (C.c_uint64 * 96)()
    assert raw.bofi_debug_gemm_prof(buf, 1) == 0
    for _ in range(a.steps):
        step()
    assert raw.bofi_debug_gemm_prof(buf, 0) == 0
    rows = []
    print("%-26s %8s %9s | MMA thread: %%operands %%epilogue | TMA %%slot | epilogue warp: %%accum %%staging %%tmem_ld %%math+st.shared %%fence+store %%arrive" % ("class", "CTAs", "Mcyc/CTA"))
    for b, name in enumerate(NAMES):
        v = [int(buf[b * 12 + i]) for i in range(12)]
        if v[7] == 0:
            continue
        mma, w_acc, w_full, w_slot, epi, w_tfull, w_stage, n, w_tld, t_math, t_store, t_arrive = v[:12]
        rows.append(dict(cls=name, ctas=n, mma_cycles=mma, mma_wait_operands=w_full, mma_wait_epilogue=w_acc,
                         tma_wait_slot=w_slot, epi_cycles=epi, epi_wait_accum=w_tfull, epi_wait_staging=w_stage,
                         epi_tmem_ld=w_tld, epi_math=t_math - w_stage, epi_store=t_store, epi_arrive=t_arrive))
        print("%-26s %8d %9.3f |            %8.1f %9.1f | %8.1f |               %6.1f %8.1f %8.1f %8.1f %8.1f %8.1f" % (
            name, n, mma / n / 1e6, 100 * w_full / mma, 100 * w_acc / mma, 100 * w_slot / mma,
            100 * w_tfull / max(epi, 1), 100 * w_stage / max(epi, 1), 100 * w_tld / max(epi, 1),
            100 * (t_math - w_stage) / max(epi, 1), 100 * t_store / max(epi, 1), 100 * t_arrive / max(epi, 1)))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_stalls.json", "w") as f:
        json.dump(dict(batch=a.batch, mode=a.mode, steps=a.steps, rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
