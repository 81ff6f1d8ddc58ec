#!/usr/bin/env python
"""Benchmark of the BoFi greedy decode hot path (BASELINE.json: captions/sec/GPU, uic_sd, 36x2048 regions).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--regions R]
                  [--precision bf16|fp32] [--mode NAIC] [--no-logprobs]

One "step" = one `_sample` of a batch of B synthetic images on each GPU (encode + bounding + filling +
vocab log-softmax/argmax).  N > 1 runs one process per GPU under torchrun (batch sharded by image, no
collective on the data path, weak scaling: B images per GPU).  Rank 0 prints ONE JSON line.

The `reference` arm times the reference algorithm's CPU restatement (oracle/bofi_oracle.py, "port") on the
host cores of this box on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "captions/sec/GPU greedy BoFi decode (uic_sd, 36x2048 regions)"
# roofline.traffic is NOT measured by this run (ncu cannot wrap a bench value): it is the DRAM read + write bytes of one
# launch of the dominant kernel's largest shape from the committed ncu --set full capture named here
TRAFFIC_CONSTANT = {"bytes": 135.1e6,
                    "source": "constant from profiles/r01d_gemm_ncu_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum of one "
                              "gemm_tc2_kernel launch M=36864 N=2048 K=512 (39.9 MB + 95.2 MB; algorithmic 189 MB, most of the output "
                              "stays in the 126 MB L2)"}


def useful_flops_per_caption(R, S, V=9491, n_enc=6, n_dec=6, d=512, dff=2048, L=20, Lb=22):
    """SURVEY.md Appendix C, 'useful' formulation (N_len = 1)."""
    mm = lambda m, n, k: 2.0 * m * n * k
    att = lambda q, k: 2.0 * mm(q, k, d)
    enc = 4 * mm(R, d, d) + att(R, R) + 2 * mm(R, dff, d)
    dec = 6 * mm(L, d, d) + 2 * mm(R, d, d) + att(L, L) + att(L, R) + 2 * mm(L, dff, d)
    heads = 2 * mm(1, 100, d) + mm(1, 20, 100) + mm(1, 10, 100)
    bound = 2 * mm(Lb, d, d) + 2 * mm(R, d, d) + S * (4 * mm(1, d, d) + att(1, Lb) + att(1, R) + 2 * mm(1, dff, d) + heads)
    return mm(R, d, 2048) + n_enc * enc + bound + n_dec * dec + mm(L, V, d)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions (NVML, every 5 ms; B200_PROFILING.md clocks line)."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.sm, self.bits, self.mx, self.power = index, [], 0, 0, 0.0
        self.h = None
        self._run = False
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _sample(self):
        try:
            self.sm.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            self.power = max(self.power, self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
        except Exception:
            pass

    def _loop(self):
        while self._run:
            self._sample()
            time.sleep(0.005)

    def start(self):
        if self.h is None:
            return
        self._run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def pause(self):
        if self.h is None or not self._run:
            return
        self._run = False
        self.t.join()

    def stop(self):
        self.pause()
        if self.h is None or not self.sm:
            return None
        sm = sorted(self.sm)
        return {"sm_mhz": float(sm[len(sm) // 2]), "sm_min_mhz": float(sm[0]), "sm_max_mhz": float(self.mx),
                "reasons": [n for b, n in self.REASONS if self.bits & b], "samples": len(sm), "power_w_max": round(self.power, 1)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 1400.0, 6650.0, "fallback"


def xe_eager_bar(cfg, sd, att, masks, bt, iters=2):
    """The GPU bar of the XE step: the reference algorithm (oracle forward_xe + LanguageModelCriterion_UIC + autograd backward) as eager
    PyTorch on the same B200, same batch, dropout off, fp32 and bf16 autocast.  No optimiser step in the timed region."""
    from oracle.bofi_oracle import BofiOracle, OracleConfig
    out = {}
    dev = lambda t: t.cuda() if t is not None else None
    for name, ctx in (("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16)), ("fp32", torch.autocast("cuda", enabled=False))):
        o = BofiOracle(sd, OracleConfig(**cfg.to_dict()), device="cuda")
        for k, v in o.sd.items():
            if k != "model.pos_embed.pe":
                v.requires_grad_(True)

        def step():
            with ctx, torch.device("cuda"):      # index / mask tensors are created on the oracle's device (as BofiOracle.sample does)
                outs = o.forward_xe(dev(att), dev(masks), dev(bt["labels"]), dev(bt["phrase_num"]), dev(bt["phrase_length"]),
                                    dev(bt["extend_phrase_syn_seq"]), dev(bt["extend_phrase_seq"]), dev(bt["extend_phrase_seq_mask"]))
                loss, _ = o.loss_xe([t.float() for t in outs], bt["phrase_num"], bt["phrase_length"], bt["phrase_syn"], bt["labels"])
            loss.backward()
            for v in o.sd.values():
                v.grad = None
            return float(loss)
        try:
            step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                loss = step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / iters
            out[name] = {"value": att.shape[0] / dt, "unit": "images/s", "ms_per_step": dt * 1e3, "loss": loss}
        except Exception as ex:                 # e.g. out of memory: reported, not fatal
            out[name] = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
        del o
        torch.cuda.empty_cache()
    out["what"] = "oracle/bofi_oracle.py forward_xe + loss_xe + backward on cuda:0 (eager PyTorch, cuBLAS / ATen), dropout off, %d calls each" % iters
    return out


def bench_xe(a, rank, local_rank, world):
    """Config 5 of BASELINE.json: uic_sd XE training step, `--batch` images x 5 captions per GPU (default 256), forward +
    LanguageModelCriterion_UIC + backward in the library, NCCL all-reduce of the flat gradient buffer, Adam on the flat
    parameter buffer (torch's fused optimiser: plumbing), weight refresh.  One JSON line (not the headline metric)."""
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from boficap_b200 import synth
    from boficap_b200.captioning import models
    from boficap_b200.layout import BofiConfig
    from boficap_b200.parallel import OverlappedGradReduce
    cfg = BofiConfig(N_len=a.n_len)           # 1: uic_sd.yml; 2: uic_sd_N2.yml (two bounding layers)
    B = 256 if a.batch == 1024 else a.batch
    R, spi = a.regions, 5
    infos = synth.make_infos(cfg)
    opt = infos["opt"]
    opt.vocab = infos["vocab"]
    opt.bofi_precision = a.precision
    model = models.setup(opt)
    model.load_state_dict(synth.synth_state_dict(cfg, 0, a.calib))
    model = model.cuda()
    model.train(not a.no_dropout)             # train(): dropout 0.1 / 0.5 on, as in the reference's training loop
    model.train_bind()
    flat_w, flat_g = model.flat_params(), model.flat_grads()
    flat_p = torch.nn.Parameter(flat_w)          # same storage: Adam over the whole buffer == Adam over every tensor
    flat_p.grad = flat_g
    optim = torch.optim.Adam([flat_p], lr=1e-5, betas=(0.9, 0.98), eps=1e-9, fused=True)
    fc, att, masks = synth.synth_inputs(B, R, seed=1 + rank, adaptive=a.adaptive)
    bt = synth.synth_xe_batch(B, seq_per_img=spi, seed=11 + rank, vocab_size=cfg.vocab_size)
    dev = lambda t: t.cuda() if t is not None else None
    args = (dev(fc), dev(att), dev(bt["labels"]), dev(masks), dev(bt["phrase_num"]), dev(bt["phrase_length"]), dev(bt["phrase_syn"]),
            dev(bt["extend_phrase_syn_seq"]), dev(bt["extend_phrase_seq"]), dev(bt["extend_phrase_seq_mask"]))
    eng = model._engine
    # bucketed all-reduce of the flat gradient buffer, the decoder-side buckets underneath the encoder's backward pass
    reducer = OverlappedGradReduce(model, buckets=a.buckets, layer_buckets=a.layer_buckets) if dist is not None else None

    def step():
        flat_g.zero_()
        losses = model.xe_step(*args)
        if reducer is not None:
            reducer.reduce()
        optim.step()
        eng.refresh_weights()
        return losses

    for _ in range(max(3, a.warmup)):
        losses = step()
    torch.cuda.synchronize()
    first = float(losses[0])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        losses = step()
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    launches = eng.train_launches()
    # per-class profile of one more step
    eng.set_profiling(True)
    flat_g.zero_()
    model.xe_step(*args)
    prof = eng.get_profile()
    eng.set_profiling(False)
    eager = None
    if rank == 0 and world == 1 and not a.no_extras:
        try:
            eager = xe_eager_bar(cfg, synth.synth_state_dict(cfg, 0, a.calib), att, masks, bt)
        except Exception as ex:
            eager = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    if rank == 0:
        g = prof["gemm_tcgen05"] if a.precision == "bf16" else prof["gemm_ffma"]
        burst, sustained, hbm, how = measured_peaks()
        tf = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
        total_ms = sum(v["ms"] for v in prof.values())
        line = {"metric": "XE training images/sec (uic_sd, %d captions/image, %dx2048 regions)" % (spi, R),
                "value": world * B * a.steps / (ms / 1e3), "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
                "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.precision,
                "data": "synthetic",
                "config": {"workload": "uic_sd%s XE training step (forward + criterion + backward + all-reduce + Adam), %d images x %d captions per GPU, "
                                       "%d regions, %s, dropout %s" % ("" if a.n_len == 1 else "_N%d" % a.n_len, B, spi, R, a.precision,
                                                                       "off" if a.no_dropout else "on (p=0.1, att_embed 0.5)"),
                           "parallelism": "data-parallel replicas x%d, NCCL all-reduce of the %.0f MB flat gradient buffer in %d + %s buckets, the decoder-side "
                                          "buckets overlapped with the encoder's backward pass%s"
                                          % (world, flat_g.numel() * 4 / 1e6, a.buckets, "%d + 1" % cfg.N_enc if a.layer_buckets else "1",
                                             ", the encoder layers' buckets with the backward pass of the layers below" if a.layer_buckets else "")},
                "loss_first": first, "loss_last": float(losses[0]), "gpu_launches": launches * a.steps, "clocks": clocks,
                "gpu_eager": eager,
                "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05), all %d launches of a step" % g["launches"],
                             "achieved": tf, "peak": sustained, "unit": "TFLOP/s", "frac": tf / sustained, "traffic": None,
                             "share_of_profiled_kernel_time": g["ms"] / total_ms if total_ms else None,
                             "classes": {k: {"launches": v["launches"], "ms": round(v["ms"], 3)} for k, v in prof.items()}}}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def numa_bind(local_rank):
    """Pin this process (and with it the first-touch placement of its pinned host buffers) to the CPUs of the NUMA node
    the GPU hangs off.  Round 1: every rank allocated on node 0 and the 8-GPU end-to-end rate stopped at 181 GB/s of
    aggregate H2D.  Returns a description for the JSON line (None when the topology is not exposed)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(phys)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return {"gpu_numa_node": node, "bound": False}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"gpu_numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"gpu_numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as ex:           # no NVML / sysfs entry: leave the placement to the OS
        return {"bound": False, "why": type(ex).__name__}


def workload_config(a, world):
    """The workload description shared by both arms (`--impl ours` and `--impl reference` time the same thing)."""
    return {"workload": "uic_sd greedy BoFi decode (%s), batch %d/GPU, %d regions x 2048" % (a.mode, a.batch, a.regions),
            "checkpoint": "synthetic seed 0, calibration " + a.calib, "global_batch": a.batch * world,
            "regions": ("adaptive 10..%d, prefix masks" % a.regions) if a.adaptive else a.regions,
            "parallelism": "image-sharded replicas x%d" % world,
            "l2": "per-step input (%.0f MB/GPU as fp32) and activations exceed the 126 MB L2" % (a.batch * a.regions * 2048 * 4 / 1e6),
            "logprobs": "skipped" if a.no_logprobs else ("materialised [B,20,V] fp32" + (
                "" if a.mode != "NAIC" else ", rows on a 16-byte pitch (the [:, :, :V] view of a [B,20,9492] buffer, bofi_decode_ex)"))}


def reference_arm(a, cfg, config):
    """`--impl reference`: the reference's CPU implementation of the path (its PyTorch-fp32 restatement, oracle/bofi_oracle.py --
    the reference itself needs /root/reference at run time and cannot travel to the GPU box) on ALL host cores, on the arm's
    own workload: every step is one `_sample` of the full batch unless that would run for more than a few minutes, in
    which case the step is a bounded sample of the batch (stated in `sample`)."""
    from boficap_b200 import synth
    from oracle.bofi_oracle import BofiOracle, OracleConfig
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.synth_state_dict(cfg, 0, a.calib)
    o = BofiOracle(sd, OracleConfig(**cfg.to_dict()))
    kw = {"sample_method": "greedy", "beam_size": 1, "sample_n": 1, "train_mode": a.mode}
    images = a.batch if a.cpu_images <= 0 else min(a.batch, a.cpu_images)
    fc, att, masks = synth.synth_inputs(images, a.regions, seed=1, adaptive=a.adaptive)
    steps, warmup = max(1, a.steps), max(1, a.warmup)
    t0 = time.perf_counter()
    o.sample(fc, att, masks, kw)                                  # first warm-up step, also the size probe
    probe = time.perf_counter() - t0
    if probe * (steps + warmup) > a.cpu_budget and images > 64:     # bounded sample of the batch
        images = max(64, int(images * a.cpu_budget / (probe * (steps + warmup))) // 8 * 8)
        fc, att = fc[:images].contiguous(), att[:images].contiguous()
        masks = masks[:images].contiguous() if masks is not None else None
        o.sample(fc, att, masks, kw)
    for _ in range(warmup - 1):
        o.sample(fc, att, masks, kw)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.sample(fc, att, masks, kw)
    dt = time.perf_counter() - t0
    cps = images * steps / dt
    sample = ("%d of the %d images of a step, %d steps + %d warm-up; PyTorch fp32 restatement of the reference formulation "
              "(all 22 bounding rows per step, K/V re-projected per call), %d host threads, S=%d"
              % (images, a.batch, steps, warmup, torch.get_num_threads(), o.last_steps))
    return {"impl": "reference", "metric": METRIC, "value": cps, "unit": "captions/s", "n_gpus": a.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config, "bounding_steps": o.last_steps,
            "cpu_baseline": {"value": cps, "unit": "captions/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": cps, "unit": "captions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def gpu_eager_bar(cfg, sd, att, masks, mode, iters=3):
    """The GPU bar SURVEY.md section 8(d) asks for: the reference algorithm as eager PyTorch on the SAME B200 (the oracle port on
    `cuda`: library kernels, vectorised box bookkeeping -- faster than the reference's own per-row python loops), fp32
    and bf16 autocast, same batch, device-resident inputs."""
    from oracle.bofi_oracle import BofiOracle, OracleConfig
    o = BofiOracle(sd, OracleConfig(**cfg.to_dict()), device="cuda")
    kw = {"sample_method": "greedy", "beam_size": 1, "sample_n": 1, "train_mode": mode}
    out = {}
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        with ctx:
            o.sample(None, att, masks, kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                o.sample(None, att, masks, kw)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
        out[name] = {"value": att.shape[0] / dt, "unit": "captions/s", "ms_per_step": dt * 1e3, "bounding_steps": o.last_steps}
    out["what"] = "oracle/bofi_oracle.py on cuda:0 (eager PyTorch, cuBLAS / ATen kernels), batch %d, %d calls each" % (att.shape[0], iters)
    del o
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1024, help="images per GPU per step")
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--mode", default="NAIC", choices=["NAIC", "SAIC"])
    ap.add_argument("--adaptive", action="store_true", help="10..R valid regions per image with prefix masks (config 3)")
    ap.add_argument("--calib", default=None, help="synthetic checkpoint calibration (default: s_real for NAIC, s_cap for SAIC)")
    ap.add_argument("--no-logprobs", action="store_true", help="skip materialising the [B,20,V] log-prob tensor")
    ap.add_argument("--cpu-images", type=int, default=0, help="reference arm: images per step (0 = the full batch, bounded by --cpu-budget)")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="reference arm: seconds of CPU work the whole run may take")
    ap.add_argument("--workload", default="decode", choices=["decode", "xe"],
                    help="decode = BASELINE.json's headline metric; xe = XE training step (config 5: 256 images x 5 captions per GPU)")
    ap.add_argument("--no-dropout", action="store_true", help="xe workload: eval() arithmetic (dropout off)")
    ap.add_argument("--depth", type=int, default=3, help="engine handles x streams in flight (boficap_b200/pipeline.py)")
    ap.add_argument("--compact", action="store_true", help="with --adaptive: the e2e leg sends compact features (valid regions only, "
                                                            "bofi_stage_compact) instead of the padded [B, R, F] batch")
    ap.add_argument("--layer-buckets", action="store_true", help="xe workload: the encoder part of the gradients as one all-reduce per encoder "
                                                                  "layer underneath the backward pass instead of ONE after it (measured: no gain)")
    ap.add_argument("--n-len", type=int, default=1, help="xe workload: bounding layers (configs/uic_sd_N2.yml: 2)")
    ap.add_argument("--group", type=int, default=2, help="consecutive batches decoded by ONE library call on a slot, every batch with its own "
                                                          "fill window (bofi_set_shard): bit-identical results, one bounding loop per group")
    ap.add_argument("--host-dtype", default="bf16", choices=["bf16", "fp16", "fp32"], help="element type of the pinned host features of the e2e leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary legs (fp32-host e2e, drop-in class, GPU-eager bar, CPU baseline)")
    ap.add_argument("--buckets", type=int, default=4, help="xe workload: gradient all-reduce buckets overlapped with the backward pass")
    a = ap.parse_args()
    if a.calib is None:
        # SAIC on the NAIC-only calibration aborts at step 1 ("phrase nan!"): its bench needs the joint calibration
        a.calib = "s_cap" if (a.mode == "SAIC" and a.workload == "decode") else "s_real"

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if a.workload == "xe":
        return bench_xe(a, rank, local_rank, world)

    from boficap_b200 import synth
    from boficap_b200.layout import BofiConfig
    cfg = BofiConfig()
    config = workload_config(a, world)

    if a.impl == "reference":
        if rank != 0:
            return 0
        print(json.dumps(reference_arm(a, cfg, config)))
        return 0

    # ------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    all_cpus = os.sched_getaffinity(0)
    numa = numa_bind(local_rank)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from boficap_b200.parallel import gather_captions_packed
    from boficap_b200.pipeline import BofiPipeline

    sd = synth.synth_state_dict(cfg, 0, a.calib)
    pipe = BofiPipeline(cfg, sd, local_rank, a.precision, depth=max(1, a.depth), group=max(1, a.group) if a.mode == "NAIC" else 1)
    eng = pipe.engines[0]
    B, R = a.batch, a.regions
    fc, att_host, masks = synth.synth_inputs(B, R, seed=1 + rank, adaptive=a.adaptive)
    att_host = att_host.pin_memory()
    # features in the feeder's format (boficap_b200/data: 2-byte elements, pinned): half the H2D bytes of the loader's fp32,
    # and a bf16 engine's att_embed GEMM reads them in place (no conversion pass on the device)
    host_dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.host_dtype]
    att_host_lp = att_host.to(host_dt).pin_memory() if host_dt != torch.float32 else att_host
    att = att_host_lp.cuda(non_blocking=True)
    len_host = masks.long().sum(1).int().pin_memory() if masks is not None else None
    att_len = len_host.cuda() if len_host is not None else None
    notes = {"feature_dtype": a.host_dtype}
    want_lp = not a.no_logprobs
    main = torch.cuda.current_stream()

    def step():
        eng.encode(att, att_len)
        return eng.decode(a.mode, 1, 1, want_lp)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # The reference masks every row's fill attention with last[B-1] (TransformerModel.py:1871-1873): an image
    # that predicts no phrase in the LAST slot would turn the whole batch into NaN.  Boxes do not depend on
    # the batch, so put an image that does produce phrases last (same work, no NaN batch).
    out = step()
    torch.cuda.synchronize()
    tokens = out[3].sum(1)
    seed_shift = 0
    while int(tokens.sum()) == 0 and seed_shift < 32 and a.mode == "NAIC":      # tiny batches: draw other images until one yields a phrase
        seed_shift += 1
        att_host.copy_(synth.synth_inputs(B, R, seed=1 + rank + 1000 * seed_shift, adaptive=False)[1])
        att_host_lp.copy_(att_host)
        att.copy_(att_host_lp)
        out = step()
        torch.cuda.synchronize()
        tokens = out[3].sum(1)
    if seed_shift:
        notes["input_seed_shift"] = seed_shift
    if a.mode == "NAIC" and int(tokens[-1]) == 0:
        j = int((tokens > 0).nonzero()[-1])
        att_host[[j, B - 1]] = att_host[[B - 1, j]]
        att_host_lp.copy_(att_host)
        att.copy_(att_host_lp)
        if len_host is not None:
            len_host[[j, B - 1]] = len_host[[B - 1, j]]
            att_len.copy_(len_host)
        notes["last_image"] = "swapped with image %d so that the fill window is non-empty" % j
    out = step()
    torch.cuda.synchronize()
    ref_seq = out[0].clone()
    info0 = eng.decode_info()
    if info0["nan_batch"]:
        # a NaN batch decodes nothing (the reference returns NaN log-probs / aborts SAIC at step 1): no captions/s to report
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": None, "unit": "captions/s", "n_gpus": world, "config": config,
                              "invalid": "nan_batch: checkpoint calibration %r yields no caption in %s mode on this batch "
                                         "(use --calib s_cap for SAIC)" % (a.calib, a.mode), "decode_info": info0}))
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 1

    gathered = [None]

    def finish_ticket(t, st):
        """Final caption gather of the N-GPU path (the one collective north_star allows), inside the timed region."""
        if dist is None:
            return
        with torch.cuda.stream(st):
            o = t.out
            seq, pnum, plen, psyn = (o[0], o[2], o[3], o[4]) if isinstance(o, tuple) else (o["seq"], o["pnum"], o["plen"], o["psyn"])
            if not seq.is_cuda:
                return
            gathered[0] = gather_captions_packed(seq, pnum, plen, psyn)

    def run_device(n):
        """n whole decodes, round robin over the in-flight slots; returns the tickets."""
        pipe.fork_from(main)
        ts = []
        for _ in range(n):
            ts.append(pipe.submit_device(att, att_len, a.mode, 1, 1, want_lp, reuse_outputs=True))
            for u in pipe.just_launched:                     # (a grouped submission is enqueued when its group is complete)
                finish_ticket(u, pipe.streams[u.slot])
        pipe.flush()
        for u in pipe.just_launched:
            finish_ticket(u, pipe.streams[u.slot])
        pipe.join_into(main)
        return ts

    for t in run_device(max(a.warmup, 2 * pipe.depth * pipe.group)):     # every slot: eager run, graph capture, replay
        t.wait()
    torch.cuda.synchronize()
    info = eng.decode_info()
    launches_per_step = info["kernel_launches"]               # of one library call = one group of `pipe.group` batches (+ the staging copies)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(main)
    tickets = run_device(a.steps)
    ev1.record(main)
    barrier()
    ms = ev0.elapsed_time(ev1)
    sampler.pause()
    out = tickets[-1].out
    for t in tickets[-pipe.depth * pipe.group:]:             # every slot (and every batch of a group) reproduces the stand-alone decode
        assert torch.equal(t.out[0], ref_seq), "pipelined decode differs from the stand-alone decode"
    del tickets

    # ---- e2e: the host-buffer entry point (H2D of the features and D2H of captions + boxes inside the timed region),
    # pinned host input, `depth` batches in flight through bofi_sample_host_async_ex
    def run_host(n, feats):
        pipe.fork_from(main)
        if feats.dim() == 2:                                     # compact features: only the valid regions cross PCIe
            ts = [pipe.submit_host_compact(feats, len_host, R, a.mode, 1, 1) for _ in range(n)]
        else:
            ts = [pipe.submit_host(feats, len_host, a.mode, 1, 1) for _ in range(n)]
        pipe.join_into(main)
        return ts

    def time_host(feats):
        for t in run_host(2 * pipe.depth * pipe.group, feats):
            host_out = t.wait()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(main)
        tickets = run_host(a.steps, feats)
        e1.record(main)
        for t in tickets:
            host_out = t.wait()                                  # results are in host memory here
        if dist is not None:                                     # caption gather of the N-GPU path (host results -> every rank)
            packed = torch.cat([host_out["seq"].int(), host_out["pnum"].int()[:, None], host_out["plen"], host_out["psyn"].int()], 1).cuda()
            allc = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device="cuda")
            dist.all_gather_into_tensor(allc, packed)
            torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        barrier()
        return max(e0.elapsed_time(e1), wall), host_out          # device span vs host wall clock until the last result landed

    if rank == 0:
        sampler.start()
    e2e_feats = att_host_lp
    if a.compact and len_host is not None:
        # the feeder's compact format (boficap_b200/data: PinnedFeeder(compact=True)): the valid regions only, image after image
        e2e_feats = torch.cat([att_host_lp[b, :int(len_host[b])] for b in range(B)]).contiguous().pin_memory()
    ms_e2e, host_out = time_host(e2e_feats)
    clocks = sampler.stop() if rank == 0 else None
    assert torch.equal(host_out["seq"], ref_seq.cpu()), "host-path captions differ from the device path"
    h2d = e2e_feats.numel() * e2e_feats.element_size() + (len_host.numel() * 4 if len_host is not None else 0)
    d2h = sum(host_out[k].numel() * host_out[k].element_size() for k in ("seq", "pnum", "plen", "psyn"))
    ms_e2e32 = None
    if not a.no_extras and host_dt != torch.float32:
        ms_e2e32, host_out32 = time_host(att_host)
        notes["fp32_host_token_agreement"] = float((host_out32["seq"] == ref_seq.cpu()).float().mean())   # 1.0 for bf16 features

    # ---- the reference-facing class: model(fc, att, masks, opt, mode='sample') on pinned host tensors, .cuda() and the
    # read-back of the captions inside the timed region (eval_utils.py:431-445 does exactly this per batch), one batch in flight
    ms_dropin = ms_dropin_serial = None
    if not a.no_extras and world == 1:
        from boficap_b200.captioning import models
        infos = synth.make_infos(cfg)
        mopt = infos["opt"]
        mopt.vocab = infos["vocab"]
        mopt.bofi_precision = a.precision
        model = models.setup(mopt)
        model.load_state_dict(sd)
        model = model.cuda().eval()
        fc_host = fc.pin_memory()
        kw = {"sample_method": "greedy", "beam_size": 1, "sample_n": 1, "train_mode": a.mode, "bofi_logprobs": want_lp}
        mk_host = masks.pin_memory() if masks is not None else None

        def dropin_step():
            f, x = fc_host.cuda(non_blocking=True), att_host.cuda(non_blocking=True)
            m = mk_host.cuda(non_blocking=True) if mk_host is not None else None
            r = model(f, x, m, opt=kw, mode="sample")
            return r[0].cpu()

        for _ in range(3):
            seq_d = dropin_step()
        notes["dropin_token_agreement"] = float((seq_d == ref_seq.cpu()).float().mean())
        nd = max(3, a.steps // 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(nd):
            dropin_step()
        torch.cuda.synchronize()
        ms_dropin_serial = (time.perf_counter() - t0) * 1e3 / nd
        # the same class fed the way boficap_b200/data feeds it: pinned bf16 features, the next batch's H2D copy on a copy
        # stream underneath the current decode (two device buffers); `model(...)` itself stays the synchronous reference call
        copy_stream = torch.cuda.Stream()
        dbuf = [torch.empty_like(att), torch.empty_like(att)]
        evs = [torch.cuda.Event(), torch.cuda.Event()]
        f_dev = fc_host.cuda()
        m_dev = mk_host.cuda() if mk_host is not None else None

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                dbuf[i & 1].copy_(att_host_lp, non_blocking=True)
                evs[i & 1].record(copy_stream)

        def dropin_run(n):
            prefetch(0)
            seq_h = None
            for i in range(n):
                main.wait_event(evs[i & 1])
                if i + 1 < n:
                    prefetch(i + 1)
                seq_h = model(f_dev, dbuf[i & 1], m_dev, opt=kw, mode="sample")[0].cpu()
            return seq_h

        seq_d = dropin_run(3)
        notes["dropin_prefetch_token_agreement"] = float((seq_d == ref_seq.cpu()).float().mean())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dropin_run(nd)
        torch.cuda.synchronize()
        ms_dropin = (time.perf_counter() - t0) * 1e3 / nd
        del model, dbuf
        torch.cuda.empty_cache()

    if dist is not None:
        t = torch.tensor([ms, ms_e2e, ms_e2e32 or 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, m32 = t.tolist()
        ms_e2e32 = m32 if ms_e2e32 is not None else None

    # ---- per-kernel-class profile of one more step (CUDA events around every launch of the library)
    eng.set_profiling(True)
    step()
    prof = eng.get_profile()
    top = eng.get_profile_top_gemm() if a.precision == "bf16" else None
    eng.set_profiling(False)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    S = info["bounding_steps"]
    mean_regions = float(len_host.float().mean()) if len_host is not None else float(R)
    value = world * B * a.steps / (ms / 1e3)
    e2e_value = world * B * a.steps / (ms_e2e / 1e3)
    burst, sustained, hbm, how = measured_peaks()
    g = prof["gemm_tcgen05"] if a.precision == "bf16" else prof["gemm_ffma"]
    gemm_tflops = g["flops"] / (g["ms"] / 1e3) / 1e12 if g["ms"] > 0 else 0.0
    # the per-launch durations come from one profiled step (CUDA events around every launch, kernels run one at a time at
    # boost clocks): the burst figure is the matching denominator; the sustained one is reported beside it
    peak = burst if a.precision == "bf16" else 75.0
    total_ms = sum(v["ms"] for v in prof.values())
    if a.precision == "bf16" and g["ms"] > 0:
        # dominant kernel = the tcgen05 GEMM: algorithmic flops = sum of 2*M*N*K over its launches in one
        # step, duration = sum of the CUDA-event durations around each launch (library-side, launching stream)
        achieved, kname = gemm_tflops, "gemm_tc2_kernel / gemm_tc_kernel (tcgen05), all %d launches of a step" % g["launches"]
        traffic = TRAFFIC_CONSTANT["bytes"]
        dom = {"launches_per_step": g["launches"], "us_per_launch": g["ms"] / g["launches"] * 1e3,
               "gflop_per_launch": g["flops"] / g["launches"] / 1e9, "share_of_step": g["ms"] / total_ms,
               "traffic_source": TRAFFIC_CONSTANT["source"]}
        if top and top["ms"] > 0:
            dom["slowest_shape"] = {"M": top["M"], "N": top["N"], "K": top["K"], "launches": top["launches"],
                                    "us_per_launch": top["ms"] / top["launches"] * 1e3,
                                    "tflops": top["flops"] / (top["ms"] / 1e3) / 1e12}
    else:
        achieved, kname, traffic = gemm_tflops, "gemm_simt_kernel (FFMA), all launches", None
        dom = {"launches_per_step": g["launches"], "share_of_step": g["ms"] / total_ms if total_ms else None}
    useful = useful_flops_per_caption(mean_regions, S)
    roofline = {"bound": "tensor", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "peak_source": "%s bf16 burst (per-launch events of one profiled step)" % how if a.precision == "bf16" else "nominal fp32 FFMA",
                "frac_of_sustained": achieved / sustained if a.precision == "bf16" else None,
                "traffic": traffic, "dominant": dom,
                "all_gemm_launches": {"launches_per_step": g["launches"], "ms_per_step": g["ms"], "tflops": gemm_tflops,
                                      "share_of_step": g["ms"] / total_ms if total_ms else None},
                "whole_path_useful_tflops": value / world * useful / 1e12,
                "whole_path_frac_of_burst": value / world * useful / 1e12 / burst,
                "classes": {k: {"launches": v["launches"], "ms": round(v["ms"], 4),
                                "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 2) if v["ms"] > 0 else 0.0,
                                "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["ms"] > 0 else 0.0}
                            for k, v in prof.items()}}

    cpu = eager = None
    if world == 1 and not a.no_extras:
        # the GPU bar: the same algorithm as eager PyTorch on this B200
        try:
            eager = gpu_eager_bar(cfg, sd, att, masks.cuda() if masks is not None else None, a.mode)
        except Exception as ex:                 # e.g. out of memory on a shared box: reported, not fatal
            eager = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
        # the CPU baseline in a fresh process (this one is pinned to one NUMA node and holds the GPU): all host cores
        os.sched_setaffinity(0, all_cpus)
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1", "--batch", str(B),
               "--regions", str(R), "--mode", a.mode, "--calib", a.calib, "--cpu-images", "512"] + (["--adaptive"] if a.adaptive else [])
        env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
            cpu = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:
            cpu = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}

    e2e = {"value": e2e_value, "unit": "captions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": ms_e2e / a.steps, "host_feature_dtype": a.host_dtype,
           "api": "%s, pinned host buffers (%s features), %d batches in flight%s"
                  % ("bofi_stage_compact + bofi_encode_staged_compact + bofi_decode_host_async (compact features: sum(att_len) rows per batch)"
                     if e2e_feats.dim() == 2 else "bofi_sample_host_async_ex" if pipe.group == 1 else "bofi_stage_part x %d + bofi_sample_staged (batches of a group share one library call)" % pipe.group,
                     a.host_dtype, pipe.depth * pipe.group, ", caption all_gather inside the timed region" if world > 1 else "")}
    line = {"metric": METRIC, "value": value, "unit": "captions/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": a.precision, "data": "synthetic", "config": config,
            "in_flight": "%d batches (%d engine handles + streams, round robin%s)%s"
                         % (pipe.depth * pipe.group, pipe.depth,
                            "" if pipe.group == 1 else "; %d consecutive batches of %d share one library call, each with its own fill window" % (pipe.group, B),
                            "; caption all_gather (244 B/image) after every batch, inside the timed region" if world > 1 else ""),
            "e2e": e2e,
            "e2e_fp32_host": None if ms_e2e32 is None else {
                "value": world * B * a.steps / (ms_e2e32 / 1e3), "unit": "captions/s", "ms_per_step": ms_e2e32 / a.steps,
                "h2d_bytes_per_step": att_host.numel() * 4 + (len_host.numel() * 4 if len_host is not None else 0)},
            "e2e_dropin": None if ms_dropin is None else {
                "value": B / (ms_dropin / 1e3), "unit": "captions/s", "ms_per_step": ms_dropin,
                "api": "model(fc, att, masks, opt, mode='sample') of boficap_b200.captioning.models, one batch in flight; features as the "
                       "feeder delivers them (pinned %s), the next batch's H2D copy on a copy stream under the current decode, seq.cpu() per step"
                       % a.host_dtype,
                "fp32_host_serial": {"value": B / (ms_dropin_serial / 1e3), "unit": "captions/s", "ms_per_step": ms_dropin_serial,
                                     "api": "pinned fp32 host tensors, .cuda() + _sample + seq.cpu() in series (the reference eval loop's pattern, "
                                            "eval_utils.py:431-445)"}},
            "numa": numa, "notes": notes, "mean_regions": mean_regions, "gpu_launches": launches_per_step * ((a.steps + pipe.group - 1) // pipe.group), "bounding_steps": S, "fill_width": info["fill_width"],
            "nan_batch": info["nan_batch"], "mean_caption_tokens": float(out[3].sum(1).float().mean()), "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu, "gpu_eager": eager}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
