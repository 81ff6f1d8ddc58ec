"""Aggregate host-to-device ceiling of the box: every rank copies one pinned batch of region features (B x R x 2048, bf16 by
default) to its GPU in a loop with plain cudaMemcpyAsync (one call per copy), all ranks at once.  Run under torchrun with
the same rank count as the bench; rank 0 prints one JSON line.  This is the denominator of the N-GPU end-to-end numbers."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--regions", type=int, default=36)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--numa", action="store_true", help="bind to the GPU's NUMA node first (bench.py: numa_bind)")
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    numa = None
    if a.numa:
        from bench import numa_bind
        numa = numa_bind(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dt = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    host = torch.empty(a.batch, a.regions, 2048, dtype=dt).pin_memory()
    host.fill_(1)
    dev = torch.empty_like(host, device="cuda")
    for _ in range(3):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        dev.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    nbytes = host.numel() * host.element_size()
    if rank == 0:
        print(json.dumps({"what": "pinned H2D ceiling, %d rank(s) copying at once" % world, "bytes_per_copy": nbytes, "iters": a.iters,
                          "ms_per_copy": ms / a.iters, "gbs_per_gpu": nbytes / (ms / a.iters) / 1e6, "gbs_aggregate": world * nbytes / (ms / a.iters) / 1e6,
                          "captions_per_s_ceiling": world * a.batch / (ms / a.iters) * 1e3, "dtype": a.dtype, "numa": numa}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
