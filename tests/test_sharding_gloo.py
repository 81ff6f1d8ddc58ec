"""CPU, world_size 2 over gloo: the image-sharded decode driver (boficap_b200/parallel.py).  The per-rank
decoder is the CPU oracle here; on the GPU box it is the CUDA engine (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from boficap_b200 import parallel, synth
from boficap_b200.layout import BofiConfig


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 64, 1000, 4096):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.bofi_oracle import BofiOracle, OracleConfig
    cfg = BofiConfig()
    sd = synth.synth_state_dict(cfg, 0, "s_real")
    oracle = BofiOracle(sd, OracleConfig(**cfg.to_dict()))
    fc, att, masks = synth.synth_inputs(7, 36, seed=7, adaptive=True)

    def decode(a, m):
        seq, _, pnum, plen, psyn, _ = oracle.sample(None, a, m, {"train_mode": "NAIC"})
        return seq, pnum, plen, psyn

    full = parallel.sample_sharded(decode, att, masks)
    # per-shard reference: what the model returns for exactly the rows a rank sees
    lo, hi = parallel.shard_bounds(7, world, rank)
    mine = decode(att[lo:hi], masks[lo:hi])
    ok = all(torch.equal(f[lo:hi], m) for f, m in zip(full, mine))
    ok = ok and full[0].shape == (7, 20) and full[1].shape == (7,)
    # boxes never depend on the batch composition: they equal the unsharded decode
    whole = decode(att, masks)
    ok = ok and torch.equal(full[2], whole[2]) and torch.equal(full[3], whole[3])
    # the one-collective packed gather of equal shards (bench.py's timed region): every rank decodes 4 images
    a4, m4 = att[rank * 3:rank * 3 + 4], masks[rank * 3:rank * 3 + 4]
    mine4 = decode(a4, m4)
    packed = parallel.gather_captions_packed(*mine4)
    ok = ok and packed[0].shape == (4 * world, 20) and packed[0].dtype == torch.int64 and packed[3].dtype == torch.int64
    ok = ok and all(torch.equal(p[rank * 4:rank * 4 + 4].to(m.dtype), m) for p, m in zip(packed, mine4))
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_decode_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
