"""Image-sharded multi-GPU decode: one process per GPU, no collective on the data path.

Images are independent in the encoder and in the bounding step; the filling step of the reference couples
every row of a batch to the LAST row of that batch (stale loop index, TransformerModel.py:1871-1873), so
parity is defined per shard: each rank reproduces what the reference model would return for the rows it
sees (this is also what nn.DataParallel does to the reference, tools/train.py:99-101).  The only
communication is the final gather of captions and boxes (160 B + 244 B per image) to every rank.

`decode_fn(att_feats, att_masks) -> (seq, phrase_num, phrase_length, phrase_syn)` is the per-rank decoder
(BofiEngine on the GPU box; the tests inject the CPU oracle and run under gloo).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous, balanced split of n images: the first n % world ranks get one extra image."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_inputs(att_feats, att_masks, world, rank):
    lo, hi = shard_bounds(att_feats.shape[0], world, rank)
    return att_feats[lo:hi], (att_masks[lo:hi] if att_masks is not None else None)


def gather_captions(local, n_total, group=None):
    """all_gather of ragged per-rank results (tuple of tensors with images on dim 0) -> full-batch tensors."""
    world = dist.get_world_size(group)
    out = []
    for t in local:
        sizes = [shard_bounds(n_total, world, r) for r in range(world)]
        pad = max(hi - lo for lo, hi in sizes)
        buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        buf[: t.shape[0]] = t
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        out.append(torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], 0))
    return tuple(out)


def sample_sharded(decode_fn, att_feats, att_masks=None, group=None):
    """Every rank holds the full (host) batch description, decodes its own shard, gathers all captions."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    a, m = shard_inputs(att_feats, att_masks, world, rank)
    local = decode_fn(a, m) if a.shape[0] > 0 else None
    if local is None:
        raise ValueError("a rank received an empty shard (batch smaller than world size)")
    return gather_captions(local, att_feats.shape[0], group)


def allreduce_gradients(model, group=None, average=True):
    """Data-parallel XE training: ONE all-reduce of the flat gradient buffer the parameters' `.grad`s are views of
    (TransformerModel.train_bind), instead of one collective per parameter tensor."""
    flat = model.flat_grads()
    dist.all_reduce(flat, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    return flat
