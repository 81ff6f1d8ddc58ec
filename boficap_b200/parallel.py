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


def gather_captions_packed(seq, phrase_num, phrase_length, phrase_syn, group=None):
    """The final caption gather as ONE collective for equal shards (bench.py, weak scaling): tokens and boxes of a shard
    packed into one int32 row per image ([L | 1 | L | L] = 244 B at L = 20), all_gather_into_tensor, unpacked views.
    Returns (seq i64 [W*rows, L], phrase_num i32, phrase_length i32, phrase_syn i64) of the whole job on every rank."""
    L = seq.shape[1]
    packed = torch.cat([seq.to(torch.int32), phrase_num.to(torch.int32)[:, None], phrase_length.to(torch.int32),
                        phrase_syn.to(torch.int32)], 1).contiguous()
    world = dist.get_world_size(group)
    out = torch.empty((world * packed.shape[0], packed.shape[1]), dtype=torch.int32, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    return out[:, :L].long(), out[:, L], out[:, L + 1:2 * L + 1], out[:, 2 * L + 1:].long()


def sample_sharded(decode_fn, att_feats, att_masks=None, group=None):
    """Every rank holds the full (host) batch description, decodes its own shard, gathers all captions."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    a, m = shard_inputs(att_feats, att_masks, world, rank)
    local = decode_fn(a, m) if a.shape[0] > 0 else None
    if local is None:
        raise ValueError("a rank received an empty shard (batch smaller than world size)")
    return gather_captions(local, att_feats.shape[0], group)


class OverlappedGradReduce:
    """Data-parallel XE training with the gradient all-reduce overlapped with the backward pass.  The flat gradient buffer
    is cut where the library's backward pass finishes first: everything from the decoders on (embeddings, generator, bounding
    head, the memory K/V projections) is final before the encoder's backward pass starts (bofi_train_set_grad_event), so
    that part -- `buckets` NCCL calls so that the first bytes move early -- is reduced on a side stream underneath it; the
    encoder / att_embed part follows on the training stream.  `finish()` joins the side stream."""

    def __init__(self, model, group=None, buckets=4, layer_buckets=False):
        """layer_buckets: the encoder part as well is reduced layer by layer underneath the backward pass of the layers below
        (bofi_train_set_layer_event); only att_embed (4 MB) is left for after the backward pass.  Same gradients
        (tools/check_grad_reduce.py); measured at 8 GPUs: 23.99 ms per step against 23.93 ms with ONE encoder bucket after the
        backward pass -- what an 8-GPU step costs over a 1-GPU step (+0.6 ms) is not the tail all-reduce -- so it is off."""
        self.model, self.group = model, group
        eng = model._engine
        self.flat = model.flat_grads()
        layout = eng.param_layout()
        self.cut = layout["model.decoder.layers.0.self_attn.linears.0.weight"][0]
        self.layers = []                         # (event, gradient slice) from the last encoder layer down to the first
        if layer_buckets:
            n_enc = model.cfg.N_enc
            starts = [layout["model.encoder.layers.%d.self_attn.linears.0.weight" % l][0] for l in range(n_enc)] + [self.cut]
            for l in range(n_enc - 1, -1, -1):
                ev = torch.cuda.Event()
                ev.record()
                eng.train_set_layer_event(l, ev)
                self.layers.append((ev, self.flat[starts[l]:starts[l + 1]]))
            self.front_cut = starts[0]
        n_back = self.flat.numel() - self.cut
        step = -(-n_back // max(1, buckets))
        self.back = [self.flat[self.cut + i:min(self.cut + i + step, self.flat.numel())] for i in range(0, n_back, step)]
        self.front = self.flat[:self.front_cut] if self.layers else self.flat[:self.cut]
        self.side = torch.cuda.Stream(self.flat.device)
        self.event = torch.cuda.Event()
        self.event.record()                      # creates the handle the library records into
        eng.train_set_grad_event(self.event)

    def reduce(self):
        """Call right after the (asynchronous) backward pass has been enqueued on the current stream."""
        main = torch.cuda.current_stream(self.flat.device)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.event)     # recorded by the library inside the backward pass that was just enqueued
            for b in self.back:
                dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group)
            for ev, b in self.layers:            # encoder layers, last to first, as their backward passes finish
                self.side.wait_event(ev)
                dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.group)
            done = torch.cuda.Event()
            done.record(self.side)
        dist.all_reduce(self.front, op=dist.ReduceOp.AVG, group=self.group)
        main.wait_event(done)

    def close(self):
        self.model._engine.train_set_grad_event(None)
        for l in range(len(self.layers)):
            self.model._engine.train_set_layer_event(l, None)


def allreduce_gradients(model, group=None, average=True):
    """Data-parallel XE training: ONE all-reduce of the flat gradient buffer the parameters' `.grad`s are views of
    (TransformerModel.train_bind), instead of one collective per parameter tensor."""
    flat = model.flat_grads()
    dist.all_reduce(flat, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    return flat
