"""Double-buffered feeder over the C ABI: `depth` engine handles, each on its own CUDA stream, fed round robin.

One batch is a chain of ~330 dependent kernels: dense encoder / filling GEMMs that fill the GPU, and a bounding
loop of ~200 tiny latency-bound kernels (20 dependent steps over M = B rows) that does not.  With two batches in
flight the H2D copy and the bounding loop of one batch run underneath the dense phases of the other; every
batch still sees exactly the arithmetic (and the `last[B-1]` fill window) of a stand-alone call, so parity is
per batch, as for `nn.DataParallel` replicas of the reference.

`group` > 1 (NAIC): that many consecutive submissions are decoded by ONE library call -- their features are staged
side by side (bofi_stage_part), the call runs with bofi_set_shard(batch size) so that every batch keeps its own fill
window, and each ticket gets its slice of the outputs.  Bit for bit the results of separate calls (the fill window
is the only coupling between the rows of a batch); the bounding loop's launches are paid once per group and the
GEMMs see `group` times the rows.  A ticket of a group that is still filling launches it when waited for.

Each engine owns its packed weights and workspace (about 0.4 GB + 2.5 GB at B=1024, R=36, bf16), which is
what 180 GB of HBM is for.
"""
import torch

from .engine import BofiEngine


class _Group:
    """The submissions that will be decoded by one library call on one slot."""
    __slots__ = ("slot", "key", "parts", "event", "out", "launched", "host", "kw", "size", "rows")

    def __init__(self, slot, key, host, kw, size):
        self.slot, self.key, self.host, self.kw, self.size = slot, key, host, kw, size
        self.parts = []            # batch sizes
        self.rows = []             # compact batches: rows (valid regions) per batch
        self.event = None
        self.out = None
        self.launched = False


class Ticket:
    __slots__ = ("event", "_out", "slot", "_group", "_index", "_pipe")

    def __init__(self, event, out, slot, group=None, index=0, pipe=None):
        self.event, self._out, self.slot, self._group, self._index, self._pipe = event, out, slot, group, index, pipe

    @property
    def launched(self):
        return self._group is None or self._group.launched

    def _resolve(self):
        g = self._group
        if g is None or self._out is not None:
            return
        if not g.launched:
            self._pipe.just_launched = self._pipe._launch(g)
        lo = sum(g.parts[:self._index])
        hi = lo + g.parts[self._index]
        sn = g.kw["sample_n"]
        cut = lambda t: t[lo * sn:hi * sn] if t is not None else None
        self._out = {k: cut(v) for k, v in g.out.items()} if isinstance(g.out, dict) else tuple(cut(v) for v in g.out)
        self.event = g.event

    @property
    def out(self):
        self._resolve()
        return self._out

    def wait(self):
        self._resolve()
        self.event.synchronize()
        return self._out


class BofiPipeline:
    def __init__(self, cfg, state_dict, device=0, precision="bf16", depth=2, group=1):
        assert depth >= 1 and group >= 1
        self.device = torch.device("cuda", device)
        self.engines = [BofiEngine(cfg, device, precision).load_state_dict(state_dict) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        # host submissions: the features of a slot's NEXT call are staged on the slot's copy stream as soon as the encoder of its
        # current call has consumed the staging buffer (enc_done), i.e. underneath the bounding loop / filling decoder of that call
        self.copy_streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self.enc_done = [torch.cuda.Event() for _ in range(depth)]
        self.host_out = [None] * depth
        self.dev_out = [None] * depth
        self.n = 0
        self.group = group
        self._open = None           # the group that is still filling
        self.just_launched = []     # tickets whose library call was enqueued by the last submit_* / flush (in submission order)
        self._tickets = {}

    @property
    def depth(self):
        return len(self.engines)

    def close(self):
        for e in self.engines:
            e.close()

    def _next(self):
        slot = self.n % self.depth
        self.n += 1
        return slot

    def fork_from(self, stream):
        """Make every pipeline stream wait for what has been enqueued on `stream` so far."""
        ev = torch.cuda.Event()
        ev.record(stream)
        for st in self.streams + self.copy_streams:
            st.wait_event(ev)

    def join_into(self, stream):
        """Make `stream` wait for everything enqueued on the pipeline streams (launches a group that is still filling)."""
        self.flush()
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)

    # ---- grouped submissions ---------------------------------------------------------------------------------------------
    def flush(self):
        """Launch the group that is still filling (fewer than `group` submissions)."""
        self.just_launched = self._launch(self._open) if self._open is not None else []

    def _launch(self, g):
        """Enqueues the library call of a group; returns its tickets (submission order)."""
        if g.launched:
            return []
        eng, st = self.engines[g.slot], self.streams[g.slot]
        total = sum(g.parts)
        code, have_len, R = g.key[0], g.key[1], g.key[2]
        kw = g.kw
        if g.host:
            staged = torch.cuda.Event()
            staged.record(self.copy_streams[g.slot])
        with torch.cuda.stream(st):
            eng.set_shard(g.parts[0] if len(g.parts) > 1 else 0)
            if g.host:
                st.wait_event(staged)
                if g.rows:
                    eng.encode_staged_compact(code, sum(g.rows), total, R)
                else:
                    eng.encode_staged(code, have_len, total, R)
                self.enc_done[g.slot].record(st)            # the staging buffers may be refilled from here on
                g.out = eng.decode_host(kw["mode"], kw["sample_n"], kw["output_logsoftmax"], out=self.host_out[g.slot],
                                        want_logprobs=kw["want_logprobs"])
                self.host_out[g.slot] = g.out
            else:
                eng.encode_staged(code, have_len, total, R)
                g.out = eng.decode(kw["mode"], kw["sample_n"], kw["output_logsoftmax"], kw["want_logprobs"],
                                   out=self.dev_out[g.slot] if kw["reuse_outputs"] else None)
                if kw["reuse_outputs"]:
                    self.dev_out[g.slot] = g.out
            eng.set_shard(0)
            g.event = torch.cuda.Event()
            g.event.record(st)
        g.launched = True
        if self._open is g:
            self._open = None
        return self._tickets.pop(id(g), [])

    def _submit_grouped(self, host, att_feats, att_len, kw, size, compact_regions=0):
        from .engine import _feat_code
        compact = compact_regions > 0                       # att_feats [sum(att_len), F]: only the valid regions
        B, R = (int(att_len.shape[0]), compact_regions) if compact else att_feats.shape[:2]
        key = (_feat_code(att_feats), att_len is not None, R, B, host, size, compact, tuple(sorted(kw.items())))
        launched = []
        g = self._open
        if g is not None and g.key != key:                  # another shape / mode: the open group goes as it is
            launched += self._launch(g)
            g = None
        if g is None:
            g = _Group(self._next(), key, host, kw, size)
            self._open = g
            self._tickets[id(g)] = []
        eng = self.engines[g.slot]
        st = self.copy_streams[g.slot] if host else self.streams[g.slot]
        with torch.cuda.stream(st):                         # the copy starts now, underneath whatever is running
            if host and not g.parts:
                st.wait_event(self.enc_done[g.slot])        # (a no-op before the slot's first call)
            if compact:
                eng.stage_compact(att_feats, att_len, R, row0=sum(g.rows), image0=sum(g.parts), total_images=B * size)
                g.rows.append(int(att_feats.shape[0]))
            else:
                eng.stage_part(att_feats, att_len, sum(g.parts), B * size)
        g.parts.append(B)
        t = Ticket(None, None, g.slot, g, len(g.parts) - 1, self)
        self._tickets[id(g)].append(t)
        if len(g.parts) == g.size:
            launched += self._launch(g)
        self.just_launched = launched
        return t

    def submit_device(self, att_feats, att_len=None, mode="NAIC", sample_n=1, output_logsoftmax=1, want_logprobs=True, reuse_outputs=False):
        """Device-resident inputs: encode + decode enqueued on the next slot's stream.  reuse_outputs: write into the
        slot's output tensors of the previous round (valid until the slot is reused; no allocation per call)."""
        if self.group > 1 and mode == "NAIC":
            return self._submit_grouped(False, att_feats, att_len, dict(mode=mode, sample_n=sample_n, output_logsoftmax=output_logsoftmax,
                                                                         want_logprobs=want_logprobs, reuse_outputs=reuse_outputs), self.group)
        launched = self._launch(self._open) if self._open is not None else []
        slot = self._next()
        eng, st = self.engines[slot], self.streams[slot]
        with torch.cuda.stream(st):
            eng.encode(att_feats, att_len)
            out = eng.decode(mode, sample_n, output_logsoftmax, want_logprobs, out=self.dev_out[slot] if reuse_outputs else None)
            if reuse_outputs:
                self.dev_out[slot] = out
            ev = torch.cuda.Event()
            ev.record(st)
        t = Ticket(ev, out, slot)
        self.just_launched = launched + [t]
        return t

    def submit_host(self, att_feats, att_len=None, mode="NAIC", sample_n=1, output_logsoftmax=1, want_logprobs=False):
        """Pinned host inputs -> pinned host outputs (owned by the slot, valid until the slot is reused)."""
        if att_feats.is_pinned():
            # staged on the slot's copy stream, decoded by a call shared with the other batches of its group (NAIC)
            if att_len is not None and not att_len.is_cuda and not att_len.is_pinned():
                att_len = att_len.to(torch.int32).pin_memory()       # (4 bytes per image; asynchronous copies need pinned memory)
            return self._submit_grouped(True, att_feats, att_len, dict(mode=mode, sample_n=sample_n, output_logsoftmax=output_logsoftmax,
                                                                        want_logprobs=want_logprobs, reuse_outputs=True),
                                        self.group if mode == "NAIC" else 1)
        launched = self._launch(self._open) if self._open is not None else []
        slot = self._next()
        eng, st = self.engines[slot], self.streams[slot]
        with torch.cuda.stream(st):
            out = eng.sample_host(att_feats, att_len, mode, sample_n, output_logsoftmax, out=self.host_out[slot],
                                  want_logprobs=want_logprobs, sync=False)
            self.host_out[slot] = out
            ev = torch.cuda.Event()
            ev.record(st)
        t = Ticket(ev, out, slot)
        self.just_launched = launched + [t]
        return t

    def submit_host_compact(self, att_compact, att_len, max_regions, mode="NAIC", sample_n=1, output_logsoftmax=1, want_logprobs=False):
        """Pinned COMPACT host features ([sum(att_len), F]: only the valid regions, image after image; boficap_b200/data:
        PinnedFeeder(compact=True)) -> pinned host outputs.  sum(att_len) instead of B * R rows cross PCIe and att_embed runs on
        them directly.  Staged on the slot's copy stream and grouped like submit_host (the compact rows of a group's batches
        follow each other in the staging buffer)."""
        assert att_compact.is_pinned() and att_compact.dim() == 2 and att_len is not None
        if not att_len.is_cuda and not att_len.is_pinned():
            att_len = att_len.to(torch.int32).pin_memory()
        return self._submit_grouped(True, att_compact, att_len, dict(mode=mode, sample_n=sample_n, output_logsoftmax=output_logsoftmax,
                                                                      want_logprobs=want_logprobs, reuse_outputs=True),
                                    self.group if mode == "NAIC" else 1, compact_regions=int(max_regions))
