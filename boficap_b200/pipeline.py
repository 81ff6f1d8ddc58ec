"""Double-buffered feeder over the C ABI: `depth` engine handles, each on its own CUDA stream, fed round robin.

One batch is a chain of ~330 dependent kernels: dense encoder / filling GEMMs that fill the GPU, and a bounding
loop of ~200 tiny latency-bound kernels (20 dependent steps over M = B rows) that does not.  With two batches in
flight the H2D copy and the bounding loop of one batch run underneath the dense phases of the other; every
batch still sees exactly the arithmetic (and the `last[B-1]` fill window) of a stand-alone call, so parity is
per batch, as for `nn.DataParallel` replicas of the reference.

Each engine owns its packed weights and workspace (about 0.4 GB + 2.5 GB at B=1024, R=36, bf16), which is
what 180 GB of HBM is for.
"""
import torch

from .engine import BofiEngine


class Ticket:
    __slots__ = ("event", "out", "slot")

    def __init__(self, event, out, slot):
        self.event, self.out, self.slot = event, out, slot

    def wait(self):
        self.event.synchronize()
        return self.out


class BofiPipeline:
    def __init__(self, cfg, state_dict, device=0, precision="bf16", depth=2):
        assert depth >= 1
        self.device = torch.device("cuda", device)
        self.engines = [BofiEngine(cfg, device, precision).load_state_dict(state_dict) for _ in range(depth)]
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]
        self.host_out = [None] * depth
        self.dev_out = [None] * depth
        self.n = 0

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
        for st in self.streams:
            st.wait_event(ev)

    def join_into(self, stream):
        """Make `stream` wait for everything enqueued on the pipeline streams."""
        for st in self.streams:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)

    def submit_device(self, att_feats, att_len=None, mode="NAIC", sample_n=1, output_logsoftmax=1, want_logprobs=True, reuse_outputs=False):
        """Device-resident inputs: encode + decode enqueued on the next slot's stream.  reuse_outputs: write into the
        slot's output tensors of the previous round (valid until the slot is reused; no allocation per call)."""
        slot = self._next()
        eng, st = self.engines[slot], self.streams[slot]
        with torch.cuda.stream(st):
            eng.encode(att_feats, att_len)
            out = eng.decode(mode, sample_n, output_logsoftmax, want_logprobs, out=self.dev_out[slot] if reuse_outputs else None)
            if reuse_outputs:
                self.dev_out[slot] = out
            ev = torch.cuda.Event()
            ev.record(st)
        return Ticket(ev, out, slot)

    def submit_host(self, att_feats, att_len=None, mode="NAIC", sample_n=1, output_logsoftmax=1, want_logprobs=False):
        """Pinned host inputs -> pinned host outputs (owned by the slot, valid until the slot is reused)."""
        slot = self._next()
        eng, st = self.engines[slot], self.streams[slot]
        with torch.cuda.stream(st):
            out = eng.sample_host(att_feats, att_len, mode, sample_n, output_logsoftmax, out=self.host_out[slot],
                                  want_logprobs=want_logprobs, sync=False)
            self.host_out[slot] = out
            ev = torch.cuda.Event()
            ev.record(st)
        return Ticket(ev, out, slot)
