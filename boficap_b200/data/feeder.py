"""Pinned, 2-byte, double-buffered feature feeder.

The reference's loader yields fp32 numpy batches that `eval_utils.py:431-437` / `tools/train.py:205-207` copy to the GPU
from pageable memory.  Here a background thread reads the images of the NEXT batches (FeatureReader), converts them to
the feeder's element type (bf16 by default: half the host-to-device bytes, and a bf16 engine's att_embed GEMM reads them in
place) while packing them into a ring of page-locked batch buffers; the consumer gets `(att_feats, att_len, keys)` whose
tensors go straight to `bofi_sample_host_async_ex` (BofiPipeline.submit_host) or `.cuda(non_blocking=True)`.

Of the `depth` slots one is being filled and at least one waits in the ready queue, so a batch stays valid until
`depth - 2` newer ones have been handed out: use depth = pipeline depth + 2 (5 for the default three batches in flight).
"""
import queue
import threading

import torch


class PinnedFeeder:
    def __init__(self, reader, keys, batch_size, max_regions, feat_size=2048, dtype=torch.bfloat16, depth=5, drop_last=False, pin=None,
                 compact=False):
        """compact=True: a batch is `(att_compact [sum(att_len), F], att_len [n], keys)` -- only the valid regions, image after image,
        what pack_wrapper (AttModel.py:33-51) feeds att_embed -- for BofiPipeline.submit_host_compact / bofi_stage_compact: with
        10..100 regions per image 44 % fewer bytes cross PCIe than with the padded [B, R, F] batch."""
        self.reader, self.keys, self.B, self.R = reader, list(keys), batch_size, max_regions
        self.dtype, self.depth, self.drop_last, self.compact = dtype, max(2, depth), drop_last, compact
        pin = torch.cuda.is_available() if pin is None else pin
        mk = lambda *shape, dt: torch.zeros(*shape, dtype=dt).pin_memory() if pin else torch.zeros(*shape, dtype=dt)
        self.slots = [(mk(batch_size, max_regions, feat_size, dt=dtype), mk(batch_size, dt=torch.int32)) for _ in range(self.depth)]
        self.free = queue.Queue()
        for i in range(self.depth):
            self.free.put(i)
        self.ready = queue.Queue(maxsize=self.depth)
        self.held = []
        self.thread = threading.Thread(target=self._produce, daemon=True)
        self.thread.start()

    def _produce(self):
        try:
            for lo in range(0, len(self.keys), self.B):
                keys = self.keys[lo:lo + self.B]
                if len(keys) < self.B and self.drop_last:
                    break
                i = self.free.get()
                att, lens = self.slots[i]
                n = len(keys)
                if self.compact:
                    rows = att.view(-1, att.shape[-1])          # the same pinned memory as [B * R, F]
                    pos = 0
                    for b, k in enumerate(keys):
                        cnt = self.reader.get_into(k, rows[pos:pos + self.R])
                        lens[b] = cnt
                        pos += cnt
                    self.ready.put((i, n, keys, pos))
                    continue
                for b, k in enumerate(keys):
                    cnt = self.reader.get_into(k, att[b])
                    if cnt < self.R:
                        att[b, cnt:].zero_()                  # padded regions are zero, as dataloader.py:333-336 leaves them
                    lens[b] = cnt
                self.ready.put((i, n, keys))
            self.ready.put(None)
        except Exception as ex:                                # surfaced to the consumer
            self.ready.put(ex)

    def __iter__(self):
        return self

    def __next__(self):
        item = self.ready.get()
        if item is None:
            raise StopIteration
        if isinstance(item, Exception):
            raise item
        i, n, keys = item[:3]
        self.held.append(i)
        if len(self.held) >= self.depth - 1:                    # the oldest batch handed out is finished by now
            self.free.put(self.held.pop(0))
        att, lens = self.slots[i]
        if self.compact:
            return att.view(-1, att.shape[-1])[:item[3]], lens[:n], keys
        return att[:n], lens[:n], keys
