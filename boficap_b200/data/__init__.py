"""Input side of the hot path: the region-feature reader (captioning/data/dataloader.py:24-86),
the phrase-tensor collate of train_mode UIC (:231-452) and a pinned, 2-byte, double-buffered feeder that hands batches to
bofi_sample_host_async_ex / the XE step.  Host-side code: numpy for the integer tensors, torch only as the pinned-memory
allocator."""
from .collate import collate_features, collate_phrases, collate_uic  # noqa: F401
from .feeder import PinnedFeeder  # noqa: F401
from .reader import FeatureReader  # noqa: F401
