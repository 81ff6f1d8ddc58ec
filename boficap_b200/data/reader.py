"""Region-feature reader: the dispatch of the reference's HybridLoader (captioning/data/dataloader.py:24-86).

  db_path a directory      -> one file per image, `<key><ext>`; ext '.npy' (an array) or '.npz' (key 'feat', or 'z' as the
                              reference's cocotest_bu files have it, :43-44)
  db_path ending in '.pth' -> a {key: array / tensor} dictionary saved with torch.save (ext ignored, :50-54)
  db_path ending in '.lmdb' / 'h5' -> the reference reads these through `lmdbdict` / `h5py`; neither package is in this
                              image, so they are imported lazily and their absence is an ImportError that says so.
`in_memory=True` keeps the raw bytes of every file that has been read (:58-60, :79-80).  `get(key)` returns float32
`[regions, feat]` like the reference; `get_into(key, out, dtype)` decodes straight into a row block of a (pinned) batch
buffer in the feeder's element type, which is what PinnedFeeder uses.
"""
import io
import os

import numpy as np
import torch


class FeatureReader:
    def __init__(self, db_path, ext=".npz", in_memory=False):
        self.db_path, self.ext, self.in_memory = db_path, ext, in_memory
        self.cache = {} if in_memory else None
        if db_path.endswith(".lmdb"):
            try:
                from lmdbdict import lmdbdict
                from lmdbdict.methods import DUMPS_FUNC, LOADS_FUNC
            except ImportError as ex:
                raise ImportError("reading %s needs the `lmdbdict` package (dataloader.py:46-49), which is not installed" % db_path) from ex
            self.kind = "lmdb"
            self.db = lmdbdict(db_path, unsafe=True)
            self.db._key_dumps = DUMPS_FUNC["ascii"]
            self.db._value_loads = LOADS_FUNC["identity"]
        elif db_path.endswith(".pth"):
            self.kind = "pth"
            self.db = torch.load(db_path, map_location="cpu")
        elif db_path.endswith("h5"):
            try:
                import h5py
            except ImportError as ex:
                raise ImportError("reading %s needs the `h5py` package (dataloader.py:55-57), which is not installed" % db_path) from ex
            self.kind = "h5"
            self.db = h5py.File(db_path, "r")
        else:
            if not os.path.isdir(db_path):
                raise FileNotFoundError("feature directory %s does not exist" % db_path)
            self.kind = "dir"

    def _decode(self, raw):
        if self.kind == "pth":
            return raw.numpy() if isinstance(raw, torch.Tensor) else np.asarray(raw)
        if self.kind == "h5":
            return np.array(raw).astype("float32")
        arr = np.load(io.BytesIO(raw))
        if isinstance(arr, np.lib.npyio.NpzFile):
            arr = arr["feat"] if "feat" in arr else arr["z"]
        return arr

    def _raw(self, key):
        if self.cache is not None and key in self.cache:
            return self.cache[key]
        if self.kind in ("lmdb", "pth", "h5"):
            raw = self.db[key]
        else:
            with open(os.path.join(self.db_path, str(key) + self.ext), "rb") as f:
                raw = f.read()
        if self.cache is not None:
            self.cache[key] = raw
        return raw

    def get(self, key):
        """float32 [regions, feat] (2-D features are flattened to rows like dataloader.py:485: reshape(-1, shape[-1]))."""
        a = np.asarray(self._decode(self._raw(key)), dtype=np.float32)
        return a.reshape(-1, a.shape[-1])

    def get_into(self, key, out):
        """Decode image `key` into the leading rows of `out` (a [R_max, feat] row block of a torch batch buffer, any float
        dtype: the conversion to the feeder's 2-byte type happens in this copy).  Returns the region count."""
        a = torch.from_numpy(np.ascontiguousarray(self.get(key)))
        n = min(a.shape[0], out.shape[0])
        out[:n].copy_(a[:n])
        return n
