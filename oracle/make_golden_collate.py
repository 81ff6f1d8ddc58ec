"""Golden vectors of the reference's batch construction (captioning/data/dataloader.py:231-452, `Dataset.collate_func`,
train_mode UIC / preprocess_mode 'phrase') from the UNMODIFIED reference function (build container only; outputs committed
under tests/golden/collate_uic.npz).

TEST INFRASTRUCTURE.  Usage:  python oracle/make_golden_collate.py

dataloader.py cannot be used as a Dataset here (h5py / lmdbdict are not installed and `Dataset.__init__` reads attributes
that are never assigned), but `collate_func` itself is plain numpy: the two missing imports are stubbed
and the function is called unbound on a namespace carrying exactly the attributes it reads.
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_ROOT = os.environ.get("BOFI_REFERENCE_ROOT", "/root/reference")


def reference_collate():
    for name in ("h5py", "lmdbdict", "lmdbdict.methods"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.lmdbdict = object
            m.DUMPS_FUNC, m.LOADS_FUNC = {}, {}
            sys.modules[name] = m
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from captioning.data import dataloader
    return dataloader.Dataset.collate_func


def random_samples(rng, B, spi, L, feat, rmin, rmax, vocab):
    """Per-image samples in the layout `Dataset.__getitem__` returns (dataloader.py:454-520)."""
    samples = []
    for ix in range(B):
        regions = int(rng.randint(rmin, rmax + 1))
        att = rng.rand(regions, feat).astype("float32")
        seq = np.zeros((spi, L), dtype="int")
        pnum = np.zeros(spi, dtype="int")
        plen = np.zeros((spi, L), dtype="int")
        psyn = np.zeros((spi, L), dtype="int")
        for q in range(spi):
            lens, total = [], 0
            for _ in range(int(rng.randint(1, 9))):
                ln = int(rng.randint(1, 7))
                if total + ln > L:
                    break
                lens.append(ln)
                total += ln
            k = len(lens)
            seq[q, :total] = rng.randint(4, vocab + 4, size=total)
            pnum[q] = k
            plen[q, :k] = lens
            psyn[q, :k] = rng.randint(4, 7, size=k)
        samples.append((att.mean(0), att, seq, pnum, plen, psyn, ix, ix, False))
    return samples


def main():
    collate = reference_collate()
    rng = np.random.RandomState(20)
    B, spi, L, feat = 6, 5, 16, 8
    samples = random_samples(rng, B, spi, L, feat, 3, 9, 9487)
    fake = types.SimpleNamespace(seq_per_img=spi, pp_mode="phrase", seq_length=L, bos_idx=1, eos_idx=2, len_idx=3, pad_idx=0,
                                 train_mode="UIC", h5_label_file=True, label=np.zeros((B * spi, L), dtype="int"),
                                 label_start_ix=np.arange(B) * spi + 1, label_end_ix=(np.arange(B) + 1) * spi,
                                 info={"images": [{"id": i} for i in range(B)]}, split_ix={"train": list(range(B))})
    data = collate(fake, samples, "train")
    out = {"B": B, "spi": spi, "L": L}
    for i, s in enumerate(samples):
        out["att_%d" % i] = s[1]
    out["seq"] = np.vstack([s[2] for s in samples])
    out["pnum_in"] = np.concatenate([s[3] for s in samples])
    out["plen_in"] = np.vstack([s[4] for s in samples])
    out["psyn_in"] = np.vstack([s[5] for s in samples])
    for k in ("fc_feats", "att_feats", "att_masks", "labels", "phrase_num", "phrase_length", "phrase_syn", "extend_phrase_syn_seq",
              "extend_phrase_seq", "extend_phrase_seq_mask", "phrase", "masks"):
        out["ref_" + k] = data[k].numpy()
    path = os.path.join(ROOT, "tests", "golden", "collate_uic.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and k.startswith("ref_")})


if __name__ == "__main__":
    main()
