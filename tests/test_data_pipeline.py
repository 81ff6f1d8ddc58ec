"""CPU: the input side (boficap_b200/data) -- the phrase-tensor collate against fixtures recorded
from the reference's own `Dataset.collate_func` (oracle/make_golden_collate.py), the feature reader's dispatch
(dataloader.py:24-86) and the pinned 2-byte feeder's batching."""
import os

import numpy as np
import pytest
import torch

from boficap_b200 import synth
from boficap_b200.data import FeatureReader, PinnedFeeder, collate_features, collate_phrases, collate_uic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate_uic.npz")


def test_collate_matches_reference_collate_func():
    z = np.load(GOLDEN)
    B, spi, L = int(z["B"]), int(z["spi"]), int(z["L"])
    samples = []
    for i in range(B):
        att = z["att_%d" % i]
        samples.append((att.mean(0), att, z["seq"][i * spi:(i + 1) * spi], z["pnum_in"][i * spi:(i + 1) * spi],
                        z["plen_in"][i * spi:(i + 1) * spi], z["psyn_in"][i * spi:(i + 1) * spi]))
    data = collate_uic(samples, L, spi)
    for k in ("labels", "phrase_num", "phrase_length", "phrase_syn", "extend_phrase_syn_seq", "extend_phrase_seq", "extend_phrase_seq_mask",
              "phrase", "masks"):
        ref = z["ref_" + k]
        got = data[k].numpy()
        assert got.shape == ref.shape, (k, got.shape, ref.shape)
        assert np.array_equal(got, ref), k
        assert got.dtype == ref.dtype, (k, got.dtype, ref.dtype)
    assert np.array_equal(data["att_feats"].numpy(), z["ref_att_feats"])
    assert np.array_equal(data["att_masks"].numpy(), z["ref_att_masks"])
    np.testing.assert_allclose(data["fc_feats"].numpy(), z["ref_fc_feats"], rtol=1e-6)
    assert np.array_equal(data["att_len"].numpy(), z["ref_att_masks"].sum(1).astype(np.int32))


def test_collate_masks_are_none_for_equal_region_counts_and_edge_phrases():
    att, masks, lens = collate_features([np.ones((4, 8), np.float32)] * 3)
    assert masks is None and lens.tolist() == [4, 4, 4]                     # dataloader.py:340-342
    # one caption with a single phrase filling all L slots, one with L phrases of one word
    L = 6
    seq = np.array([[4, 5, 6, 7, 8, 9], [10, 11, 12, 13, 14, 15]])
    out = collate_phrases(seq, [1, 6], np.array([[6, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1]]), np.array([[5, 0, 0, 0, 0, 0], [4, 5, 6, 4, 5, 6]]), L, 2)
    assert out["phrase_num"].tolist() == [[2, 7]]
    assert out["extend_phrase_seq"][0, 0].tolist() == [1] * 6              # the first phrase copies the BOS pseudo-phrase
    assert out["extend_phrase_seq"][0, 1].tolist() == [1, 10, 11, 12, 13, 14]
    assert out["extend_phrase_syn_seq"][0, 1].tolist() == [3, 4, 5, 6, 4, 5, 6, 0]
    m = out["extend_phrase_seq_mask"][0, 1].reshape(L, L)
    assert torch.equal(m, torch.tril(torch.ones(L, L, dtype=torch.bool)))


def test_collated_batch_feeds_the_oracle_forward():
    """The collate output has exactly the tensors `_forward` takes (TransformerModel.py:1713): run the oracle on it."""
    from boficap_b200.layout import BofiConfig
    from oracle.bofi_oracle import BofiOracle, OracleConfig
    z = np.load(GOLDEN)
    B, spi, L = 2, int(z["spi"]), int(z["L"])
    samples = [(z["att_%d" % i].mean(0), np.tile(z["att_%d" % i], (1, 256)), z["seq"][i * spi:(i + 1) * spi], z["pnum_in"][i * spi:(i + 1) * spi],
                z["plen_in"][i * spi:(i + 1) * spi], z["psyn_in"][i * spi:(i + 1) * spi]) for i in range(B)]
    data = collate_uic(samples, L, spi)
    cfg = BofiConfig(N_enc=1, N_dec=1)
    o = BofiOracle(synth.synth_state_dict(cfg, 0, None), OracleConfig(**cfg.to_dict()))
    with torch.no_grad():
        outs = o.forward_xe(data["att_feats"], data["att_masks"], data["labels"], data["phrase_num"], data["phrase_length"],
                            data["extend_phrase_syn_seq"], data["extend_phrase_seq"], data["extend_phrase_seq_mask"])
        loss, parts = o.loss_xe(outs, data["phrase_num"], data["phrase_length"], data["phrase_syn"], data["labels"])
    assert outs[2].shape == (B * spi, L, cfg.tgt_vocab) and torch.isfinite(loss)


def test_reader_dispatch_and_feeder_batches(tmp_path):
    rng = np.random.RandomState(0)
    feats = {str(i): rng.rand(int(rng.randint(3, 9)), 16).astype(np.float32) for i in range(11)}
    d_npz, d_npy = tmp_path / "att_npz", tmp_path / "att_npy"
    d_npz.mkdir()
    d_npy.mkdir()
    for k, a in feats.items():
        np.savez(d_npz / (k + ".npz"), **({"feat": a} if int(k) % 2 else {"z": a}))      # 'z': the cocotest_bu spelling (:43-44)
        np.save(d_npy / (k + ".npy"), a)
    torch.save({k: torch.from_numpy(a) for k, a in feats.items()}, str(tmp_path / "att.pth"))
    readers = [FeatureReader(str(d_npz), ".npz", in_memory=True), FeatureReader(str(d_npy), ".npy"), FeatureReader(str(tmp_path / "att.pth"))]
    for r in readers:
        for k, a in feats.items():
            assert np.array_equal(r.get(k), a)
    assert len(readers[0].cache) == 11
    for bad in ("x.lmdb", "x.h5"):
        with pytest.raises(ImportError):
            FeatureReader(str(tmp_path / bad))
    with pytest.raises(FileNotFoundError):
        FeatureReader(str(tmp_path / "missing_dir"))
    # feeder: 11 images in batches of 4 (last one short), bf16, padded to 8 regions, ring of 4 slots
    keys = [str(i) for i in range(11)]
    got = []
    for att, lens, ks in PinnedFeeder(readers[0], keys, 4, 8, feat_size=16, dtype=torch.bfloat16, depth=4, pin=False):
        assert att.dtype == torch.bfloat16 and att.shape[1:] == (8, 16) and att.shape[0] == len(ks) == lens.shape[0]
        for b, k in enumerate(ks):
            n = feats[k].shape[0]
            assert int(lens[b]) == n
            assert torch.equal(att[b, :n], torch.from_numpy(feats[k]).to(torch.bfloat16))
            assert (att[b, n:] == 0).all()
        got += ks
    assert got == keys
    assert [len(k) for *_, k in PinnedFeeder(readers[1], keys, 4, 8, feat_size=16, depth=3, pin=False, drop_last=True)] == [4, 4]
    # compact batches: only the valid regions, image after image (what pack_wrapper hands to att_embed)
    got = []
    for att, lens, ks in PinnedFeeder(readers[0], keys, 4, 8, feat_size=16, dtype=torch.bfloat16, depth=4, pin=False, compact=True):
        assert att.dim() == 2 and att.shape[0] == int(lens.sum()) and lens.shape[0] == len(ks)
        pos = 0
        for b, k in enumerate(ks):
            n = feats[k].shape[0]
            assert int(lens[b]) == n
            assert torch.equal(att[pos:pos + n], torch.from_numpy(feats[k]).to(torch.bfloat16))
            pos += n
        got += ks
    assert got == keys
