"""Batch construction of the reference's `Dataset.collate_func` (captioning/data/dataloader.py:231-452) for train_mode UIC
and preprocess_mode 'phrase' (configs/uic_sd.yml): the padded region features + prefix masks (:333-342) and the phrase
tensors the teacher-forced forward consumes (:343-428).  numpy restatement, row loops vectorised where the reference
loops per caption; tests/test_data_pipeline.py holds it to fixtures recorded from the reference's own function
(oracle/make_golden_collate.py)."""
import numpy as np
import torch


def collate_features(att_list, out=None, lens_out=None, dtype=torch.float32, pad_to=None):
    """att_list: per-image [regions_i, feat] arrays.  Returns (att_feats [B, R, feat], att_masks [B, R] f32 or None, att_len
    i32 [B]) with R = max regions (or `pad_to`); `att_masks` is None when every image has R regions, as :340-342 does.
    `out` / `lens_out`: preallocated (pinned) buffers to fill instead of allocating."""
    B = len(att_list)
    R = max(a.shape[0] for a in att_list) if pad_to is None else pad_to
    F = att_list[0].shape[1]
    if out is None:
        out = torch.zeros(B, R, F, dtype=dtype)
    else:
        out = out[:B, :R]
        out.zero_()
    lens = np.zeros(B, dtype=np.int32)
    for i, a in enumerate(att_list):
        n = min(a.shape[0], R)
        out[i, :n].copy_(torch.from_numpy(np.ascontiguousarray(a[:n])))
        lens[i] = n
    att_len = torch.from_numpy(lens) if lens_out is None else lens_out[:B].copy_(torch.from_numpy(lens))
    masks = None
    if int(lens.min()) != R:
        masks = (torch.arange(R)[None, :] < torch.from_numpy(lens)[:, None]).float()
    return out, masks, att_len


def collate_phrases(seq, phrase_num, phrase_length, phrase_syn, seq_length, seq_per_img, pad_idx=0, bos_idx=1, eos_idx=2,
                    len_idx=3):
    """The phrase tensors of :296-428.  Inputs are the per-caption rows the dataset stores (stacked over the batch):
      seq [N, L] words; phrase_num [N]; phrase_length [N, L]; phrase_syn [N, L]   (N = images * seq_per_img, L = seq_length)
    Returns the reference's dict entries `labels`, `phrase_num`, `phrase_length`, `phrase_syn`, `extend_phrase_syn_seq`,
    `extend_phrase_seq`, `extend_phrase_seq_mask`, `phrase`, `masks` with the reference's shapes and dtypes."""
    seq = np.asarray(seq, dtype=np.int64)
    N, L = seq.shape[0], seq_length
    pn = np.asarray(phrase_num, dtype=np.int64).reshape(-1)
    pl = np.asarray(phrase_length, dtype=np.int64)
    ps = np.asarray(phrase_syn, dtype=np.int64)
    labels = np.zeros((N, L + 2), dtype=np.int64)               # :296-302  bos, words, eos
    labels[:, 1:L + 1] = seq[:, :L]
    labels[:, 0] = bos_idx
    labels[:, L + 1] = eos_idx
    plen = np.zeros((N, L + 2), dtype=np.int64)                 # :361-365
    psyn = np.zeros((N, L + 2), dtype=np.int64)
    plen[:, 0] = 1
    psyn[:, 0] = bos_idx
    ext_syn = np.zeros((N, L + 2), dtype=np.int64)              # :354-357
    ext_syn[:, 0] = len_idx
    ext_seq = np.zeros((N, L), dtype=np.int64)
    ext_mask = np.zeros((N, L, L), dtype=bool)
    slot = np.arange(L)
    for n in range(N):
        k = int(pn[n])
        plen[n, 1:k + 1] = pl[n, :k]                            # :373-375
        psyn[n, 1:k + 1] = ps[n, :k]
        psyn[n, k + 1] = eos_idx
        lens = pl[n, :k]
        ends = np.cumsum(lens)
        total = int(ends[-1]) if k else 0
        # :376-379  syn label of every word slot: slot s (1-based inside the row) belongs to phrase searchsorted(ends, s)
        if total:
            ext_syn[n, 1:1 + total] = ps[n, np.searchsorted(ends, slot[:total], side="right")]
        # :381-399  position-wise copy of the previous phrase's words; :400 the phrase-block-causal mask
        seq_last, phrase_last = 0, 0
        for j in range(1, k + 1):
            cur, prev = int(plen[n, j]), int(plen[n, j - 1])
            if cur <= prev:
                pre_pad = prev - cur
                ext_seq[n, phrase_last:phrase_last + cur] = labels[n, seq_last + pre_pad:seq_last + pre_pad + cur]
            else:
                pre_less, times = prev - (cur % prev), cur // prev
                reps = np.where(np.arange(prev) < pre_less, times, times + 1)
                ext_seq[n, phrase_last:phrase_last + cur] = np.repeat(labels[n, seq_last:seq_last + prev], reps)[:L - phrase_last]
            ext_mask[n, phrase_last:, :phrase_last + cur] = True
            seq_last += prev
            phrase_last += cur
    pnum = pn + 1                                               # :350  (+ the BOS pseudo-phrase)
    # :403-425  `phrase`: the words regrouped so that phrase j of every caption starts at the same column
    max_pn = int(pn.max()) + 2
    max_len = plen[:, :max_pn].max(0)
    start = np.concatenate([[0], np.cumsum(max_len)[:-1]])
    phrase = np.full((N, int(max_len.sum())), pad_idx, dtype=np.int64)
    for n in range(N):
        last = 0
        for j in range(int(pnum[n])):
            c = int(plen[n, j])
            phrase[n, start[j]:start[j] + c] = labels[n, last:last + c]
            last += c
    B = N // seq_per_img
    r = lambda a: torch.from_numpy(a).reshape(B, seq_per_img, *a.shape[1:])
    return dict(labels=r(labels), phrase_num=r(pnum), phrase_length=r(plen), phrase_syn=r(psyn), extend_phrase_syn_seq=r(ext_syn),
                extend_phrase_seq=r(ext_seq), extend_phrase_seq_mask=r(ext_mask.reshape(N, L * L)), phrase=r(phrase),
                masks=r(phrase != pad_idx))


def collate_uic(samples, seq_length, seq_per_img, dtype=torch.float32, **idx):
    """`collate_func` for a list of samples (fc [feat], att [regions, feat], seq [spi, L], phrase_num [spi],
    phrase_length [spi, L], phrase_syn [spi, L]) -> the reference's batch dict (features in `dtype`)."""
    fc = torch.from_numpy(np.stack([np.asarray(s[0], dtype=np.float32) for s in samples]))
    att, masks, att_len = collate_features([np.asarray(s[1], dtype=np.float32) for s in samples], dtype=dtype)
    data = collate_phrases(np.vstack([s[2] for s in samples]), np.concatenate([np.asarray(s[3]).reshape(-1) for s in samples]),
                           np.vstack([s[4] for s in samples]), np.vstack([s[5] for s in samples]), seq_length, seq_per_img, **idx)
    data.update(fc_feats=fc, att_feats=att, att_masks=masks, att_len=att_len)
    return data
