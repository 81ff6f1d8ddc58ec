"""Python handle over the C ABI: owns one bofi_handle_t, uploads a reference-layout state_dict,
and runs encode / decode on raw device pointers of torch tensors (torch is only the allocator and
the stream provider here)."""
import ctypes as C

import torch

from . import _lib
from .layout import BofiConfig, state_spec


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _feat_code(t):
    """BOFI_FEAT_* of a region-feature tensor (fp32 as the reference's loader yields it, or the feeder's 2-byte forms)."""
    name = str(t.dtype).replace("torch.", "")
    if name not in _lib.FEAT:
        raise TypeError("att_feats must be float32, bfloat16 or float16, got %s" % t.dtype)
    return _lib.FEAT[name]


class BofiEngine:
    def __init__(self, cfg: BofiConfig, device=0, precision="fp32"):
        if not torch.cuda.is_available():
            raise RuntimeError("boficap_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.precision = precision
        self.device = torch.device("cuda", device if isinstance(device, int) else (device.index or 0))
        c = _lib.BofiConfigC(
            abi_version=_lib.ABI_VERSION, tgt_vocab=cfg.tgt_vocab, att_feat_size=cfg.att_feat_size,
            n_enc=cfg.N_enc, n_dec=cfg.N_dec, n_len=cfg.N_len, d_model=cfg.d_model, d_ff=cfg.d_ff, heads=cfg.h,
            seq_length=cfg.seq_length, pad_idx=cfg.pad_idx, bos_idx=cfg.bos_idx, eos_idx=cfg.eos_idx,
            len_idx=cfg.len_idx, precision=_lib.PRECISION[precision])
        self.handle = C.c_void_p()
        _lib.check(self.lib.bofi_create(C.byref(c), self.device.index, C.byref(self.handle)))
        self._batch = None

    def close(self):
        if getattr(self, "handle", None) and self.handle.value:
            self.lib.bofi_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- weights ---------------------------------------------------------------------------
    def load_state_dict(self, sd):
        spec = state_spec(self.cfg)
        missing = [k for k in spec if k not in sd]
        unexpected = [k for k in sd if k not in spec]
        if missing or unexpected:
            raise RuntimeError("state_dict mismatch: missing %s unexpected %s" % (missing[:4], unexpected[:4]))
        for name, (shape, _) in spec.items():
            t = sd[name].detach().to("cpu", torch.float32).contiguous()
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError("size mismatch for %s: %s vs %s" % (name, tuple(t.shape), tuple(shape)))
            _lib.check(self.lib.bofi_set_weight(self.handle, name.encode(), C.c_void_p(t.data_ptr()), t.numel()))
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_finalize_weights(self.handle, self._stream()))
        return self

    # ---- device path -------------------------------------------------------------------------
    def encode(self, att_feats, att_len=None, want_memory=False):
        assert att_feats.is_cuda
        code = _feat_code(att_feats)
        att_feats = att_feats.contiguous()
        B, R, _ = att_feats.shape
        if att_len is not None:
            att_len = att_len.to(device=att_feats.device, dtype=torch.int32).contiguous()
        memory = torch.empty(B, R, self.cfg.d_model, device=att_feats.device) if want_memory else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_encode_ex(self.handle, self._stream(), _ptr(att_feats), code, _ptr(att_len), B, R, _ptr(memory)))
        self._batch = (B, R, att_feats, att_len)      # keep inputs alive until decode is enqueued
        return memory

    def decode(self, mode="NAIC", sample_n=1, output_logsoftmax=1, want_logprobs=True, out=None, dense_logprobs=False):
        """out: optional (seq, logp, pnum, plen, psyn) tensors of a previous call with the same shapes to write into.
        NAIC log-probs come back as `buf[:, :, :V]` of a [rows, L, V rounded up to a multiple of 4] allocation unless
        dense_logprobs=True: every row then starts on a 16-byte boundary and the library writes the tensor with TMA stores
        (bofi_decode_ex).  Same values and indexing as the reference's seq_logprob, not contiguous (`.contiguous()` if needed)."""
        B = self._batch[0]
        rows, L, V = B * sample_n, self.cfg.seq_length, self.cfg.tgt_vocab
        dev = self.device
        if out is not None and out[0].shape[0] == rows and (out[1] is not None) == bool(want_logprobs):
            seq, logp, pnum, plen, psyn = out
        else:
            seq = torch.empty(rows, L, dtype=torch.int64, device=dev)
            Vp = V if (dense_logprobs or mode != "NAIC") else (V + 3) // 4 * 4
            logp = torch.empty(rows, L, Vp, dtype=torch.float32, device=dev)[:, :, :V] if want_logprobs else None
            pnum = torch.empty(rows, dtype=torch.int32, device=dev)
            plen = torch.empty(rows, L, dtype=torch.int32, device=dev)
            psyn = torch.empty(rows, L, dtype=torch.int64, device=dev)
        ld = int(logp.stride(1)) if logp is not None else 0
        if logp is not None:
            assert logp.stride(2) == 1 and logp.stride(0) == L * ld, "log-prob buffer: [rows, L, V] rows with a common pitch"
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_decode_ex(self.handle, self._stream(), _lib.MODE[mode], sample_n, int(output_logsoftmax),
                                               _ptr(seq), _ptr(logp), ld, _ptr(pnum), _ptr(plen), _ptr(psyn)))
        return seq, logp, pnum, plen, psyn

    # ---- compact features: only the valid regions, image after image (bofi_encode_compact / bofi_stage_compact) ----------------
    def encode_compact(self, att_compact, att_len, R):
        """att_compact [sum(att_len), F] device tensor (fp32 / bf16 / fp16), att_len [B]: _prepare_feature + encoder on the valid
        regions only; same memory as encode() on the padded [B, R, F] tensor."""
        assert att_compact.is_cuda and att_compact.is_contiguous() and att_compact.dim() == 2
        att_len = att_len.to(device=att_compact.device, dtype=torch.int32).contiguous()
        B = att_len.shape[0]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_encode_compact(self.handle, self._stream(), _ptr(att_compact), _feat_code(att_compact), _ptr(att_len),
                                                    int(att_compact.shape[0]), B, int(R)))
        self._batch = (B, R, att_compact, att_len)

    def stage_compact(self, att_compact, att_len, R, row0=0, image0=0, total_images=None):
        """Asynchronous copy of a compact batch (pinned host or device) into the library's staging buffer; row0 / image0 /
        total_images: the batch is one of several that will share a call (its compact rows follow those of the batches before)."""
        assert att_compact.is_contiguous() and att_compact.dim() == 2
        att_len = att_len.to(torch.int32).contiguous()
        B = int(att_len.shape[0])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_stage_compact_part(self.handle, self._stream(), _ptr(att_compact), _feat_code(att_compact), _ptr(att_len),
                                                        int(att_compact.shape[0]), int(row0), int(image0), B,
                                                        int(total_images if total_images is not None else B), int(R)))
        gens = getattr(self, "_stage_gens", None)
        if gens is None:
            gens = self._stage_gens = []
        if image0 == 0 or not gens:
            gens.append([])
            del gens[:-3]
        gens[-1].append((att_compact, att_len))
        return _feat_code(att_compact)

    def encode_staged_compact(self, code, total_rows, B, R):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_encode_staged_compact(self.handle, self._stream(), int(code), int(total_rows), int(B), int(R)))
        self._batch = (B, R, None, None)

    # ---- several batches in one call (bofi_set_shard / bofi_stage_part / bofi_encode_staged / bofi_sample_staged) --------------
    def set_shard(self, shard_images):
        """The images of the following encode / decode calls are batches of `shard_images` images that keep their own fill
        window (0: the call is one batch): bit for bit the results of separate calls, one bounding loop for all of them."""
        _lib.check(self.lib.bofi_set_shard(self.handle, int(shard_images)))

    def stage_part(self, att_feats, att_len, row0, total):
        """Copy one batch (host pinned or device tensor) to rows [row0, row0 + B) of the library's staging buffer for `total`
        images; asynchronous on the current stream (the source must stay alive until the stream has passed the copy)."""
        assert att_feats.is_contiguous()
        code = _feat_code(att_feats)
        B, R, _ = att_feats.shape
        if att_len is not None:
            att_len = att_len.to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_stage_part(self.handle, self._stream(), _ptr(att_feats), code, _ptr(att_len), int(row0), B, int(total), R))
        # keep the sources alive while their copies may still be pending: this call's batch and the two before it on this handle
        # (a handle's calls are stream-ordered, and a caller that reuses a slot has waited for the slot's previous ticket)
        gens = getattr(self, "_stage_gens", None)
        if gens is None:
            gens = self._stage_gens = []
        if row0 == 0:
            gens.append([])
            del gens[:-3]
        if not gens:
            gens.append([])
        gens[-1].append((att_feats, att_len))
        return code

    def encode_staged(self, code, have_len, B, R):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_encode_staged(self.handle, self._stream(), int(code), int(bool(have_len)), int(B), int(R)))
        self._batch = (B, R, None, None)

    def decode_host(self, mode="NAIC", sample_n=1, output_logsoftmax=1, out=None, want_logprobs=False):
        """decode of the encoded batch, results copied to pinned host tensors (asynchronous; bofi_decode_host_async)."""
        B = self._batch[0]
        rows, L, V = B * sample_n, self.cfg.seq_length, self.cfg.tgt_vocab
        if out is not None and (out["seq"].shape[0] != rows or (out.get("logp") is not None) != bool(want_logprobs)):
            out = None
        if out is None:
            out = dict(seq=torch.empty(rows, L, dtype=torch.int64).pin_memory(),
                       logp=torch.empty(rows, L, V).pin_memory() if want_logprobs else None,
                       pnum=torch.empty(rows, dtype=torch.int32).pin_memory(),
                       plen=torch.empty(rows, L, dtype=torch.int32).pin_memory(),
                       psyn=torch.empty(rows, L, dtype=torch.int64).pin_memory())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_decode_host_async(
                self.handle, self._stream(), _lib.MODE[mode], sample_n, int(output_logsoftmax),
                _ptr(out["seq"]), _ptr(out.get("logp")), _ptr(out["pnum"]), _ptr(out["plen"]), _ptr(out["psyn"])))
        self._host_keep = out
        return out

    def sample_staged(self, code, have_len, B, R, mode="NAIC", sample_n=1, output_logsoftmax=1, out=None, want_logprobs=False):
        """encode + decode of the staged batch, results copied to pinned host tensors (asynchronous, as sample_host(sync=False))."""
        rows, L, V = B * sample_n, self.cfg.seq_length, self.cfg.tgt_vocab
        if out is not None and (out["seq"].shape[0] != rows or (out.get("logp") is not None) != bool(want_logprobs)):
            out = None
        if out is None:
            out = dict(seq=torch.empty(rows, L, dtype=torch.int64).pin_memory(),
                       logp=torch.empty(rows, L, V).pin_memory() if want_logprobs else None,
                       pnum=torch.empty(rows, dtype=torch.int32).pin_memory(),
                       plen=torch.empty(rows, L, dtype=torch.int32).pin_memory(),
                       psyn=torch.empty(rows, L, dtype=torch.int64).pin_memory())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_sample_staged(
                self.handle, self._stream(), _lib.MODE[mode], sample_n, int(output_logsoftmax), int(code), int(bool(have_len)), int(B), int(R),
                _ptr(out["seq"]), _ptr(out.get("logp")), _ptr(out["pnum"]), _ptr(out["plen"]), _ptr(out["psyn"])))
        self._host_keep = out
        return out

    def sample_host(self, att_feats, att_len=None, mode="NAIC", sample_n=1, output_logsoftmax=1, out=None, want_logprobs=False,
                    sync=True):
        """End-to-end call on HOST tensors (pinned or pageable): H2D, encode, decode, D2H, sync.
        sync=False only enqueues (bofi_sample_host_async): inputs / outputs must be pinned and stay alive until the
        current stream has drained."""
        assert not att_feats.is_cuda and att_feats.is_contiguous()
        code = _feat_code(att_feats)
        B, R, _ = att_feats.shape
        rows, L, V = B * sample_n, self.cfg.seq_length, self.cfg.tgt_vocab
        # a slot's buffers of an earlier call are only reused when they have exactly this call's shapes (the short last
        # batch of an epoch, another sample_n or want_logprobs would otherwise be written past their end)
        if out is not None and (out["seq"].shape[0] != rows or (out.get("logp") is not None) != bool(want_logprobs)
                                or (want_logprobs and out["logp"].shape[0] != rows)):
            out = None
        if out is None:
            out = dict(seq=torch.empty(rows, L, dtype=torch.int64).pin_memory(),
                       logp=torch.empty(rows, L, V).pin_memory() if want_logprobs else None,
                       pnum=torch.empty(rows, dtype=torch.int32).pin_memory(),
                       plen=torch.empty(rows, L, dtype=torch.int32).pin_memory(),
                       psyn=torch.empty(rows, L, dtype=torch.int64).pin_memory())
        if att_len is not None:
            att_len = att_len.to(torch.int32).contiguous()
        with torch.cuda.device(self.device):
            if not sync:
                assert att_feats.is_pinned() and out["seq"].is_pinned(), "asynchronous host path needs pinned buffers"
                self._host_keep = (att_feats, att_len, out)
            _lib.check(self.lib.bofi_sample_host_async_ex(
                self.handle, self._stream(), _lib.MODE[mode], sample_n, int(output_logsoftmax), _ptr(att_feats), code, _ptr(att_len),
                B, R, _ptr(out["seq"]), _ptr(out.get("logp")), _ptr(out["pnum"]), _ptr(out["plen"]), _ptr(out["psyn"])))
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
        return out

    def masks_to_len(self, att_masks):
        """att_masks [B, R] on the device -> int32 valid-region counts, with the device-side prefix-mask check
        (bofi_masks_to_len); `check_masks()` raises later, at a point where the caller synchronises anyway."""
        m = att_masks.to(torch.float32).contiguous()
        out = torch.empty(m.shape[0], dtype=torch.int32, device=m.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_masks_to_len(self.handle, self._stream(), _ptr(m), m.shape[0], m.shape[1], _ptr(out)))
        self._mask_keep = m
        return out

    def check_masks(self):
        bad = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_check_masks(self.handle, self._stream(), C.byref(bad)))
        if bad.value:
            raise ValueError("att_masks is not a prefix mask (valid regions first, as dataloader.py:333-338 builds it): this library "
                             "keeps only the per-image count of valid regions")

    def set_decode_stats(self, slot_entropy=None, slot_logp=None):
        self._stats_keep = (slot_entropy, slot_logp)
        _lib.check(self.lib.bofi_set_decode_stats(self.handle, _ptr(slot_entropy), _ptr(slot_logp)))

    def set_sampling(self, method="greedy", temperature=1.0, seed=0):
        _lib.check(self.lib.bofi_set_sampling(self.handle, 0 if method == "greedy" else 1, float(temperature), int(seed) & 0xFFFFFFFF))

    def decode_info(self):
        info = _lib.DecodeInfoC()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_get_decode_info(self.handle, self._stream(), C.byref(info)))
        return {n: getattr(info, n) for n, _ in info._fields_}

    PROFILE_CLASSES = ("gemm_tcgen05", "gemm_ffma", "attention", "layernorm", "vocab_epilogue", "other")

    def set_profiling(self, on):
        _lib.check(self.lib.bofi_set_profiling(self.handle, int(on)))

    def get_profile(self):
        n = len(self.PROFILE_CLASSES)
        launches, ms, fl, by = (C.c_int32 * n)(), (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_get_profile(self.handle, self._stream(), launches, ms, fl, by))
        return {c: dict(launches=launches[i], ms=ms[i], flops=fl[i], bytes=by[i]) for i, c in enumerate(self.PROFILE_CLASSES)}

    def get_profile_top_gemm(self):
        mnk, n, ms, fl = (C.c_int32 * 3)(), C.c_int32(), C.c_double(), C.c_double()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_get_profile_top_gemm(self.handle, self._stream(), mnk, C.byref(n), C.byref(ms), C.byref(fl)))
        return dict(M=mnk[0], N=mnk[1], K=mnk[2], launches=n.value, ms=ms.value, flops=fl.value)

    def workspace_bytes(self, B, R, sample_n=1):
        return int(self.lib.bofi_workspace_bytes(self.handle, B, R, sample_n))

    # ---- XE training (A16) ----------------------------------------------------------------------
    def param_numel(self):
        return int(self.lib.bofi_param_numel(self.handle))

    def param_layout(self):
        """name -> (offset, numel) inside the flat parameter / gradient buffers."""
        out = {}
        for name in state_spec(self.cfg):
            off, n = C.c_int64(), C.c_int64()
            _lib.check(self.lib.bofi_param_offset(self.handle, name.encode(), C.byref(off), C.byref(n)))
            out[name] = (off.value, n.value)
        return out

    def train_bind(self):
        """Allocates the flat parameter / gradient tensors, hands them to the library and returns them."""
        n = self.param_numel()
        self.flat_w = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_train_bind(self.handle, self._stream(), _ptr(self.flat_w), _ptr(self.flat_g)))
        return self.flat_w, self.flat_g

    def refresh_weights(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_refresh_weights(self.handle, self._stream()))

    @staticmethod
    def _i32(t):
        return t.to(dtype=torch.int32).contiguous()

    def _xe_inputs(self, att_feats, att_len, batch):
        """batch: dict of 2-D caption-row tensors on the device (labels, phrase_num, phrase_length, [phrase_syn],
        extend_phrase_syn_seq, extend_phrase_seq, sa_vis); returns the shapes and the int32 tensors (kept alive)."""
        assert att_feats.is_cuda and att_feats.dtype == torch.float32
        att_feats = att_feats.contiguous()
        B, R, _ = att_feats.shape
        labels = self._i32(batch["labels"])
        N, Lb = labels.shape
        assert N % B == 0, "caption rows must be a multiple of the image count"
        ints = dict(labels=labels, phrase_num=self._i32(batch["phrase_num"]), phrase_length=self._i32(batch["phrase_length"]),
                    ext_syn=self._i32(batch["extend_phrase_syn_seq"]), ext_seq=self._i32(batch["extend_phrase_seq"]),
                    sa_vis=self._i32(batch["sa_vis"]))
        if "phrase_syn" in batch and batch["phrase_syn"] is not None:
            ints["phrase_syn"] = self._i32(batch["phrase_syn"])
        if att_len is not None:
            att_len = att_len.to(device=att_feats.device, dtype=torch.int32).contiguous()
        P = int(batch["P"]) if "P" in batch else int(ints["phrase_num"].max())
        return att_feats, att_len, ints, (B, R, N // B, Lb - 2, P, N, Lb)

    def train_forward(self, att_feats, att_len, batch):
        att_feats, att_len, ints, (B, R, spi, L, P, N, Lb) = self._xe_inputs(att_feats, att_len, batch)
        V, dev = self.cfg.tgt_vocab, self.device
        outs = [torch.empty(N, Lb - 1, 20, device=dev), torch.empty(N, Lb - 1, 10, device=dev), torch.empty(N, L, V, device=dev),
                torch.empty(N, Lb - 1, 20, device=dev), torch.empty(N, Lb - 1, 10, device=dev), torch.empty(N, L, V, device=dev)]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_train_forward(
                self.handle, self._stream(), _ptr(att_feats), _ptr(att_len), B, R, spi, L, P, _ptr(ints["labels"]),
                _ptr(ints["phrase_num"]), _ptr(ints["phrase_length"]), _ptr(ints["ext_syn"]), _ptr(ints["ext_seq"]), _ptr(ints["sa_vis"]),
                *[_ptr(o) for o in outs]))
        self._train_keep = (att_feats, att_len, ints)      # fp32 mode reads att_feats again in the backward pass
        return outs

    def train_backward(self, grads, outs):
        grads = [g.contiguous().float() for g in grads]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_train_backward(self.handle, self._stream(), *[_ptr(g) for g in grads], *[_ptr(o) for o in outs]))
        self._train_keep = None

    def train_step_xe(self, att_feats, att_len, batch):
        """Fused forward + criterion + backward; returns the seven losses (device f32[7])."""
        att_feats, att_len, ints, (B, R, spi, L, P, N, Lb) = self._xe_inputs(att_feats, att_len, batch)
        losses = torch.empty(7, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_train_step_xe(
                self.handle, self._stream(), _ptr(att_feats), _ptr(att_len), B, R, spi, L, P, _ptr(ints["labels"]),
                _ptr(ints["phrase_num"]), _ptr(ints["phrase_length"]), _ptr(ints["phrase_syn"]), _ptr(ints["ext_syn"]),
                _ptr(ints["ext_seq"]), _ptr(ints["sa_vis"]), _ptr(losses)))
        self._train_keep = (att_feats, att_len, ints)
        return losses

    # ---- self-critical sampling with a tape -------------------------------------
    def sc_sample(self, att_feats, att_len=None, mode="NAIC", sample_n=1):
        """bofi_sc_sample: (seq, logprobs, phrase_num, phrase_length, phrase_syn) for B * sample_n rows; the tape of the
        encoder and of the scoring decoder pass stays in the handle until sc_backward."""
        assert att_feats.is_cuda and att_feats.dtype == torch.float32
        att_feats = att_feats.contiguous()
        B, R, _ = att_feats.shape
        rows, L, V, dev = B * sample_n, self.cfg.seq_length, self.cfg.tgt_vocab, self.device
        if att_len is not None:
            att_len = att_len.to(device=att_feats.device, dtype=torch.int32).contiguous()
        seq = torch.empty(rows, L, dtype=torch.int64, device=dev)
        logp = torch.empty(rows, L, V, dtype=torch.float32, device=dev)
        pnum = torch.empty(rows, dtype=torch.int32, device=dev)
        plen = torch.empty(rows, L, dtype=torch.int32, device=dev)
        psyn = torch.empty(rows, L, dtype=torch.int64, device=dev)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_sc_sample(self.handle, self._stream(), _lib.MODE[mode], sample_n, _ptr(att_feats), _ptr(att_len), B, R,
                                               _ptr(seq), _ptr(logp), _ptr(pnum), _ptr(plen), _ptr(psyn)))
        self._train_keep = (att_feats, att_len)
        return seq, logp, pnum, plen, psyn

    def sc_backward(self, g_logp, logp):
        g_logp = g_logp.contiguous().float()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_sc_backward(self.handle, self._stream(), _ptr(g_logp), _ptr(logp)))
        self._train_keep = None

    def sc_inputs(self, rows):
        """Decoder inputs of the taped pass of the last sc_sample: (word_ids, syn_ids, visible keys) [rows, L], total [rows]."""
        L, dev = self.cfg.seq_length, self.device
        w, s_, v = (torch.empty(rows, L, dtype=torch.int32, device=dev) for _ in range(3))
        t = torch.empty(rows, dtype=torch.int32, device=dev)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bofi_sc_inputs(self.handle, self._stream(), _ptr(w), _ptr(s_), _ptr(v), _ptr(t)))
        return w, s_, v, t

    def train_set_dropout(self, p, p_att_embed, seed):
        _lib.check(self.lib.bofi_train_set_dropout(self.handle, float(p), float(p_att_embed), int(seed) & 0xFFFFFFFF))

    def train_set_layer_event(self, layer, event):
        """event: torch.cuda.Event (recorded once) or None; recorded once encoder layer `layer`'s gradients are final."""
        keep = getattr(self, "_layer_event_keep", None)
        if keep is None:
            keep = self._layer_event_keep = {}
        keep[layer] = event
        _lib.check(self.lib.bofi_train_set_layer_event(self.handle, int(layer), C.c_void_p(event.cuda_event) if event is not None else None))

    def train_set_grad_event(self, event):
        """event: torch.cuda.Event (recorded once so that its handle exists) or None; see bofi_train_set_grad_event."""
        self._grad_event_keep = event
        _lib.check(self.lib.bofi_train_set_grad_event(self.handle, C.c_void_p(event.cuda_event) if event is not None else None))

    def train_set_glat(self, glat_p, seed=0):
        _lib.check(self.lib.bofi_train_set_glat(self.handle, float(glat_p), int(seed) & 0xFFFFFFFF))

    def train_launches(self):
        return int(self.lib.bofi_train_launches(self.handle))

    # ---- unit entry points (parity tests) ------------------------------------------------------
    def layernorm(self, x, a2, b2):
        out = torch.empty_like(x)
        _lib.check(self.lib.bofi_layernorm_f32(self.handle, self._stream(), _ptr(x), _ptr(a2), _ptr(b2), _ptr(out),
                                               x.numel() // x.shape[-1]))
        return out

    def linear(self, a, w, bias=None, residual=None, relu=False, inplace=False):
        """Unit entry of the GEMM backends.  inplace=True: `residual` is updated in place (x += a . w^T + b), the form
        the decode path uses for the O-projections and FFN2 (TMA reduce stores on the tcgen05 backends)."""
        M, K = a.shape
        N = w.shape[0]
        out = residual if inplace else torch.empty(M, N, device=a.device, dtype=torch.float32)
        _lib.check(self.lib.bofi_linear_f32(self.handle, self._stream(), _ptr(a), _ptr(w), _ptr(bias), _ptr(residual),
                                            _ptr(out), M, N, K, int(relu)))
        return out

    def linear_resid_ln(self, a, w, bias, x, ln_a, ln_b, rows_dev=None):
        """Unit entry of the residual GEMM with the following LayerNorm in its epilogue (gemm_tc2_ln.cuh; bf16 engines):
        x += a . w^T + bias in place, returns y = LayerNorm(x; ln_a, ln_b) as the kernel's bf16 values widened to fp32."""
        M, K = a.shape
        assert w.shape == (512, K) and x.shape == (M, 512)
        y = torch.empty(M, 512, device=a.device, dtype=torch.float32)
        _lib.check(self.lib.bofi_linear_resid_ln(self.handle, self._stream(), _ptr(a), _ptr(w), _ptr(bias), _ptr(x), _ptr(ln_a),
                                                 _ptr(ln_b), _ptr(y), M, K, _ptr(rows_dev)))
        return y

    def attention(self, q, k, v, vis=None):
        B, Tq, _ = q.shape
        Tk = k.shape[1]
        out = torch.empty_like(q)
        _lib.check(self.lib.bofi_attention_f32(self.handle, self._stream(), _ptr(q), _ptr(k), _ptr(v), _ptr(vis),
                                               _ptr(out), B, Tq, Tk))
        return out
