"""ctypes binding of libbofi_b200.so (include/bofi_b200.h).  There is no fallback: if the shared
library is missing or does not export the ABI declared in the header, importing raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BOFI_LIB_PATH: load another build of the same ABI (the stall-counter build of tools/gemm_stalls.py)
LIB_PATH = os.environ.get("BOFI_LIB_PATH") or os.path.join(_HERE, "lib", "libbofi_b200.so")

ABI_VERSION = 2
OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM = 0, 1, 2, 3, 4
PRECISION = {"fp32": 0, "bf16": 1}
MODE = {"NAIC": 0, "SAIC": 1}
FEAT = {"float32": 0, "bfloat16": 1, "float16": 2}


class BofiConfigC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "tgt_vocab", "att_feat_size", "n_enc", "n_dec", "n_len", "d_model", "d_ff", "heads",
        "seq_length", "pad_idx", "bos_idx", "eos_idx", "len_idx", "precision")]


class DecodeInfoC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("bounding_steps", "fill_width", "kernel_launches", "nan_batch")]


class BofiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbofi_b200 error %d: %s" % (code, msg))
        self.code = code


_P = C.c_void_p
_I = C.c_int32
_SIGNATURES = {
    "bofi_last_error": (C.c_char_p, []),
    "bofi_abi_version": (C.c_int, []),
    "bofi_create": (C.c_int, [C.POINTER(BofiConfigC), C.c_int, C.POINTER(_P)]),
    "bofi_destroy": (C.c_int, [_P]),
    "bofi_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int64]),
    "bofi_missing_weights": (C.c_int, [_P]),
    "bofi_finalize_weights": (C.c_int, [_P, _P]),
    "bofi_workspace_bytes": (C.c_int64, [_P, _I, _I, _I]),
    "bofi_encode": (C.c_int, [_P, _P, _P, _P, _I, _I, _P]),
    "bofi_encode_ex": (C.c_int, [_P, _P, _P, _I, _P, _I, _I, _P]),
    "bofi_masks_to_len": (C.c_int, [_P, _P, _P, _I, _I, _P]),
    "bofi_check_masks": (C.c_int, [_P, _P, C.POINTER(C.c_int32)]),
    "bofi_sample_host_async_ex": (C.c_int, [_P, _P, _I, _I, _I, _P, _I, _P, _I, _I, _P, _P, _P, _P, _P]),
    "bofi_decode": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "bofi_decode_ex": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, C.c_int64, _P, _P, _P]),
    "bofi_encode_compact": (C.c_int, [_P, _P, _P, _I, _P, _I, _I, _I]),
    "bofi_stage_compact": (C.c_int, [_P, _P, _P, _I, _P, _I, _I, _I]),
    "bofi_stage_compact_part": (C.c_int, [_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I]),
    "bofi_encode_staged_compact": (C.c_int, [_P, _P, _I, _I, _I, _I]),
    "bofi_set_shard": (C.c_int, [_P, _I]),
    "bofi_stage_part": (C.c_int, [_P, _P, _P, _I, _P, _I, _I, _I, _I]),
    "bofi_encode_staged": (C.c_int, [_P, _P, _I, _I, _I, _I]),
    "bofi_sample_staged": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "bofi_decode_host_async": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "bofi_sample_host": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "bofi_sample_host_async": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "bofi_set_decode_stats": (C.c_int, [_P, _P, _P]),
    "bofi_set_sampling": (C.c_int, [_P, _I, C.c_float, C.c_uint32]),
    "bofi_get_decode_info": (C.c_int, [_P, _P, C.POINTER(DecodeInfoC)]),
    "bofi_set_profiling": (C.c_int, [_P, _I]),
    "bofi_get_profile": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "bofi_get_profile_top_gemm": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "bofi_param_numel": (C.c_int64, [_P]),
    "bofi_param_offset": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bofi_train_bind": (C.c_int, [_P, _P, _P, _P]),
    "bofi_refresh_weights": (C.c_int, [_P, _P]),
    "bofi_train_forward": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I] + [_P] * 12),
    "bofi_train_backward": (C.c_int, [_P, _P] + [_P] * 12),
    "bofi_train_step_xe": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I] + [_P] * 8),
    "bofi_train_set_dropout": (C.c_int, [_P, C.c_float, C.c_float, C.c_uint32]),
    "bofi_train_launches": (C.c_int, [_P]),
    "bofi_sc_sample": (C.c_int, [_P, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P]),
    "bofi_sc_backward": (C.c_int, [_P, _P, _P, _P]),
    "bofi_sc_inputs": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "bofi_train_set_grad_event": (C.c_int, [_P, _P]),
    "bofi_train_set_layer_event": (C.c_int, [_P, _I, _P]),
    "bofi_train_set_glat": (C.c_int, [_P, C.c_float, C.c_uint32]),
    "bofi_layernorm_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _I]),
    "bofi_linear_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I]),
    "bofi_linear_resid_ln": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "bofi_attention_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m boficap_b200.build` "
                          "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export the header's ABI
        fn.restype = res
        fn.argtypes = args
    if lib.bofi_abi_version() != ABI_VERSION:
        raise ImportError("libbofi_b200.so ABI %d != binding ABI %d" % (lib.bofi_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code):
    if code != OK:
        raise BofiError(code, load().bofi_last_error().decode("utf-8", "replace"))
