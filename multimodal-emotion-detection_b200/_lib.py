"""ctypes binding of libavhot.so (include/avhot.h).  No CPU fallback: if the library is missing it is
built with nvcc, and if that fails the import raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AVHOT_LIB_PATH") or os.path.join(HERE, "libavhot.so")   # override: A/B runs of two builds

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
PREC_FP32, PREC_TC = 0, 2
POOL_ATTENTION, POOL_AVERAGE, POOL_MAX = 0, 1, 2
SITE_FRAME_MLP, SITE_FRAME_POOL, SITE_GATE, SITE_MODMASK, SITE_ATTN = 0, 1, 2, 3, 16
SITE_EARLY0, SITE_EARLY1, SITE_LATE = 4, 5, 32
SITE_CNN0, SITE_CNN1 = 6, 7
SITE_UNC = 48
SITE_TFM, SITE_PAIR_MSG = 64, 80
TFM_MAX_LAYERS = 4
MAX_MODALITIES = 4

c_float_p = C.c_void_p  # device pointers travel as integers


class LstmDesc(C.Structure):
    _fields_ = [("input_dim", C.c_int), ("hidden_dim", C.c_int), ("num_layers", C.c_int),
                ("output_dim", C.c_int),
                ("weight_ih", C.c_void_p * 4), ("weight_hh", C.c_void_p * 4),
                ("bias_ih", C.c_void_p * 4), ("bias_hh", C.c_void_p * 4),
                ("proj_weight", C.c_void_p), ("proj_bias", C.c_void_p)]


class CnnDesc(C.Structure):
    _fields_ = [("input_dim", C.c_int), ("hidden_dim", C.c_int), ("output_dim", C.c_int)] + [(n, C.c_void_p) for n in (
        "conv1_weight", "conv1_bias", "bn1_weight", "bn1_bias", "bn1_mean", "bn1_var",
        "conv2_weight", "conv2_bias", "bn2_weight", "bn2_bias", "bn2_mean", "bn2_var",
        "proj_weight", "proj_bias")] + [("bn_eps", C.c_float)]


class TfmDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("input_dim", "hidden_dim", "output_dim", "num_layers", "num_heads", "ffn_dim",
                                       "max_len")] + [("ln_eps", C.c_float)] + \
               [(n, C.c_void_p) for n in ("input_w", "input_b", "pos", "proj_w", "proj_b")] + \
               [(n, C.c_void_p * 4) for n in ("in_proj_w", "in_proj_b", "out_w", "out_b", "lin1_w", "lin1_b", "lin2_w",
                                              "lin2_b", "n1_w", "n1_b", "n2_w", "n2_b")]


class TattnDesc(C.Structure):
    _fields_ = [("feature_dim", C.c_int), ("hidden_dim", C.c_int), ("num_heads", C.c_int)] + [(n, C.c_void_p) for n in (
        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class FrameDesc(C.Structure):
    _fields_ = [("frame_dim", C.c_int), ("hidden_dim", C.c_int), ("output_dim", C.c_int),
                ("mlp_weight", C.c_void_p), ("mlp_bias", C.c_void_p),
                ("attn_weight", C.c_void_p), ("attn_bias", C.c_void_p),
                ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p),
                ("proj_weight", C.c_void_p), ("proj_bias", C.c_void_p)]


class XattnDesc(C.Structure):
    _fields_ = [("query_dim", C.c_int), ("key_dim", C.c_int), ("hidden_dim", C.c_int),
                ("num_heads", C.c_int)] + [(n, C.c_void_p) for n in (
                    "q_ln_w", "q_ln_b", "k_ln_w", "k_ln_b", "v_ln_w", "v_ln_b",
                    "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class FusionDesc(C.Structure):
    _fields_ = [("num_modalities", C.c_int), ("hidden_dim", C.c_int), ("num_heads", C.c_int),
                ("num_classes", C.c_int), ("gate_hidden", C.c_int),
                ("modality_dim", C.c_int * MAX_MODALITIES),
                ("proj_w", C.c_void_p * MAX_MODALITIES), ("proj_b", C.c_void_p * MAX_MODALITIES),
                ("attn", XattnDesc * MAX_MODALITIES)] + [(n, C.c_void_p) for n in (
                    "pre_ln_w", "pre_ln_b", "post_ln_w", "post_ln_b",
                    "gate0_w", "gate0_b", "gate3_w", "gate3_b", "cls_w", "cls_b")]


class EarlyDesc(C.Structure):
    _fields_ = [("num_modalities", C.c_int), ("hidden_dim", C.c_int), ("num_classes", C.c_int),
                ("modality_dim", C.c_int * MAX_MODALITIES)] + [(n, C.c_void_p) for n in (
                    "w0", "b0", "ln1_w", "ln1_b", "w4", "b4", "ln5_w", "ln5_b", "w8", "b8")]


class LateDesc(C.Structure):
    _fields_ = [("num_modalities", C.c_int), ("hidden_dim", C.c_int), ("num_classes", C.c_int),
                ("modality_dim", C.c_int * MAX_MODALITIES),
                ("w0", C.c_void_p * MAX_MODALITIES), ("b0", C.c_void_p * MAX_MODALITIES),
                ("w3", C.c_void_p * MAX_MODALITIES), ("b3", C.c_void_p * MAX_MODALITIES),
                ("fusion_logits", C.c_void_p)]


class UncDesc(C.Structure):
    _fields_ = [("num_modalities", C.c_int), ("hidden_dim", C.c_int), ("num_classes", C.c_int),
                ("modality_dim", C.c_int * MAX_MODALITIES)] + [(n, C.c_void_p * MAX_MODALITIES) for n in (
                    "cls_w0", "cls_b0", "cls_w3", "cls_b3", "unc_w0", "unc_b0", "unc_w3", "unc_b3")] + \
               [("epsilon", C.c_float)]


class HeadDesc(C.Structure):
    _fields_ = [("in_dim", C.c_int), ("hidden_dim", C.c_int), ("num_classes", C.c_int),
                ("w0", C.c_void_p), ("b0", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p)]


_P, _I, _U32, _U64, _F, _SZ = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/avhot.h declares
PROTOTYPES = {
    "avhot_version": (_I, []),
    "avhot_last_error": (C.c_char_p, []),
    "avhot_check_device": (_I, []),
    "avhot_launch_count": (C.c_longlong, []),
    "avhot_prof_enable": (_I, [_I]),
    "avhot_prof_reset": (_I, []),
    "avhot_prof_read": (_I, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(_I)]),
    "avhot_prof_names": (_I, [C.c_char_p, _I]),
    "avhot_gemm": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "avhot_dropout_keep": (_I, [_P, _I, _I, _U64, _U32, _U32, _F, _U64, _P]),
    "avhot_modality_mask": (_I, [_P, _I, _I, _U64, _F, _U64, _P]),
    "avhot_synth_audio": (_I, [_P, _I, _I, _U64, _U64, _P]),
    "avhot_synth_video": (_I, [_P, _P, _I, _I, _U64, _U64, _P]),
    "avhot_condition_audio": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "avhot_condition_audio_i16": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "avhot_frontend": (_I, [_P, _P, _I, _I, _P]),
    "avhot_frontend_ex": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "avhot_frontend_tables_host": (_I, [_P, _P, _P]),
    "avhot_lstm_create": (_I, [C.POINTER(LstmDesc), C.POINTER(_P), _P]),
    "avhot_lstm_destroy": (None, [_P]),
    "avhot_lstm_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "avhot_lstm_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _SZ, _P]),
    "avhot_cnn_create": (_I, [C.POINTER(CnnDesc), C.POINTER(_P), _P]),
    "avhot_cnn_destroy": (None, [_P]),
    "avhot_cnn_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "avhot_cnn_forward": (_I, [_P, _P, _P, _I, _I, _I, _F, _U64, _U64, _U32, _P, _SZ, _P]),
    "avhot_gru_create": (_I, [C.POINTER(LstmDesc), C.POINTER(_P), _P]),
    "avhot_gru_destroy": (None, [_P]),
    "avhot_gru_workspace_bytes": (_SZ, [_P, _I, _I]),
    "avhot_gru_forward": (_I, [_P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "avhot_lstm_split_pitch": (_I, [_P]),
    "avhot_lstm_forward_split": (_I, [_P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "avhot_lstm_forward_frontend": (_I, [_P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "avhot_frame_create": (_I, [C.POINTER(FrameDesc), C.POINTER(_P), _P]),
    "avhot_frame_destroy": (None, [_P]),
    "avhot_frame_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "avhot_frame_hidden": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _SZ, _P]),
    "avhot_frame_tail": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _U64, _U64, _U32, _P]),
    "avhot_frame_tail_norm": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _U64, _U64, _U32, _P]),
    "avhot_frame_tail_pool": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _U64, _U64, _U32, _I, _P]),
    "avhot_frame_attention_pool": (_I, [_P, _P, _P, _P, _I, _I, _P]),
    "avhot_frame_tail_ids": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _U64, _P, _U32, _P]),
    "avhot_fusion_forward_ids": (_I, [_P, C.POINTER(_P), C.POINTER(_I), _P, _P, _I, _I, _F, _U64, _P, _U32, _I, _P, _SZ,
                                      _P]),
    "avhot_flag_near_ties": (_I, [_P, _I, _I, _F, _U32, _U64, _P, _P, _P, _P, _I, _P]),
    "avhot_gather_flagged": (_I, [_P, _SZ, _I, _P, _P, _P, _I, _P]),
    "avhot_scatter_resolved": (_I, [_P, _P, _I, _U32, _P, _P, _P, _P, _P, _P, _I, _P]),
    "avhot_xattn_create": (_I, [C.POINTER(XattnDesc), C.POINTER(_P), _P]),
    "avhot_xattn_destroy": (None, [_P]),
    "avhot_xattn_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "avhot_xattn_forward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _U64, _U64, _U32, _U32,
                                 _P, _SZ, _P]),
    "avhot_fusion_create": (_I, [C.POINTER(FusionDesc), C.POINTER(HeadDesc), C.POINTER(_P), _P]),
    "avhot_fusion_destroy": (None, [_P]),
    "avhot_fusion_workspace_bytes": (_SZ, [_P, _I, _I]),
    "avhot_fusion_forward": (_I, [_P, C.POINTER(_P), C.POINTER(_I), _P, _P, _P, _P, _P, _I, _I, _F,
                                  _U64, _U64, _U32, _I, _P, _SZ, _P]),
    "avhot_early_create": (_I, [C.POINTER(EarlyDesc), C.POINTER(_P), _P]),
    "avhot_late_create": (_I, [C.POINTER(LateDesc), C.POINTER(_P), _P]),
    "avhot_head_destroy": (None, [_P]),
    "avhot_head_forward": (_I, [_P, C.POINTER(_P), C.POINTER(_I), _I, _P, _P, _P, _I, _I, _F, _U64, _U64, _U32, _P]),
    "avhot_unc_create": (_I, [C.POINTER(UncDesc), C.POINTER(_P), _P]),
    "avhot_unc_forward": (_I, [_P, C.POINTER(_P), C.POINTER(_I), _I, _P, _P, _P, _P, _I, _I, _F, _U64, _U64, _U32, _P]),
    "avhot_uncertainty_fuse": (_I, [_P, _P, _P, _P, _P, C.c_longlong, _I, _I, _F, _P]),
    "avhot_calibration_accumulate": (_I, [_P, _P, _P, _P, C.c_longlong, _I, _P, _I, _P, _P]),
    "avhot_mc_reduce": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "avhot_tfm_create": (_I, [C.POINTER(TfmDesc), C.POINTER(_P), _P]),
    "avhot_tfm_destroy": (None, [_P]),
    "avhot_tfm_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "avhot_tfm_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _U64, _U64, _U32, _I, _P, _SZ, _P]),
    "avhot_tattn_create": (_I, [C.POINTER(TattnDesc), C.POINTER(_P), _P]),
    "avhot_tattn_destroy": (None, [_P]),
    "avhot_tattn_workspace_bytes": (_SZ, [_P, _I, _I]),
    "avhot_tattn_forward": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _SZ, _P]),
    "avhot_tattn_pool": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "avhot_pair_combine": (_I, [_P, C.POINTER(_P), _I, _P, _P, _P, _F, _P, _I, _I, _F, _U64, _U64, _U32, _U32, _P]),
}

_lock = threading.Lock()
_lib = None


def load():
    """Load (building first if needed) libavhot.so.  Raises if it cannot be had - there is no fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.environ.get("AVHOT_LIB_PATH"):
            # stamp-checked: a library built from other sources than the ones on disk (a stale binary would be loaded
            # with mismatching ctypes struct layouts) is rebuilt; an up-to-date one costs a hash of csrc/
            from . import build as _build      # nvcc cross-compiles without a GPU
            if os.path.exists(_build.NVCC) or not os.path.exists(LIB_PATH):
                _build.build()
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)            # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    return load().avhot_last_error().decode(errors="replace")


def check(rc: int, what: str = "") -> None:
    """Translate an avhot_status: ValueError for bad arguments / unsupported configurations (the reference
    raises ValueError / assert there), RuntimeError for CUDA failures."""
    if rc == OK:
        return
    msg = f"{what}: {last_error()}" if what else last_error()
    if rc in (ERR_INVALID, ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(f"{msg} (avhot_status {rc})")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """The product path runs on an sm_100 device only."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("libavhot runs on CUDA tensors only (B200, sm_100a); there is no CPU fallback - "
                               f"got a tensor on {t.device}")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} and {t.device}")
    return dev


_checked_devices = set()


def check_device(dev: torch.device) -> None:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx in _checked_devices:
        return
    with torch.cuda.device(idx):
        check(load().avhot_check_device(), "avhot_check_device")
    _checked_devices.add(idx)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(dev=None):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def f32c(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous (the reference's tensors are fp32; anything else is converted like `.float()`)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def prof_enable(on: bool = True) -> None:
    load().avhot_prof_enable(1 if on else 0)


def prof_reset() -> None:
    load().avhot_prof_reset()


def prof_report() -> dict:
    """{kernel name: (total ms, launches)} since the last reset (synchronises)."""
    lib = load()
    buf = C.create_string_buffer(4096)
    check(lib.avhot_prof_names(buf, 4096))
    out = {}
    for name in [n for n in buf.value.decode().split("\n") if n]:
        ms, n = C.c_double(), C.c_int()
        check(lib.avhot_prof_read(name.encode(), C.byref(ms), C.byref(n)))
        out[name] = (ms.value, n.value)
    return out


def launch_count() -> int:
    return int(load().avhot_launch_count())
