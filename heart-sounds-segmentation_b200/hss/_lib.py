"""ctypes binding of ``libhssb.so`` (C-ABI declared in ``include/hssb.h``).

The library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc -shared``).  Loading
fails loudly when it is missing: the product has no other implementation to fall back to.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HSSB_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libhssb.so"))

MODE_RAW, MODE_ABS, MODE_STACK = 0, 1, 2


class ModelParams(ctypes.Structure):
    """``hssb_model_params`` of include/hssb.h."""

    _fields_ = [
        ("input_size", c_int),
        ("hidden_size", c_int),
        ("w_ih", (c_void_p * 2) * 2),
        ("w_hh", (c_void_p * 2) * 2),
        ("b_ih", (c_void_p * 2) * 2),
        ("b_hh", (c_void_p * 2) * 2),
        ("lin_w", c_void_p),
        ("lin_b", c_void_p),
    ]


_SIGNATURES = {
    "hssb_version": (c_int, []),
    "hssb_last_error": (c_char_p, []),
    "hssb_fsst_stft": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "hssb_fsst_stats_words": (c_size_t, [c_int64, c_int64]),
    "hssb_fsst_reassign": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hssb_fsst_finish": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "hssb_fsst_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int, c_int]),
    "hssb_fsst_forward": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hssb_fsst_host": (c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "hssb_model_create": (c_int, [POINTER(ModelParams), POINTER(c_void_p), c_void_p]),
    "hssb_model_destroy": (None, [c_void_p]),
    "hssb_model_split_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "hssb_model_split_input": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_void_p]),
    "hssb_model_forward_split": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hssb_model_side_gate": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "hssb_model_update": (c_int, [c_void_p, POINTER(ModelParams), c_void_p]),
    "hssb_model_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "hssb_model_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "hssb_lstm_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "hssb_lstm_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hssb_debug_inproj": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hssb_debug_trace": (c_int, [c_void_p, c_int]),
    "hssb_debug_max_clusters": (c_int, []),
    "hssb_debug_sync_offset": (ctypes.c_longlong, [ctypes.c_longlong, ctypes.c_longlong]),
    "hssb_lstm_train_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hssb_lstm_train_forward_tc": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hssb_model_uses_tensor_cores": (c_int, [c_void_p]),
    "hssb_lstm_train_backward_tc_workspace_bytes": (c_size_t, []),
    "hssb_lstm_train_backward_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "hssb_split_tf32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "hssb_lstm_train_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "hssb_ce_head_forward": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hssb_ce_head_backward": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hssb_clip_adam_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_float, c_int64, c_float,
                                    c_void_p, c_void_p, c_void_p]),
    "hssb_confusion": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "hssb_metrics_update": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "hssb_auroc_hist": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "hssb_csv_scan": (c_int, [POINTER(c_char_p), c_int, c_int, c_void_p]),
    "hssb_csv_parse": (c_int, [POINTER(c_char_p), c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hssb_prof_enable": (c_int, [c_int]),
    "hssb_prof_read": (c_int, [c_char_p, c_size_t]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib() -> ctypes.CDLL:
    """Load ``libhssb.so`` once and attach the prototypes."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libhssb.so not found at {LIB_PATH}: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback for the FSST/BiLSTM path."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    """Map the C-ABI error convention to Python exceptions (<0 argument errors, >0 cudaError_t)."""
    if rc == 0:
        return
    msg = lib().hssb_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (hssb error {rc})")
    raise RuntimeError(f"{what}: {msg} (cudaError {rc})")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("hss (B200 build) needs a CUDA device: the FSST/BiLSTM path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


def cached_workspace(cache: dict, dev, need: int):
    """Grow-only scratch tensor for ``(device, current stream)``.

    Keyed by stream because the kernels of two concurrent calls on different CUDA streams (or threads) must not share
    scratch; an outgrown buffer is handed back to the caching allocator with ``record_stream`` so that it is not reused
    while work queued on this stream may still touch it."""
    import torch

    stream = torch.cuda.current_stream(dev)
    key = (dev.index, stream.cuda_stream)
    ws = cache.get(key)
    if ws is None or ws.numel() < need:
        if ws is not None:
            ws.record_stream(stream)
        cache.pop(key, None)
        ws = None
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        cache[key] = ws
    return ws


def prof_enable(on: bool) -> None:
    lib().hssb_prof_enable(1 if on else 0)


def prof_read() -> dict:
    """{kernel name: (launches, total_ms)} since the last read."""
    buf = ctypes.create_string_buffer(1 << 16)
    lib().hssb_prof_read(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
