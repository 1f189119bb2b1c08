"""ctypes binding of libnnue_b200.so (the C ABI declared in include/nnue_b200.h).

There is NO fallback: if the shared library is missing or does not load, every hot-path call
raises.  Build it with `python nnue-vision_b200/build.py` (or `__graft_entry__.build()`).
"""
import ctypes
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libnnue_b200.so"

i32, i64, f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
vp, sz = ctypes.c_void_p, ctypes.c_size_t


ABI_VERSION = 4  # NNUE_B200_ABI_VERSION of include/nnue_b200.h


class AllreduceSlice(ctypes.Structure):
    """struct nnue_allreduce_slice (include/nnue_b200.h)."""
    _fields_ = [("peer_recv_h", ctypes.c_void_p), ("peer_flags_h", ctypes.c_void_p), ("state_d", ctypes.c_void_p),
                ("n", ctypes.c_size_t), ("buf_d", ctypes.c_void_p), ("phase", ctypes.c_int)]


class NnueShape(ctypes.Structure):
    """struct nnue_shape (include/nnue_b200.h)."""
    _fields_ = [(n, i32) for n in (
        "B", "H", "W", "C", "G", "L1", "L2", "L3", "NC", "stride",
        "F", "Gh", "Gw", "P", "CW", "NW", "PP", "BW")]


SHAPE_P = ctypes.POINTER(NnueShape)

# name -> (restype, argtypes); tests/test_abi.py checks this table against the header
SIGNATURES = {
    "nnue_b200_abi_version": (ctypes.c_int, []),
    "nnue_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "nnue_last_cuda_error": (ctypes.c_char_p, []),
    "nnue_launch_count": (ctypes.c_ulonglong, [ctypes.c_int]),
    "nnue_launch_count_add": (ctypes.c_ulonglong, [ctypes.c_ulonglong]),
    "nnue_allreduce_max_world": (ctypes.c_int, []),
    "nnue_allreduce_recv_floats": (sz, [ctypes.c_int, sz]),
    "nnue_allreduce_ll_max_floats": (sz, []),
    "nnue_allreduce_oneshot": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "nnue_allreduce_oneshot_slices": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(AllreduceSlice), ctypes.c_int,
                                                      ctypes.c_void_p]),
    "nnue_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "nnue_shape_init": (ctypes.c_int, [SHAPE_P] + [ctypes.c_int] * 10),
    "nnue_workspace_bytes": (sz, [SHAPE_P]),
    "nnue_extract_fwd": (ctypes.c_int, [SHAPE_P] + [vp] * 9),
    "nnue_sparse_from_bits": (ctypes.c_int, [SHAPE_P, vp, ctypes.c_int, vp, vp, vp]),
    "nnue_ft_fwd": (ctypes.c_int, [SHAPE_P, vp, vp, vp, vp, vp, sz, vp]),
    "nnue_ft_fwd_indexed": (ctypes.c_int, [ctypes.c_int] * 4 + [vp] * 6),
    "nnue_ft_sort_pairs_workspace_bytes": (sz, [ctypes.c_int] * 3),
    "nnue_ft_sort_pairs": (ctypes.c_int, [ctypes.c_int] * 3 + [vp] * 6 + [sz, vp]),
    "nnue_ft_bwd_indexed": (ctypes.c_int, [ctypes.c_int] * 4 + [vp, vp, vp, ctypes.c_int] + [vp] * 7),
    "nnue_head_fwd": (ctypes.c_int, [SHAPE_P] + [vp] * 11),
    "nnue_ce_fwd_bwd": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, vp, vp, f32, vp, vp, vp, vp, vp, sz, vp]),
    "nnue_head_bwd": (ctypes.c_int, [SHAPE_P] + [vp] * 15 + [sz, vp]),
    "nnue_head_is_fused": (ctypes.c_int, [SHAPE_P]),
    "nnue_head_uses_umma": (ctypes.c_int, [SHAPE_P]),
    "nnue_head_train": (ctypes.c_int, [SHAPE_P, vp, vp, f32] + [vp] * 14 + [vp, sz, vp]),
    "nnue_head_train_overlapped": (ctypes.c_int, [SHAPE_P, vp, vp, f32] + [vp] * 14 + [vp, sz, vp, vp, sz, vp]),
    "nnue_head_side_workspace_bytes": (sz, [SHAPE_P]),
    "nnue_ft_bwd_is_fused": (ctypes.c_int, [SHAPE_P]),
    "nnue_ft_bwd": (ctypes.c_int, [SHAPE_P] + [vp] * 7 + [sz, vp]),
    "nnue_wants_transposed_bits": (ctypes.c_int, [SHAPE_P]),
    "nnue_ft_bwd_dw": (ctypes.c_int, [SHAPE_P, vp, vp, vp, vp, vp, vp, sz, vp]),
    "nnue_input_bwd_is_dense": (ctypes.c_int, [SHAPE_P]),
    "nnue_ft_bwd_gbin": (ctypes.c_int, [SHAPE_P, vp, vp, vp, vp, vp, sz, vp]),
    "nnue_ft_uses_mma": (ctypes.c_int, [SHAPE_P]),
    "nnue_ft_tables_bytes": (sz, [SHAPE_P]),
    "nnue_ft_format_tables": (ctypes.c_int, [SHAPE_P, vp, vp, vp]),
    "nnue_ft_fwd_tables": (ctypes.c_int, [SHAPE_P, vp, vp, vp, vp, vp]),
    "nnue_ft_bwd_gbin_tables": (ctypes.c_int, [SHAPE_P, vp, vp, vp, vp, vp, sz, vp]),
    "nnue_ft_uses_umma": (ctypes.c_int, [SHAPE_P]),
    "nnue_conv_bwd": (ctypes.c_int, [SHAPE_P] + [vp] * 8 + [sz, vp]),
    "nnue_input_bwd": (ctypes.c_int, [SHAPE_P] + [vp] * 9 + [sz, vp]),
    "nnue_input_bwd_wants_activations": (ctypes.c_int, [SHAPE_P]),
    "nnue_input_bwd_fused_ok": (ctypes.c_int, [SHAPE_P]),
    "nnue_input_bwd_fused": (ctypes.c_int, [SHAPE_P] + [vp] * 9 + [sz, vp]),
    "nnue_input_bwd_stored": (ctypes.c_int, [SHAPE_P] + [vp] * 10 + [sz, vp]),
    "nnue_ft_bwd_dval": (ctypes.c_int, [SHAPE_P] + [vp] * 8 + [sz, vp]),
    "nnue_extract_bwd": (ctypes.c_int, [SHAPE_P] + [vp] * 5 + [sz, vp]),
    "nnue_scale_flat": (ctypes.c_int, [ctypes.c_longlong, vp, vp, vp, vp]),
    "nnue_opt_workspace_bytes": (sz, [ctypes.c_longlong]),
    "nnue_opt_grad_sqnorm": (ctypes.c_int, [ctypes.c_longlong, vp, vp, vp, sz, vp]),
    "nnue_opt_sgd_step": (ctypes.c_int, [ctypes.c_longlong, vp, vp, vp, f32, f32, f32, f32, vp, ctypes.c_int, vp]),
    "nnue_opt_adam_step": (ctypes.c_int, [ctypes.c_longlong, vp, vp, vp, vp, f32, f32, f32, f32, f32, ctypes.c_int, f32,
                                          vp, vp]),
    "nnue_q_load": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(vp)]),
    "nnue_q_load_memory": (ctypes.c_int, [vp, sz, ctypes.POINTER(vp)]),
    "nnue_q_free": (None, [vp]),
    "nnue_q_dims": (ctypes.c_int, [vp, vp, ctypes.POINTER(f32)]),
    "nnue_q_infer": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp]),
    "nnue_q_conv_bound": (ctypes.c_int, [f32, ctypes.c_int, vp]),
    "nnue_q_workspace_bytes": (sz, [vp, ctypes.c_int]),
    "nnue_q_infer_ws": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp, vp, sz, vp]),
    "nnue_q_acc_apply": (ctypes.c_int, [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, vp]),
    "nnue_q_acc_score": (ctypes.c_int, [vp, ctypes.c_int, vp, ctypes.c_int, vp, vp]),
    "nnue_q_infer_host": (ctypes.c_int, [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, vp]),
}


class NnueError(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded library; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NnueError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python nnue-vision_b200/build.py`). There is no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.nnue_b200_abi_version() != ABI_VERSION:
            raise NnueError("libnnue_b200.so ABI version mismatch; rebuild it")
        _lib = handle
        # NNUE_OPTIONS="key=value,key=value": nnue_set_option knobs applied at load (to run a whole test file or bench
        # under another kernel choice without touching the code)
        import os
        for kv in filter(None, os.environ.get("NNUE_OPTIONS", "").split(",")):
            key, _, val = kv.partition("=")
            if handle.nnue_set_option(key.strip().encode(), int(val)) != 0:
                raise NnueError(f"NNUE_OPTIONS: unknown option {key!r}")
    return _lib


def check(rc):
    if rc != 0:
        L = lib()
        msg = L.nnue_error_string(rc).decode()
        cuda = L.nnue_last_cuda_error().decode()
        raise NnueError(f"libnnue_b200: {msg} (code {rc})" + (f": {cuda}" if rc == -3 and cuda else ""))


def set_option(key, value):
    global _options_epoch
    check(lib().nnue_set_option(key.encode(), int(value)))
    _options_epoch += 1


def make_shape(B, H, W, C, G, L1, L2, L3, NC, stride):
    s = NnueShape()
    check(lib().nnue_shape_init(ctypes.byref(s), B, H, W, C, G, L1, L2, L3, NC, stride))
    return s


def workspace_bytes(shape):
    return int(lib().nnue_workspace_bytes(ctypes.byref(shape)))


def dptr(t, dtype=None):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NnueError("the NNUE hot path runs on CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise NnueError("expected a contiguous tensor")
    if dtype is not None and t.dtype != dtype:
        raise NnueError(f"expected dtype {dtype}, got {t.dtype}")
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device_of(t):
    """Context manager: make the device of CUDA tensor `t` the current one for the calls inside (the C side launches
    on the CURRENT device's current stream and never calls cudaSetDevice, as PyTorch's own ops guard by tensor device:
    a model moved to cuda:1 must work while cuda:0 is current).  CPU tensors get a null context; they are rejected by
    the hot-path calls themselves."""
    if t is not None and getattr(t, "is_cuda", False):
        return torch.cuda.device(t.device)
    import contextlib
    return contextlib.nullcontext()


_options_epoch = 0


def options_epoch():
    """Bumped by every set_option(): cached launch plans / captured graphs made under other options are stale."""
    return _options_epoch
