"""ctypes binding of ``libnotorch_b200.so`` (the C ABI declared in ``include/notorch_b200.h``).

There is no CPU fallback and no alternative backend: if the shared library is missing and cannot
be built, importing this module raises; if a kernel reports an error, the wrapper raises
``RuntimeError`` with ``nt_last_error_string()``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnotorch_b200.so")

ABI_VERSION = 201  # = NT_ABI_VERSION of include/notorch_b200.h; a library reporting another value is refused

# enums of include/notorch_b200.h
NT_F32, NT_BF16 = 0, 1
GEMM_TF32X3, GEMM_FP32, GEMM_TF32, GEMM_BF16 = 0, 1, 2, 3
ACT_IDENTITY, ACT_RELU, ACT_LEAKY_RELU, ACT_ELU, ACT_SILU, ACT_GELU, ACT_TANH = range(7)

_i32p, _i64p, _vp = C.c_void_p, C.c_void_p, C.c_void_p  # raw device addresses
_i64, _int, _f32, _u64, _sz = C.c_int64, C.c_int, C.c_float, C.c_uint64, C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/notorch_b200.h
SIGNATURES: dict[str, tuple[object, list[object]]] = {
    "nt_last_error_string": (C.c_char_p, []),
    "nt_version": (_int, []),
    "nt_kernel_launch_count": (C.c_longlong, []),
    "nt_debug_set_trace_buffer": (None, [_vp]),
    "nt_debug_wgrad_geometry": (_int, [_i64, _i64, _int, _i64p]),
    "nt_debug_split_item": (_int, [_i64, _i64, _i64, _i64p]),
    "nt_device_supported": (_int, []),
    "nt_collate_workspace_bytes": (_sz, [_i64]),
    "nt_collate": (_int, [_i32p, _i32p, _i64, _i32p, _i32p, _i64, _i64, _int, _i64p, _i64p, _i64p, _i64p, _i32p, _i32p, _vp, _sz, _vp]),
    "nt_build_csr_workspace_bytes": (_sz, [_i64, _i64]),
    "nt_build_csr": (_int, [_i64p, _i64, _i64, _i32p, _i32p, _i32p, _i32p, _vp, _sz, _vp]),
    "nt_seg_reduce": (_int, [_vp, _i64, _i32p, _i32p, _i64, _int, _f32, _int, _f32, _vp, _int, _vp]),
    "nt_seg_reduce_ex": (_int, [_vp, _i64, _i32p, _i32p, _i64, _int, _f32, _int, _f32, _vp, _vp, _vp, _int, _vp]),
    "nt_csr_to_ell": (_int, [_i32p, _i32p, _i64, _i32p, _vp]),
    "nt_seg_reduce_ell": (_int, [_vp, _i64, _i32p, _i32p, _i32p, _i64, _int, _f32, _int, _f32, _vp, _vp, _vp, _int, _vp]),
    "nt_dense_forward": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _f32, _u64, _u64, _vp, _int, _int, _vp]),
    "nt_gather_add": (_int, [_vp, _vp, _i32p, _i32p, _i64, _i64, _f32, _vp, _int, _vp]),
    "nt_weight_image_bytes": (_sz, [_i64]),
    "nt_weight_prepare": (_int, [_vp, _i64, _int, _vp, _int, _vp]),
    "nt_layer_forward": (_int, [_vp, _vp, _i32p, _i32p, _vp, _vp, _vp, _i64, _i64, _i64, _int, _f32, _int, _f32, _u64, _u64, _vp, _vp, _int, _int, _vp]),
    "nt_layer_backward_dgrad": (_int, [_vp, _vp, _vp, _i64, _i64, _f32, _u64, _u64, _vp, _int, _int, _vp]),
    "nt_layer_backward_wgrad_workspace_bytes": (_sz, [_i64, _i64]),
    "nt_layer_backward_wgrad": (_int, [_vp, _vp, _vp, _vp, _i32p, _i32p, _i64, _i64, _i64, _int, _f32, _f32, _u64, _u64, _vp, _vp, _vp, _sz, _int, _int, _vp]),
    "nt_layer_backward_epilogue": (_int, [_vp, _vp, _vp, _vp, _i32p, _i32p, _i32p, _i32p, _i64, _i64, _int, _f32, _int, _int, _vp, _int, _vp]),
    "nt_layer_backward_epilogue_arg": (_int, [_vp, _vp, _vp, _vp, _i32p, _i32p, _i32p, _i32p, _i64, _i64, _int, _f32, _int, _vp, _int, _vp]),
    "nt_layer_backward_epilogue_fused": (_int, [_vp, _vp, _vp, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i64, _i64, _int, _f32, _int, _int, _vp, _int, _vp]),
    "nt_weighted_colsum": (_int, [_vp, _i32p, _i64, _i64, _vp, _int, _vp]),
    "nt_pooled_message_sum_workspace_bytes": (_sz, [_i64]),
    "nt_pooled_message_sum": (_int, [_vp, _i32p, _i32p, _i32p, _i32p, _i32p, _vp, _i64, _i64, _i64, _int, _f32, _int, _int, _vp, _vp, _vp, _sz, _int, _vp]),
    "nt_layer_backward_epilogue_pooled_workspace_bytes": (_sz, [_i64]),
    "nt_layer_backward_epilogue_pooled": (_int, [_vp, _vp, _vp, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i64, _i64, _i64, _int, _f32, _int, _int, _vp,
                                                 _vp, _sz, _int, _vp]),
    "nt_linear_forward": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "nt_linear_backward_input": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _vp]),
    "nt_linear_backward_weight_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "nt_linear_backward_weight": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _int, _vp]),
    "nt_embedding_bag_sum": (_int, [_vp, _i64, _i64p, _i64, _i64, _i64, _vp, _i32p, _int, _vp]),
    "nt_embedding_bag_backward_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "nt_embedding_bag_backward": (_int, [_vp, _i64p, _i64, _i64, _i64, _i64, _vp, _vp, _sz, _int, _vp]),
    "nt_embed_edge_init": (_int, [_vp, _i64, _vp, _i64, _i64p, _i64, _i64p, _i64, _i32p, _i64, _i64, _i64, _vp, _i32p, _int, _vp]),
    "nt_embed_edge_init_backward_workspace_bytes": (_sz, [_i64, _i64, _i64, _i64]),
    "nt_embed_edge_init_backward": (_int, [_vp, _i64p, _i64, _i64p, _i64, _i32p, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _int, _vp]),
    "nt_seg_max": (_int, [_vp, _i64, _i32p, _i32p, _i64, _vp, _i32p, _int, _vp]),
    "nt_seg_extreme": (_int, [_vp, _i64, _i32p, _i32p, _i64, _int, _f32, _int, _vp, _i32p, _int, _vp]),
    "nt_seg_max_backward": (_int, [_vp, _i32p, _i32p, _i64, _i64, _vp, _int, _vp]),
    "nt_row_dot": (_int, [_vp, _vp, _i32p, _i64, _i64, _i64, _f32, _f32, _vp, _int, _vp]),
    "nt_seg_softmax": (_int, [_vp, _i32p, _i32p, _i64, _vp, _int, _vp]),
    "nt_seg_softmax_backward": (_int, [_vp, _vp, _i32p, _i32p, _i64, _vp, _int, _vp]),
    "nt_seg_weighted_sum": (_int, [_vp, _vp, _i64, _i32p, _i32p, _i64, _f32, _vp, _int, _vp]),
    "nt_row_scale_gather": (_int, [_vp, _vp, _i32p, _i64, _i64, _f32, _vp, _int, _vp]),
    "nt_weighted_col_sum_workspace_bytes": (_sz, [_i64, _i64]),
    "nt_weighted_col_sum": (_int, [_vp, _vp, _i64, _i64, _f32, _vp, _vp, _sz, _int, _vp]),
    "nt_dropout_mask": (_int, [_i64, _i64, _f32, _u64, _u64, _vp, _vp]),
}

_lock = threading.Lock()
_lib: C.CDLL | None = None


def _load() -> C.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        # in-tree build (nvcc cross-compiles for sm_100a without a GPU). build() is a no-op while the digest of csrc/ + the headers
        # matches the stamp written beside the objects, so an edited source never runs under a stale library.
        from . import build as _build

        try:
            _build.build()
        except Exception as exc:  # pragma: no cover - depends on the toolchain
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"notorch_b200: {LIB_PATH} is missing and could not be built ({exc}). "
                    "There is no CPU or PyTorch fallback; run `python -m notorch_b200.build`."
                ) from exc
            import warnings

            warnings.warn(f"notorch_b200: {LIB_PATH} is older than its sources and could not be rebuilt ({exc}); loading it as is", stacklevel=3)
        lib = C.CDLL(LIB_PATH)
        lib.nt_version.restype = C.c_int
        if lib.nt_version() != ABI_VERSION:
            raise RuntimeError(f"notorch_b200: {LIB_PATH} reports ABI version {lib.nt_version()}, this binding needs {ABI_VERSION}; "
                               "rebuild with `python -m notorch_b200.build --force`")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def lib() -> C.CDLL:
    return _lib if _lib is not None else _load()


def last_error() -> str:
    msg = lib().nt_last_error_string()
    return msg.decode(errors="replace") if msg else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"notorch_b200: {what} failed with status {rc}: {last_error()}")
