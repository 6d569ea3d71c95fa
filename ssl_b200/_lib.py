"""ctypes binding of libssl_b200.so (the C ABI of include/ssl_b200.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call fails, this module
raises.  Build it with ``python -m ssl_b200.csrc.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSL_B200_LIB names another build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("SSL_B200_LIB") or os.path.join(_HERE, "csrc", "libssl_b200.so")

F32, BF16, F16 = 0, 1, 2
ROWS_RAW, ROWS_EXP, ROWS_NORM = 0, 1, 2
ABI_VERSION = 3
PATH_AUTO, PATH_POINT, PATH_PLANE = 0, 1, 2

_c_int, _c_float, _c_void_p, _c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/ssl_b200.h one to one
SIGNATURES = {
    "ssl_b200_abi_version": (_c_int, []),
    "ssl_b200_last_error": (ctypes.c_char_p, []),
    "ssl_b200_compute_similarity": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                             _c_int, _c_void_p]),
    "ssl_b200_compute_similarity_backward": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                                      _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "ssl_b200_edge_list_workspace_bytes": (_c_size_t, [ctypes.c_int64]),
    "ssl_b200_build_edge_list": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_int,
                                          _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "ssl_b200_ssg_rows_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                           _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _c_void_p,
                                           _c_void_p, _c_void_p]),
    "ssl_b200_rows_grad_to_distance_grad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                                                     _c_float, _c_int, _c_void_p]),
    "ssl_b200_ssg_rows_backward": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                            _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p]),
    "ssl_b200_row_loss_blocks": (_c_int, []),
    "ssl_b200_row_loss": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_int,
                                   _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "ssl_b200_plane_supported": (_c_int, [_c_int, _c_int, _c_int]),
    "ssl_b200_plane_rows_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int]),
    "ssl_b200_plane_rows_forward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                             _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p,
                                             _c_size_t, _c_void_p]),
    "ssl_b200_rows_from_distance": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                             _c_int, _c_void_p]),
    "ssl_b200_plane_rows_backward_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int]),
    "ssl_b200_plane_rows_backward": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p,
                                              _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_size_t,
                                              _c_void_p]),
    "ssl_b200_loss_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int]),
    "ssl_b200_loss_forward_backward": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                                _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_float, _c_float,
                                                _c_int, _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p,
                                                _c_size_t, _c_int, _c_void_p]),
    "ssl_b200_loss_step_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int]),
    "ssl_b200_loss_step": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int,
                                    _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _c_float,
                                    _c_float, _c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "ssl_b200_loss_from_terms": (_c_int, [_c_void_p, _c_int, _c_float, _c_float, _c_float, _c_void_p, _c_void_p]),
    "ssl_b200_loss_export_distance_grad": (_c_int, [_c_void_p, _c_size_t, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                                    _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_void_p]),
    "ssl_b200_loss_step_host": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                         _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _c_float, _c_float,
                                         _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "ssl_b200_release_host_arena": (_c_int, []),
    "ssl_b200_crop": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "ssl_b200_pool_exchange": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, ctypes.c_int64, _c_int,
                                        _c_void_p]),
    "ssl_b200_profile_enable": (_c_int, [_c_int]),
    "ssl_b200_profile_num_stages": (_c_int, []),
    "ssl_b200_profile_stage_name": (ctypes.c_char_p, [_c_int]),
    "ssl_b200_profile_read": (_c_int, [_c_void_p, _c_void_p]),
    "ssl_b200_launch_count": (ctypes.c_uint64, []),
    "ssl_b200_laplacian_mask": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_void_p]),
}

_lib = None


class SSLB200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SSLB200Error(
            f"{LIB_PATH} is missing: build the sm_100a kernels with `python -m ssl_b200.csrc.build`. "
            "ssl_b200 has no CPU / PyTorch fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ssl_b200_abi_version() != ABI_VERSION:
        raise SSLB200Error(f"libssl_b200.so ABI {lib.ssl_b200_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.ssl_b200_last_error()
        raise SSLB200Error(f"{name} failed ({rc}): {msg.decode() if msg else '?'}")


def profile_enable(on: bool) -> None:
    call("ssl_b200_profile_enable", 1 if on else 0)


def profile_read() -> dict:
    """{stage name: (milliseconds, bracketed launches)} since the last read (waits for the events)."""
    lib = load()
    n = lib.ssl_b200_profile_num_stages()
    ms = (ctypes.c_float * n)()
    cnt = (ctypes.c_int * n)()
    call("ssl_b200_profile_read", ctypes.cast(ms, _c_void_p), ctypes.cast(cnt, _c_void_p))
    return {lib.ssl_b200_profile_stage_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n) if cnt[i]}


def dtype_code(dtype) -> int:
    import torch
    try:
        return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}[dtype]
    except KeyError:
        raise TypeError(f"ssl_b200 kernels take float32 / bfloat16 / float16 images, got {dtype}") from None
