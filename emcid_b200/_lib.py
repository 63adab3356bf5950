"""ctypes binding of ``libemcid_b200.so`` (the C ABI in ``include/emcid_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C emcid_b200/csrc``.
There is no fallback: if the library is missing, ``lib()`` raises, and every compute entry point
fails with ``EMCID_ERR_UNSUPPORTED`` on a device that is not sm_100.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# EMCID_B200_LIB: another build of the same library (measurement aid, e.g. one compiled with -DEMCID_POTRF_TIMING)
LIB_PATH = os.environ.get("EMCID_B200_LIB") or os.path.join(_HERE, "libemcid_b200.so")

EMCID_OK = 0
ACT_QUICK_GELU = 0
ACT_GELU_ERF = 1
ACT_NONE = 2

_ACT_BY_NAME = {"quick_gelu": ACT_QUICK_GELU, "gelu": ACT_GELU_ERF, "gelu_erf": ACT_GELU_ERF, "none": ACT_NONE}


class EmcidError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libemcid_b200 error {code}: {msg}")
        self.code = code


# name -> (restype, argtypes); mirrors include/emcid_b200.h one to one (tests check the export list).
SIGNATURES = {
    "emcid_last_error": (c_char_p, []),
    "emcid_version": (c_int, []),
    "emcid_device_check": (c_int, [c_int]),
    "emcid_hang_code": (c_uint, []),
    "emcid_gemm3x_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "emcid_gemm3x_nt": (
        c_int,
        [c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_longlong,
         c_float, c_float, c_int, c_void_p, c_size_t, c_void_p],
    ),
    "emcid_mom2_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "emcid_mom2_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t]),
    "emcid_mom2_set_chunks": (c_int, [c_void_p, c_int, c_int]),
    "emcid_mom2_set_precision": (c_int, [c_void_p, c_int]),
    "emcid_mom2_set_weights": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    "emcid_mom2_accumulate": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_longlong, c_void_p]),
    "emcid_mom2_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "emcid_mom2_reset": (c_int, [c_void_p, c_void_p]),
    "emcid_mom2_profile": (c_int, [c_void_p, c_int]),
    "emcid_mom2_get_profile": (c_int, [c_void_p, ctypes.POINTER(c_double)]),
    "emcid_mom2_destroy": (c_int, [c_void_p]),
    "emcid_nccl_available": (c_int, []),
    "emcid_mom2_reduce": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "emcid_mom2_broadcast": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "emcid_mom2_state_elems": (c_size_t, [c_int]),
    "emcid_mom2_export_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "emcid_mom2_import_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "emcid_symmetrize_lower": (c_int, [c_void_p, c_int, c_longlong, c_void_p]),
    "emcid_checksum_tensors": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "emcid_fixed_random_subset": (c_int, [c_longlong, c_longlong, ctypes.POINTER(c_longlong), c_longlong]),
    "emcid_clip_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                  c_longlong, c_int]),
    "emcid_clip_set_embeddings": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "emcid_clip_set_layer": (c_int, [c_void_p, c_int, ctypes.POINTER(c_void_p), c_void_p]),
    "emcid_clip_update_layer": (c_int, [c_void_p, c_int, ctypes.POINTER(c_void_p), c_uint, c_void_p]),
    "emcid_clip_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   ctypes.POINTER(c_int), ctypes.POINTER(c_void_p), c_void_p, c_void_p]),
    "emcid_clip_forward_keys": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                        c_void_p, c_void_p, c_int, c_void_p]),
    "emcid_clip_set_final_norm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "emcid_clip_forward_final": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                         c_void_p, c_void_p]),
    "emcid_clip_profile": (c_int, [c_void_p, c_int]),
    "emcid_clip_get_profile": (c_int, [c_void_p, ctypes.POINTER(c_double)]),
    "emcid_clip_launches": (c_longlong, [c_void_p]),
    "emcid_release_cached_memory": (c_int, []),
    "emcid_clip_destroy": (c_int, [c_void_p]),
    "emcid_solve_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "emcid_solve_layers": (
        c_int,
        [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_longlong, c_void_p, c_longlong, c_double, c_double,
         ctypes.POINTER(c_double), c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p],
    ),
    "emcid_read_npz_f32": (c_int, [ctypes.POINTER(c_char_p), c_int, c_char_p, c_void_p, c_longlong, ctypes.POINTER(c_int)]),
    "emcid_delta_update": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "emcid_factor_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_void_p, c_double, c_void_p, c_void_p]),
    "emcid_factor_solve_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "emcid_factor_solve": (
        c_int,
        [c_void_p, c_int, c_int, c_void_p, c_longlong, c_void_p, c_longlong, c_double, c_double, c_void_p, c_void_p,
         c_void_p, c_int, c_void_p, c_size_t, c_void_p, c_void_p],
    ),
    "emcid_factor_destroy": (c_int, [c_void_p]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C emcid_b200/csrc` (there is no CPU fallback for the EMCID hot path)."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error() -> str:
    msg = lib().emcid_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code: int) -> None:
    if code != EMCID_OK:
        raise EmcidError(code, last_error())


def act_code(name: str) -> int:
    try:
        return _ACT_BY_NAME[name]
    except KeyError as e:
        raise ValueError(f"unsupported activation {name!r} (quick_gelu | gelu)") from e


def current_stream_ptr() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> int:
    """Device/host address of a contiguous torch tensor (0 for None)."""
    return 0 if t is None else int(t.data_ptr())


def gemm3x_nt(A, B, C=None, alpha: float = 1.0, beta: float = 0.0, lower: bool = False,
              streamk: bool = False, n128: bool = False, chunk: int = 0, f16: bool = False):
    """C = alpha * A @ B.T + beta * C on tcgen05 with the 3xTF32 split (fp32 CUDA tensors)."""
    import torch

    assert A.is_cuda and B.is_cuda and A.dtype == torch.float32 and B.dtype == torch.float32
    assert A.dim() == 2 and B.dim() == 2 and A.shape[1] == B.shape[1]
    assert A.stride(1) == 1 and B.stride(1) == 1
    M, K = A.shape
    N = B.shape[0]
    if C is None:
        C = torch.zeros(M, N, device=A.device, dtype=torch.float32)
    assert C.stride(1) == 1
    ws_bytes = lib().emcid_gemm3x_workspace_bytes(M, N, K)
    ws = torch.empty(ws_bytes, device=A.device, dtype=torch.uint8)
    flags = (1 if lower else 0) | (2 if streamk else 0) | (4 if n128 else 0) | (8 if f16 else 0) | ((chunk & 0xFF) << 8)
    with torch.cuda.device(A.device):
        check(lib().emcid_gemm3x_nt(M, N, K, ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(C), C.stride(0),
                                    alpha, beta, flags, ptr(ws), ws_bytes, current_stream_ptr()))
    return C
