"""Closed-form EMCID update on the GPU: thin wrapper over ``emcid_solve_layers`` (C ABI).

Computes, for a batch of independent layers, exactly the quantities of the reference's solve block
(emcid/emcid_main.py:1037-1050): ``adj_k``, ``resid`` (fp64) and ``float(resid @ adj_k.T)``.
"""
from __future__ import annotations

import ctypes
from typing import Sequence, Tuple

import torch

from . import _lib

DEFAULT_REFINE_STEPS = -1  # adaptive (see include/emcid_b200.h)

_WS = {}


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    ws = _WS.get(device.index)
    if ws is None or ws.numel() < nbytes:
        _WS[device.index] = ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return ws


def release_workspace() -> None:
    _WS.clear()


def solve_layers(C32: torch.Tensor, Kt: torch.Tensor, St: torch.Tensor, mom2_update_weight: float,
                 scale: float, layers_left: Sequence[int], refine_steps: int = DEFAULT_REFINE_STEPS,
                 check: bool = True) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """C32 [B, d, d] fp32, Kt [B, n, d] fp32, St [B, n, h] fp32 (CUDA).  layers_left[b] = L - i.
    Returns (adj_k [B, d, n] fp64, resid [B, h, n] fp64, dW [B, h, d] fp32) on the same device."""
    if C32.dim() == 2:
        C32, Kt, St = C32[None], Kt[None], St[None]
    assert C32.is_cuda and Kt.is_cuda and St.is_cuda, "emcid_b200.solve needs CUDA tensors (no CPU path)"
    C32 = C32.contiguous().float()
    Kt = Kt.contiguous().float()
    St = St.contiguous().float()
    B, d, _ = C32.shape
    n, h = Kt.shape[1], St.shape[2]
    assert Kt.shape == (B, n, d) and St.shape == (B, n, h) and len(layers_left) == B
    dev = C32.device
    adj_k = torch.empty(B, d, n, dtype=torch.float64, device=dev)
    resid = torch.empty(B, h, n, dtype=torch.float64, device=dev)
    dW = torch.empty(B, h, d, dtype=torch.float32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = _lib.lib()
    nbytes = lib.emcid_solve_workspace_bytes(B, d, h, n)
    ws = _workspace(dev, nbytes)
    inv_left = (ctypes.c_double * B)(*[1.0 / float(x) for x in layers_left])
    with torch.cuda.device(dev):
        _lib.check(lib.emcid_solve_layers(dev.index, B, d, h, n, _lib.ptr(C32), _lib.ptr(Kt), d, _lib.ptr(St), h,
                                          float(mom2_update_weight), float(scale), inv_left, _lib.ptr(adj_k),
                                          _lib.ptr(resid), _lib.ptr(dW), int(refine_steps), _lib.ptr(ws), ws.numel(),
                                          _lib.ptr(status), _lib.current_stream_ptr()))
    if check:
        st = int(status.item())
        if st != 0:
            raise _lib.EmcidError(-4, f"Cholesky breakdown: lambda*C + K K^T is not positive definite (status {st})")
    return adj_k, resid, dW
