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


class SolveNotConverged(_lib.EmcidError):
    """The fp32-class factorisation did not contract: the system is too ill-conditioned for it (status bit 1)."""


def solve_layers(C32: torch.Tensor, Kt: torch.Tensor, St: torch.Tensor, mom2_update_weight: float,
                 scale: float, layers_left: Sequence[int], refine_steps: int = DEFAULT_REFINE_STEPS,
                 check: bool = True, strict: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """C32 [B, d, d] fp32, Kt [B, n, d] fp32, St [B, n, h] fp32 (CUDA).  layers_left[b] = L - i.
    Returns (adj_k [B, d, n] fp64, resid [B, h, n] fp64, dW [B, h, d] fp32) on the same device.
    check: read the status word (one D2H sync) and raise EmcidError on a Cholesky breakdown; a refinement that used all
    its sweeps without reaching its target is a RuntimeWarning, or SolveNotConverged with `strict` (callers that have a
    fallback, emcid_main._solve_one_layer)."""
    if C32.dim() == 2:
        C32, Kt, St = C32[None], Kt[None], St[None]
    assert C32.is_cuda and Kt.is_cuda and St.is_cuda, "emcid_b200.solve needs CUDA tensors (no CPU path)"
    C32 = C32.contiguous().float()
    Kt = Kt.contiguous().float()
    St = St.contiguous().float()
    B, d, _ = C32.shape
    n, h = Kt.shape[1], St.shape[2]
    assert Kt.shape == (B, n, d) and St.shape == (B, n, h) and len(layers_left) == B
    if d % 128:
        # the blocked factorisation works on 128-wide panels: embed the system in the next multiple of 128 with a
        # decoupled identity block (zero key columns there), solve, and cut the padding off again
        dp = -(-d // 128) * 128
        Cp = torch.zeros(B, dp, dp, dtype=torch.float32, device=C32.device)
        Cp[:, :d, :d] = C32
        idx = torch.arange(d, dp, device=C32.device)
        Cp[:, idx, idx] = 1.0
        Kp = torch.zeros(B, n, dp, dtype=torch.float32, device=C32.device)
        Kp[:, :, :d] = Kt
        adj_k, resid, dW = solve_layers(Cp, Kp, St, mom2_update_weight, scale, layers_left, refine_steps, check, strict)
        return adj_k[:, :d].contiguous(), resid, dW[:, :, :d].contiguous()
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
        _check_status(int(status.item()), "lambda*C + K K^T", strict)
    return adj_k, resid, dW


def _check_status(st: int, what: str, strict: bool = False) -> None:
    """status bits of include/emcid_b200.h: bit 0 = breakdown (raise), bit 1 = refinement target not reached (warn, or
    raise SolveNotConverged when `strict`)."""
    if st & 1:
        raise _lib.EmcidError(-4, f"Cholesky breakdown: {what} is not positive definite in fp32-class arithmetic (status {st})")
    if st & 2:
        if strict:
            raise SolveNotConverged(-4, f"iterative refinement of {what} did not reach its target (status {st})")
        import warnings

        warnings.warn(f"emcid_b200.solve: the iterative refinement of {what} did not reach its target "
                      "(very ill-conditioned system); the update is the best available", RuntimeWarning)


class CachedFactor:
    """Factorisation of ``A = mom2_update_weight * C32`` kept on the device for repeated edits with the same
    covariance (``emcid_factor_create`` / ``emcid_factor_solve``).

    The reference re-solves a fresh ``d x d`` system per edit and layer even when only the keys changed
    (sequential editing, experiments/sequential_editing.py:98-171; the debias factor search,
    emcid/emcid_main.py:1460-1472; the layer ablation, experiments/ablation.py:332-338).  ``solve`` returns the
    same ``(adj_k, resid, dW)`` as ``solve_layers`` for one layer in ``O(d^2 n)`` instead of ``O(d^3)`` work."""

    def __init__(self, C32: torch.Tensor, mom2_update_weight: float):
        assert C32.is_cuda and C32.dim() == 2 and C32.shape[0] == C32.shape[1], \
            "emcid_b200.solve needs a square CUDA covariance (no CPU path)"
        C32 = C32.contiguous().float()
        self.device = C32.device
        self.d = int(C32.shape[0])
        self.mom2_update_weight = float(mom2_update_weight)
        self._handle = ctypes.c_void_p()
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        lib = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(lib.emcid_factor_create(ctypes.byref(self._handle), self.device.index, self.d, _lib.ptr(C32),
                                               self.mom2_update_weight, _lib.ptr(status), _lib.current_stream_ptr()))
        st = int(status.item())
        if st != 0:
            self.close()
            raise _lib.EmcidError(-4, f"Cholesky breakdown: lambda*C is not positive definite (status {st})")

    def solve(self, Kt: torch.Tensor, St: torch.Tensor, scale: float, layers_left: int,
              refine_steps: int = DEFAULT_REFINE_STEPS, check: bool = True, strict: bool = False
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Kt [n, d] fp32, St [n, h] fp32 (CUDA) -> (adj_k [d, n] fp64, resid [h, n] fp64, dW [h, d] fp32)."""
        assert self._handle, "CachedFactor is closed"
        assert Kt.is_cuda and St.is_cuda and Kt.device == self.device and St.device == self.device
        Kt = Kt.contiguous().float()
        St = St.contiguous().float()
        n, d = Kt.shape
        h = St.shape[1]
        assert d == self.d and St.shape[0] == n
        dev = self.device
        adj_k = torch.empty(d, n, dtype=torch.float64, device=dev)
        resid = torch.empty(h, n, dtype=torch.float64, device=dev)
        dW = torch.empty(h, d, dtype=torch.float32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        lib = _lib.lib()
        ws = _workspace(dev, lib.emcid_factor_solve_workspace_bytes(d, h, n))
        with torch.cuda.device(dev):
            _lib.check(lib.emcid_factor_solve(self._handle, h, n, _lib.ptr(Kt), d, _lib.ptr(St), h, float(scale),
                                              1.0 / float(layers_left), _lib.ptr(adj_k), _lib.ptr(resid), _lib.ptr(dW),
                                              int(refine_steps), _lib.ptr(ws), ws.numel(), _lib.ptr(status),
                                              _lib.current_stream_ptr()))
        if check:
            _check_status(int(status.item()), "lambda*C / I + Ks^T A^-1 Ks", strict)
        return adj_k, resid, dW

    def close(self) -> None:
        if self._handle:
            _lib.lib().emcid_factor_destroy(self._handle)
            self._handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
