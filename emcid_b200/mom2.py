"""Device-side second-moment accumulator: thin object wrapper over the ``emcid_mom2_*`` C ABI.

One ``Mom2Accumulator`` per edited layer.  ``add(X, valid)`` consumes the LN2 output ``X`` (the
argument of the CLIP MLP) and performs, on the GPU and without materialising the d-wide features,
what the reference does with ``flatten_masked_batch`` + ``SecondMoment.add``
(dsets/stat_dataset.py:166-172, util/runningstats.py:483-493).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

from . import _lib

_WORKSPACES = {}  # (device index, bytes) -> uint8 tensor shared by all accumulators on that device


def _shared_workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    key = device.index
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


class Mom2Accumulator:
    PRECISIONS = {"tf32x3": 0, "f16x3": 1}

    def __init__(self, device, d: int, h: int, act: str = "quick_gelu", slab_tokens: int = 0,
                 fc1_chunk: Optional[int] = None, syrk_chunk: Optional[int] = None, precision: Optional[str] = None):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("Mom2Accumulator needs a CUDA (sm_100a) device; there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device, self.d, self.h = device, int(d), int(h)
        lib = _lib.lib()
        nbytes = lib.emcid_mom2_workspace_bytes(self.d, self.h, int(slab_tokens))
        # a shared scratch keeps the A^T slab of every layer in the same L2-resident addresses
        self._ws = _shared_workspace(device, nbytes)
        self._h = ctypes.c_void_p()
        _lib.check(lib.emcid_mom2_create(ctypes.byref(self._h), device.index, self.d, self.h, _lib.act_code(act),
                                         int(slab_tokens), _lib.ptr(self._ws), self._ws.numel()))
        if precision is None:
            precision = os.environ.get("EMCID_MOM2_PRECISION")
        if precision is not None:
            _lib.check(lib.emcid_mom2_set_precision(self._h, self.PRECISIONS[precision]))
        if fc1_chunk or syrk_chunk:
            _lib.check(lib.emcid_mom2_set_chunks(self._h, int(fc1_chunk or 1), int(syrk_chunk or 2)))
        self._keep = []  # tensors that must outlive asynchronous launches

    def _stream(self) -> int:
        return int(torch.cuda.current_stream(self.device).cuda_stream)

    def set_weights(self, W1: torch.Tensor, b1: Optional[torch.Tensor]) -> None:
        W1 = W1.detach()
        assert W1.is_cuda and W1.dtype == torch.float32 and W1.shape == (self.d, self.h) and W1.stride(1) == 1
        if b1 is not None:
            b1 = b1.detach().contiguous()
            assert b1.dtype == torch.float32 and b1.shape == (self.d,)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_set_weights(self._h, _lib.ptr(W1), W1.stride(0), _lib.ptr(b1),
                                                         self._stream()))

    def add(self, X: torch.Tensor, valid: Optional[torch.Tensor] = None) -> None:
        """X: [..., h] fp32 CUDA; valid: matching [...] mask (bool/uint8/int), None = every row."""
        X = X.detach()
        assert X.is_cuda and X.dtype == torch.float32 and X.shape[-1] == self.h
        X2 = X.reshape(-1, self.h)
        if X2.stride(1) != 1 or X2.stride(0) % 4 != 0 or X2.data_ptr() % 16 != 0:
            X2 = X2.contiguous()
        v = None
        if valid is not None:
            v = valid.reshape(-1)
            assert v.numel() == X2.shape[0]
            v = (v != 0).to(torch.uint8).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_accumulate(self._h, _lib.ptr(X2), X2.stride(0), _lib.ptr(v),
                                                        X2.shape[0], self._stream()))
        # the launches are asynchronous: make the caching allocator aware the stream still uses them
        X2.record_stream(torch.cuda.current_stream(self.device))
        if v is not None:
            v.record_stream(torch.cuda.current_stream(self.device))

    def finalize(self):
        """Returns (mom2 [d, d] fp32 CUDA tensor, full symmetric; count as 0-d int64 CUDA tensor)."""
        out = torch.empty(self.d, self.d, dtype=torch.float32, device=self.device)
        count = torch.zeros((), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_finalize(self._h, _lib.ptr(out), _lib.ptr(count), self._stream()))
        return out, count

    def reduce(self, nccl_comm: int, root: int) -> None:
        """The exchange step of a caption-sharded pass (emcid_mom2_reduce): sums the accumulators of all ranks of
        `nccl_comm` (address of an ncclComm_t) onto `root`; stream-ordered on the current stream."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_reduce(self._h, ctypes.c_void_p(int(nccl_comm)), int(root), self._stream()))

    def export_state(self, out: Optional[torch.Tensor] = None):
        """(packed lower triangle of the fp64 sums [d (d + 1) / 2] float64 CUDA, count 0-d int64 CUDA): what a resumable
        pass checkpoints (emcid_mom2_export_state; folds the fp32 accumulator first)."""
        n = int(_lib.lib().emcid_mom2_state_elems(self.d))
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=self.device)
        assert out.is_cuda and out.dtype == torch.float64 and out.numel() == n and out.is_contiguous()
        count = torch.zeros((), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_export_state(self._h, _lib.ptr(out), _lib.ptr(count), self._stream()))
        return out, count

    def import_state(self, packed: torch.Tensor, count) -> None:
        """Inverse of export_state on a fresh or reset accumulator."""
        packed = packed.to(self.device, dtype=torch.float64).contiguous()
        assert packed.numel() == int(_lib.lib().emcid_mom2_state_elems(self.d))
        count = torch.as_tensor(count, dtype=torch.int64).reshape(()).to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_import_state(self._h, _lib.ptr(packed), _lib.ptr(count), self._stream()))
        s = torch.cuda.current_stream(self.device)
        packed.record_stream(s)
        count.record_stream(s)

    def profile(self, enable: bool = True) -> None:
        _lib.check(_lib.lib().emcid_mom2_profile(self._h, 1 if enable else 0))

    def get_profile(self) -> dict:
        out = (ctypes.c_double * 8)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_get_profile(self._h, out))
        keys = ("fc1_ms", "fc1_launches", "fc1_rows", "syrk_ms", "syrk_launches", "syrk_rows", "launches")
        return {k: float(out[i]) for i, k in enumerate(keys)}

    def reset(self) -> None:
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().emcid_mom2_reset(self._h, self._stream()))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            torch.cuda.synchronize(self.device)
            _lib.lib().emcid_mom2_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
