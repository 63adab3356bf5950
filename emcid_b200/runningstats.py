"""Statistics objects and the on-disk cache of the second-moment pass (seam B2 of SURVEY.md §8b).

What has to stay as the reference has it (util/runningstats.py of SilentView/EMCID) is the CONTRACT, restated here from
its observable behaviour, not its code:

  object protocol   SecondMoment (:469-511): add(a [T, d]) -> count += T, mom2 += a^T a ; moment() = mom2 / count ;
                    to_(device) ; state_dict() / load_state_dict() ; attributes .count (python int) and .mom2 (the FULL
                    symmetric [d, d] tensor).  CombinedStat (:1347-1388): named children, state keys "<name>.<key>".
  npz layout        numpy.savez, uncompressed, no pickle (:1496-1512):
                        mom2.constructor  <U32     "util.runningstats.SecondMoment()"
                        mom2.count        int64 () rows added
                        mom2.mom2         float32 (d, d), the raw sum, not divided by count
                        sample_size       int64 () — or, when the pass had no sample_size, the float64 NaN with bit
                                          pattern 0xfff8000000000002 that stands for None (:1417-1454)
  cache hit         the file loads AND every tally argument (here: sample_size) equals the stored one (:1469-1493).
  caption subset    random.Random(1).shuffle(range(len(ds)))[:sample_size] (:1551-1556, via :1598-1600).

A file written here loads in the reference and vice versa (tests/test_host_logic.py checks both directions against the
live reference).  `SecondMoment.add` runs on the sm_100a kernels only: CUDA float32 batches go through the tcgen05
lower-triangle GEMM of the library (emcid_gemm3x_nt) and are mirrored on read; anything else raises.  The other two
statistics of the reference's STAT_TYPES (mean, norm_mean; emcid/layer_stats.py:26-30) are not on the edit path and are
not reimplemented: layer_stats delegates such requests to an installed reference or refuses them.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import numpy
import torch

SECOND_MOMENT_CONSTRUCTOR = "util.runningstats.SecondMoment()"


class SecondMoment:
    """Raw second moment sum_t a_t a_t^T and the number of rows behind it."""

    def __init__(self, split_batch: bool = True, state=None):
        self.count = 0
        self.split_batch = split_batch      # accepted like the reference's argument; batches are never split here
        self._sum: Optional[torch.Tensor] = None
        self._upper_stale = False           # the kernels only write the lower triangle; the mirror image is made on read
        if state is not None:
            self.load_state_dict(_as_state(state))

    # .mom2 is always the full symmetric matrix, whatever the kernels have written so far
    @property
    def mom2(self) -> Optional[torch.Tensor]:
        if self._upper_stale and self._sum is not None:
            from . import _lib

            with torch.cuda.device(self._sum.device):
                _lib.check(_lib.lib().emcid_symmetrize_lower(_lib.ptr(self._sum), self._sum.shape[0], self._sum.stride(0),
                                                             _lib.current_stream_ptr()))
            self._upper_stale = False
        return self._sum

    @mom2.setter
    def mom2(self, value: Optional[torch.Tensor]) -> None:
        self._sum, self._upper_stale = value, False

    def add(self, a) -> None:
        a = torch.as_tensor(a)
        if a.dim() == 0:
            a = a.reshape(1)
        width = 1
        for n in a.shape[1:]:
            width *= int(n)
        a = a.reshape(a.shape[0], width)    # rows = samples, everything else is the feature vector
        known = getattr(self, "_width", None)
        if known is None:
            self._width = width
        else:
            assert width == known, f"feature width changed from {known} to {width}"
        if a.shape[0] == 0:
            return
        if not (a.is_cuda and a.dtype == torch.float32):
            raise RuntimeError("emcid_b200.SecondMoment.add needs a CUDA float32 batch (sm_100a kernels only; "
                               f"got {a.device}/{a.dtype})")
        from . import _lib

        d = a.shape[1]
        if self._sum is None or self.count == 0:
            self._sum, self._upper_stale = torch.zeros(d, d, dtype=torch.float32, device=a.device), False
        elif self._sum.device != a.device:
            self.to_(a.device)
        # [d, T] with unit stride along T: the contraction (token) axis is the fast one of both operands
        k_major = torch.empty(d, a.shape[0], dtype=torch.float32, device=a.device).copy_(a.t())
        _lib.gemm3x_nt(k_major, k_major, self._sum, alpha=1.0, beta=1.0, lower=True, streamk=True)
        self._upper_stale = True
        self.count += a.shape[0]

    def moment(self) -> torch.Tensor:
        return self.mom2 / self.count

    def to_(self, device) -> None:
        if self._sum is not None:
            self._sum = self.mom2.to(device)

    def state_dict(self) -> Dict[str, object]:
        return {"constructor": SECOND_MOMENT_CONSTRUCTOR, "count": self.count, "mom2": self.mom2.cpu().numpy()}

    def load_state_dict(self, state) -> None:
        self.count = int(state["count"])
        self.mom2 = torch.from_numpy(numpy.asarray(state["mom2"]))


class CombinedStat:
    """Named bundle of statistics; `stat.<name>` reaches a child, state keys are "<name>.<key>"."""

    def __init__(self, state=None, **children):
        self._objs = dict(children)
        if state is not None:
            self.load_state_dict(_as_state(state))

    def __getattr__(self, name):
        children = self.__dict__.get("_objs", {})
        if name in children:
            return children[name]
        raise AttributeError(name)

    def add(self, batch, *args, **kwargs) -> None:
        for child in self._objs.values():
            child.add(batch, *args, **kwargs)

    def to_(self, device) -> None:
        for child in self._objs.values():
            child.to_(device)

    def state_dict(self) -> Dict[str, object]:
        return {f"{name}.{key}": value for name, child in self._objs.items() for key, value in child.state_dict().items()}

    def load_state_dict(self, state) -> None:
        for name, child in self._objs.items():
            child.load_state_dict({key[len(name) + 1:]: state[key] for key in _keys(state) if key.startswith(name + ".")})


# ---------------------------------------------------------------------------------------------------------
# the npz cache
# ---------------------------------------------------------------------------------------------------------
_NONE_BITS = numpy.uint64(0xFFF8000000000002)       # quiet NaN with payload 2: how the layout spells None
null_numpy_value = numpy.array(_NONE_BITS).view(numpy.float64).copy()


def is_null_numpy_value(v) -> bool:
    return (isinstance(v, numpy.ndarray) and v.ndim == 0 and v.dtype == numpy.float64
            and v.view(numpy.uint64) == _NONE_BITS)


def _keys(state):
    return state.files if hasattr(state, "files") else state.keys()


def _as_state(state):
    """A state dict, or the path of an npz holding one (None entries unboxed)."""
    if isinstance(state, (str, os.PathLike)):
        with numpy.load(state) as dat:
            return {k: (None if is_null_numpy_value(dat[k]) else dat[k]) for k in dat.files}
    return state


def load_cached_state(cachefile, args, quiet: bool = False, throw: bool = False):
    """The stored state if `cachefile` (a path, or a dict standing in for one) is a hit for the tally arguments `args`,
    else None."""
    if cachefile is None:
        return None
    try:
        if isinstance(cachefile, dict):
            label, dat = "state", cachefile
        else:
            label, dat = cachefile, _as_state(os.fspath(cachefile))
    except (FileNotFoundError, ValueError, OSError, EOFError) as e:      # missing, truncated or not an npz: a miss
        if throw:
            raise e
        return None
    for name, wanted in args.items():
        stored = dat.get(name) if name in dat else None
        if name not in dat or not _same(stored, wanted):
            if not quiet:
                print(f"{label} {name} changed from {stored} to {wanted}")
            return None
    if not quiet:
        print(f"Loading cached {label}")
    return dat


def _same(stored, wanted) -> bool:
    if stored is None or wanted is None:
        return stored is None and wanted is None
    return bool(stored == wanted)


def save_cached_state(cachefile, obj, args) -> None:
    """obj.state_dict() merged with the tally arguments, as one uncompressed npz (None boxed as the NaN above).
    Written to a temporary name first and renamed, so that a reader (or a crash) never sees half a file."""
    if cachefile is None:
        return
    dat = dict(obj.state_dict())
    for name, value in args.items():
        assert name not in dat or dat[name] == value
        dat[name] = value
    if isinstance(cachefile, dict):
        cachefile.clear()
        cachefile.update(dat)
        return
    path = os.fspath(cachefile)
    if not path.endswith(".npz"):
        path += ".npz"                       # numpy.savez appends it; keep the final name predictable
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    tmp = f"{path}.{os.getpid()}.tmp.npz"
    numpy.savez(tmp, **{k: (null_numpy_value if v is None else v) for k, v in dat.items()})
    os.replace(tmp, path)


# ---------------------------------------------------------------------------------------------------------
# WHICH captions a pass visits
# ---------------------------------------------------------------------------------------------------------
def fixed_random_subset(dataset_len: int, sample_size: int, seed: int = 1) -> numpy.ndarray:
    """random.Random(seed).shuffle(list(range(dataset_len)))[:sample_size] as int64, computed by the library's restatement
    of CPython's generator (emcid_fixed_random_subset; tests compare it with `random` itself)."""
    from . import _lib

    out = numpy.empty(int(sample_size), dtype=numpy.int64)
    _lib.check(_lib.lib().emcid_fixed_random_subset(int(dataset_len), int(seed),
                                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)), out.size))
    return out


def subset_indices(dataset_len: int, sample_size: Optional[int], random_sample: Optional[int]) -> List[int]:
    """The dataset indices a pass with these tally arguments visits, in order: everything (no sample_size), the first
    sample_size items (no seed), or the head of the seeded shuffle."""
    if sample_size is not None and sample_size > dataset_len:
        print("Warning: sample size %d > dataset size %d" % (sample_size, dataset_len))
        sample_size = dataset_len
    if sample_size is None:
        return list(range(dataset_len))
    if random_sample is None:
        return list(range(sample_size))
    return fixed_random_subset(dataset_len, sample_size, random_sample).tolist()


class FixedSubsetSampler(torch.utils.data.Sampler):
    """DataLoader sampler over a fixed index list."""

    def __init__(self, samples):
        self.samples = samples

    def __iter__(self):
        return iter(self.samples)

    def __len__(self):
        return len(self.samples)
