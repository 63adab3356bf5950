"""Host-side mirror of the reference's statistics runtime for the hot path.

Same object protocol, same npz cache layout as util/runningstats.py in SilentView/EMCID
(`SecondMoment` :469-511, `CombinedStat` :1347-1388, key prefixing :1391-1406, null boxing
:1417-1454, `load_cached_state` :1469-1493, `save_cached_state` :1496-1512, samplers :1515-1571),
but `SecondMoment.add` hands CUDA batches to the sm_100a kernels instead of `a.t().mm(a)`.

A stats file written here loads in the reference and vice versa:
    mom2.constructor  <U32   "util.runningstats.SecondMoment()"
    mom2.count        int64  ()
    mom2.mom2         float32 (d, d)   full symmetric raw sum (not divided by count)
    sample_size       int64  ()
"""
from __future__ import annotations

import os
import random
import struct
from typing import Dict, Optional

import numpy
import torch

# ---------------------------------------------------------------------------------------------
# Stat protocol
# ---------------------------------------------------------------------------------------------


class Stat:
    """Abstract base (reference: util/runningstats.py:144-232)."""

    def __init__(self, state=None):
        if state is not None:
            self.load_state_dict(resolve_state_dict(state))

    def add(self, x, *args, **kwargs):
        raise NotImplementedError

    def load_state_dict(self, d):
        raise NotImplementedError

    def state_dict(self):
        raise NotImplementedError

    def save(self, filename):
        save_cached_state(filename, self, {})

    def load(self, filename):
        self.load_state_dict(load_cached_state(filename, {}, quiet=True, throw=True))

    def to_(self, device):
        pass

    def cpu_(self):
        self.to_("cpu")

    def cuda_(self):
        self.to_("cuda")

    def _normalize_add_shape(self, x, attr="data_shape"):
        """Flatten to 2-D keeping the first dimension (reference :208-222)."""
        if not torch.is_tensor(x):
            x = torch.tensor(x)
        if len(x.shape) < 1:
            x = x.view(-1)
        data_shape = getattr(self, attr, None)
        if data_shape is None:
            data_shape = x.shape[1:]
            setattr(self, attr, data_shape)
        else:
            assert x.shape[1:] == data_shape
        return x.view(x.shape[0], int(numpy.prod(data_shape)))


class SecondMoment(Stat):
    """Uncentered second moment mom2 = sum_t a_t a_t^T and row count (reference :469-511).

    `add(a)` with a CUDA fp32 batch runs the tcgen05 SYRK (emcid_gemm3x_nt, lower tiles, stream-K)
    and mirrors lazily; any other input (CPU tensors, fp64) is rejected: the product has no CPU
    arithmetic path.  The fused fc1->act->mask->SYRK route that never materialises `a` is
    `emcid_b200.layer_stats` + `emcid_b200.mom2.Mom2Accumulator`; both produce this object.
    """

    CONSTRUCTOR = "util.runningstats.SecondMoment()"  # what the reference writes (:503-504)

    def __init__(self, split_batch=True, state=None):
        if state is not None:
            return super().__init__(state)
        self.count = 0
        self.mom2 = None
        self.split_batch = split_batch
        self._lower_dirty = False

    def add(self, a):
        a = self._normalize_add_shape(a)
        if len(a) == 0:
            return
        if not (a.is_cuda and a.dtype == torch.float32):
            raise RuntimeError(
                "emcid_b200.SecondMoment.add needs a CUDA float32 batch (sm_100a kernels only; "
                f"got {a.device}/{a.dtype})")
        from . import _lib

        if self.count == 0 or self.mom2 is None:
            self.mom2 = torch.zeros(a.shape[1], a.shape[1], dtype=torch.float32, device=a.device)
        elif self.mom2.device != a.device:
            self.mom2 = self.mom2.to(a.device)
        self._mirror()  # accumulate on a consistent full matrix
        a = a.contiguous() if a.stride(1) != 1 or a.stride(0) % 4 else a
        at = a.t().contiguous()  # [d, T]: both GEMM operands K-major over tokens
        self.count += a.shape[0]
        _lib.gemm3x_nt(at, at, self.mom2, alpha=1.0, beta=1.0, lower=True, streamk=True)
        self._lower_dirty = True

    def _mirror(self):
        if getattr(self, "_lower_dirty", False) and self.mom2 is not None:
            low = torch.tril(self.mom2)
            self.mom2 = low + torch.tril(self.mom2, -1).t()
            del low
            self._lower_dirty = False

    def to_(self, device):
        if self.mom2 is not None:
            self._mirror()
            self.mom2 = self.mom2.to(device)

    def moment(self):
        self._mirror()
        return self.mom2 / self.count

    def state_dict(self):
        self._mirror()
        return dict(constructor=self.CONSTRUCTOR, count=self.count, mom2=self.mom2.cpu().numpy())

    def load_state_dict(self, state):
        self.count = int(state["count"])
        self.mom2 = torch.from_numpy(numpy.asarray(state["mom2"]))
        self._lower_dirty = False


class Mean(Stat):
    """Running mean (reference :234-277); only reachable through to_collect=["mean"], never on the
    edit path.  Plain torch reductions."""

    CONSTRUCTOR = "util.runningstats.Mean()"

    def __init__(self, state=None):
        if state is not None:
            return super().__init__(state)
        self.count = 0
        self.batchcount = 0
        self._mean = None
        self.data_shape = None

    def add(self, a):
        a = self._normalize_add_shape(a)
        if len(a) == 0:
            return
        batch_count = a.shape[0]
        batch_mean = a.sum(0) / batch_count
        self.batchcount += 1
        if self._mean is None:
            self.count = batch_count
            self._mean = batch_mean
            return
        self.count += batch_count
        self._mean = self._mean + (batch_mean - self._mean) * (batch_count / self.count)

    def size(self):
        return self.count

    def mean(self):
        return self._mean.view(self.data_shape) if self.data_shape is not None else self._mean

    def to_(self, device):
        if self._mean is not None:
            self._mean = self._mean.to(device)

    def state_dict(self):
        return dict(constructor=self.CONSTRUCTOR, count=self.count,
                    data_shape=self.data_shape and tuple(self.data_shape), batchcount=self.batchcount,
                    mean=self._mean.cpu().numpy())

    def load_state_dict(self, state):
        self.count = int(state["count"])
        self.batchcount = int(state["batchcount"])
        self._mean = torch.from_numpy(numpy.asarray(state["mean"]))
        self.data_shape = None if state["data_shape"] is None else tuple(state["data_shape"])


class NormMean(Mean):
    """Mean of per-row L2 norms (reference :280-291)."""

    CONSTRUCTOR = "util.runningstats.NormMean()"

    def add(self, a):
        super().add(a.norm(dim=-1))


class CombinedStat(Stat):
    """Bundle of named stats with "name."-prefixed state keys (reference :1347-1388)."""

    def __init__(self, state=None, **kwargs):
        self._objs = kwargs
        if state is not None:
            return super().__init__(state)

    def __getattr__(self, k):
        if k != "_objs" and k in self.__dict__.get("_objs", {}):
            return self._objs[k]
        raise AttributeError(k)

    def add(self, d, *args, **kwargs):
        for obj in self._objs.values():
            obj.add(d, *args, **kwargs)

    def load_state_dict(self, state):
        for prefix, obj in self._objs.items():
            obj.load_state_dict(pull_key_prefix(prefix, state))

    def state_dict(self):
        result = {}
        for prefix, obj in self._objs.items():
            result.update(push_key_prefix(prefix, obj.state_dict()))
        return result

    def to_(self, device):
        for v in self._objs.values():
            v.to_(device)


def push_key_prefix(prefix, d):
    return {prefix + "." + k: v for k, v in d.items()}


def pull_key_prefix(prefix, d):
    pd = prefix + "."
    return {k[len(pd):]: v for k, v in d.items() if k.startswith(pd)}


# ---------------------------------------------------------------------------------------------
# npz cache: None is stored as the NaN with payload 0xfff8000000000002 (reference :1417-1454)
# ---------------------------------------------------------------------------------------------

_NULL_BITS = 0xFFF8000000000002
null_numpy_value = numpy.array(struct.unpack(">d", struct.pack(">Q", _NULL_BITS))[0], dtype=numpy.float64)


def is_null_numpy_value(v):
    return (isinstance(v, numpy.ndarray) and numpy.ndim(v) == 0 and v.dtype == numpy.float64
            and numpy.isnan(v) and struct.unpack(">Q", struct.pack(">d", v))[0] == _NULL_BITS)


def box_numpy_null(d):
    if isinstance(d, dict) or hasattr(d, "items"):
        return {k: box_numpy_null(v) for k, v in d.items()}
    return null_numpy_value if d is None else d


def unbox_numpy_null(d):
    if isinstance(d, dict) or hasattr(d, "items"):
        return {k: unbox_numpy_null(v) for k, v in d.items()}
    return None if is_null_numpy_value(d) else d


def resolve_state_dict(s):
    if isinstance(s, (str, os.PathLike)):
        return unbox_numpy_null(numpy.load(s))
    return s


global_load_cache_enabled = True


def load_cached_state(cachefile, args, quiet=False, throw=False):
    """Cache hit iff the file loads and every tally arg (sample_size) matches (reference :1469-1493)."""
    if not global_load_cache_enabled or cachefile is None:
        return None
    try:
        if isinstance(cachefile, dict):
            dat = cachefile
            cachefile = "state"
        else:
            dat = unbox_numpy_null(numpy.load(cachefile))
        for a, v in args.items():
            if a not in dat or dat[a] != v:
                if not quiet:
                    print("%s %s changed from %s to %s" % (cachefile, a, dat[a] if a in dat else None, v))
                return None
    except (FileNotFoundError, ValueError) as e:
        if throw:
            raise e
        return None
    else:
        if not quiet:
            print("Loading cached %s" % cachefile)
        return dat


def save_cached_state(cachefile, obj, args):
    """numpy.savez(cachefile, **state ∪ args), uncompressed, no pickle (reference :1496-1512)."""
    if cachefile is None:
        return
    dat = obj.state_dict()
    for a, v in args.items():
        if a in dat:
            assert dat[a] == v
        dat[a] = v
    if isinstance(cachefile, dict):
        cachefile.clear()
        cachefile.update(dat)
    else:
        os.makedirs(os.path.dirname(os.fspath(cachefile)) or ".", exist_ok=True)
        numpy.savez(cachefile, **box_numpy_null(dat))


# ---------------------------------------------------------------------------------------------
# fixed subsets (reference :1515-1571) — define WHICH captions a statistics pass visits
# ---------------------------------------------------------------------------------------------


class FixedSubsetSampler(torch.utils.data.sampler.Sampler):
    def __init__(self, samples):
        self.samples = samples

    def __iter__(self):
        return iter(self.samples)

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, key):
        return self.samples[key]


class FixedRandomSubsetSampler(FixedSubsetSampler):
    """random.Random(seed).shuffle(range(len(ds)))[start:end]."""

    def __init__(self, data_source, start=None, end=None, seed=1):
        rng = random.Random(seed)
        shuffled = list(range(len(data_source)))
        rng.shuffle(shuffled)
        self.data_source = data_source
        super().__init__(shuffled[start:end])


def subset_indices(dataset_len: int, sample_size: Optional[int], random_sample: Optional[int]):
    """Index list make_loader (reference :1574-1603) iterates for the given tally arguments."""
    if sample_size is not None and sample_size > dataset_len:
        print("Warning: sample size %d > dataset size %d" % (sample_size, dataset_len))
        sample_size = dataset_len
    if sample_size is None:
        return list(range(dataset_len))
    if random_sample is None:
        return list(range(sample_size))
    rng = random.Random(random_sample)
    shuffled = list(range(dataset_len))
    rng.shuffle(shuffled)
    return shuffled[:sample_size]


def tally(stat, dataset, cache=None, quiet=False, **kwargs):
    """Load-or-iterate convention of the reference (:54-121): if the cache file holds a matching
    state the stat is loaded and an empty iterator returned; otherwise a DataLoader over the fixed
    subset is returned and, once exhausted, the stat is moved to the CPU and saved."""
    assert isinstance(stat, Stat)
    args = {k: kwargs[k] for k in ["sample_size"] if k in kwargs}
    cached_state = load_cached_state(cache, args, quiet=quiet)
    if cached_state is not None:
        stat.load_state_dict(cached_state)
        return iter(())
    loader = make_loader(dataset, **kwargs)

    def wrapped_loader():
        yield from loader
        stat.to_(device="cpu")
        if cache is not None:
            save_cached_state(cache, stat, args)

    return wrapped_loader()


def make_loader(dataset, sample_size=None, batch_size=1, sampler=None, random_sample=None, **kwargs):
    if callable(dataset) and not isinstance(dataset, torch.utils.data.Dataset):
        dataset = dataset()
    if isinstance(dataset, torch.Tensor):
        dataset = torch.utils.data.TensorDataset(dataset)
    if sample_size is not None:
        assert sampler is None, "sampler cannot be specified with sample_size"
        sampler = FixedSubsetSampler(subset_indices(len(dataset), sample_size, random_sample))
    return torch.utils.data.DataLoader(dataset, sampler=sampler, batch_size=batch_size, **kwargs)
