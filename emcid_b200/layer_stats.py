"""Load-or-compute cached second-moment statistics of CLIP text-encoder MLP layers on B200.

Drop-in for emcid/layer_stats.py::layer_stats_text_encoder (:140-220) of SilentView/EMCID: same
signature, same cache file names and npz layout, same return type (a CombinedStat whose `.mom2`
holds the raw fp32 [d, d] sum on the CPU and the integer token count).  What changes is how the
numbers are produced:

  * ONE forward pass serves every requested layer (`layer_stats_text_encoder_multi`); the
    reference runs a full pass over the dataset per layer (emcid/layer_stats.py:112-134).
  * For a plain fp32 HF CLIP text tower the whole forward runs in the library (csrc/clip.cuh): packed
    valid tokens, 3xFP16 tcgen05 GEMMs for every projection, act(fc1) of the edited layers handed
    to the SYRK on the device.  The forward stops at the deepest edited layer, like `Trace(stop=True)`
    (util/nethook.py:112-113).
  * Any other module layout keeps the HF forward; a forward pre-hook on each edited layer's MLP then
    hands the h-wide LN2 output and the attention mask to `Mom2Accumulator`, whose kernels do
    fc1 -> activation -> pad masking -> lower-triangular SYRK (csrc/mom2.cuh).  Either way the
    d-wide fc2 input of the reference's `Trace(retain_input=True)` is never gathered by PyTorch.
  * Captions are sharded `subset[rank::world]` across the GPUs of a torch.distributed job; the per-rank
    sums meet in one reduction per layer inside the library (emcid_mom2_reduce: lower triangle, NCCL reduce
    to the rank that writes the layer's npz); counts are reduced as int64 (bit exact).
  * A pass can checkpoint its accumulators every N device blocks and resume after a crash
    (`checkpoint_every`): the reference saves only once the loader is exhausted (util/runningstats.py:115-119).
  * `layer_stats_cross_attn_kv` (emcid/layer_stats.py:333-427): statistics of the UNet cross-attention K/V inputs,
    i.e. of the text encoder's last_hidden_state — one pass for all K/V modules.

There is no CPU path: a model that is not on a CUDA (sm_100) device is an error.
"""
from __future__ import annotations

import os
import threading
import time
import zlib
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import clip_forward, nethook
from .runningstats import (CombinedStat, FixedSubsetSampler, SecondMoment, load_cached_state, save_cached_state,
                           subset_indices)
from .stat_dataset import (DEFAULT_BLOCK_TOKENS, PackedReblocker, TokenizedDataset, dict_to_, fixed_width_collation,
                           packed_collation, unpack_to_padded)

try:  # the reference's default progress bar
    from tqdm.auto import tqdm
except Exception:  # pragma: no cover
    tqdm = None

# statistics computed here; the reference's other two (mean, norm_mean; emcid/layer_stats.py:26-30) are delegated
STAT_TYPES = {"mom2": SecondMoment}
REFERENCE_ONLY_STATS = ("mean", "norm_mean")

# what the last computed pass did (tests and bench.py read it)
LAST_PASS_INFO: Dict[str, object] = {}

HEAD_CAPTIONS = 4096                # captions a pass collates in-process while its loader workers start (see _run_pass)
LAST_HIDDEN = "last_hidden_state"   # pseudo layer name: the text encoder's output (input of the UNet cross-attention K/V)


def get_ccs_filtered_ds(tokenizer):
    """Dataset factory (reference :137-138); the patch point for synthetic data."""
    return TokenizedDataset("./data/ccs_filtered.json", tokenizer)


def stats_filename(stats_dir, model_name, ds_name, layer_name, precision, to_collect, batch_tokens, sample_size):
    """Cache path, reference :166-174."""
    size_suffix = "" if sample_size is None else f"_{sample_size}"
    size_suffix = f"_t{batch_tokens}" + size_suffix
    return Path(stats_dir) / (f"{model_name}/{ds_name}_stats/{layer_name}_{precision}_"
                              f"{'-'.join(sorted(to_collect))}{size_suffix}.npz")


def _mlp_of(model, layer_name: str):
    """The CLIP MLP that owns `layer_name` (= "...mlp.fc2"): its argument is the fc1 input."""
    if not layer_name.endswith(".fc2"):
        raise NotImplementedError(
            f"emcid_b200 computes mom2 for CLIP MLP fc2 inputs (…mlp.fc2); got {layer_name!r}")
    nethook.get_module(model, layer_name)  # LookupError for unknown names, like the reference
    mlp = nethook.get_module(model, layer_name[: -len(".fc2")])
    if not hasattr(mlp, "fc1") or not isinstance(mlp.fc1, torch.nn.Linear):
        raise NotImplementedError(f"{layer_name}: parent module has no fc1 Linear")
    return mlp


class TextEncoderMom2Pass:
    """Streams caption blocks through the CLIP text model and accumulates mom2/count for a set of MLP layers — or, with
    `layer_names == [LAST_HIDDEN]`, for the encoder's output — in the same pass."""

    def __init__(self, model, layer_names: Sequence[str], slab_tokens: int = 0, accumulator_factory=None,
                 native: Optional[bool] = None):
        from .mom2 import Mom2Accumulator

        if native is None:
            native = os.environ.get("EMCID_NATIVE_FORWARD", "1") != "0"
        self._use_native = bool(native) and accumulator_factory is None and clip_forward.supports(model)
        self._native = None
        self.model = model
        self.layer_names = list(layer_names)
        self.final_hidden = self.layer_names == [LAST_HIDDEN]
        if LAST_HIDDEN in self.layer_names and not self.final_hidden:
            raise ValueError("last_hidden_state statistics run in a pass of their own")
        device = next(model.parameters()).device
        cfg = clip_forward.text_config(model)
        self.accs: Dict[str, object] = {}
        self._mlps = []
        self._mask = None
        self._hook_weights_set = False
        self._layer_index = {}
        self.capacity_hint = 0     # tokens per block the caller is going to send (sizes the native encoder once)
        self.fallback_blocks = 0   # blocks that went through the HF forward although the native one was available
        if self.final_hidden:
            if device.type != "cuda":
                raise RuntimeError("emcid_b200.layer_stats needs the model on a CUDA sm_100 device (no CPU fallback); "
                                   f"model is on {device}")
            h = cfg.hidden_size
            if self._use_native:
                # d == hidden: plain SYRK of the last_hidden_state planes (no fc1, no activation)
                self.accs[LAST_HIDDEN] = Mom2Accumulator(device, h, h, "none", slab_tokens=slab_tokens)
            else:
                self.accs[LAST_HIDDEN] = _SecondMomentAccumulator(device, h)
            return
        mlps = {name: _mlp_of(model, name) for name in self.layer_names}  # LookupError before anything else
        if accumulator_factory is None:
            if device.type != "cuda":
                raise RuntimeError(
                    "emcid_b200.layer_stats needs the model on a CUDA sm_100 device (no CPU fallback); "
                    f"model is on {device}")
            accumulator_factory = lambda d, h, act: Mom2Accumulator(device, d, h, act, slab_tokens=slab_tokens)
        act = getattr(cfg, "hidden_act", None)
        if act not in ("quick_gelu", "gelu"):
            raise NotImplementedError(f"unsupported CLIP MLP activation {act!r}")
        order = {n: i for i, (n, _) in enumerate(model.named_modules())}
        deepest = max(self.layer_names, key=lambda n: order[n])
        for name in self.layer_names:
            mlp = mlps[name]
            W1 = mlp.fc1.weight
            if W1.dtype != torch.float32:
                raise NotImplementedError("the statistics pass runs on fp32 weights (reference precision float32)")
            self.accs[name] = accumulator_factory(W1.shape[0], W1.shape[1], act)
            self._mlps.append((name, mlp, name == deepest))
        if self._use_native:
            tm = getattr(model, "text_model", model)
            for name, mlp, _ in self._mlps:
                idx = [i for i, ly in enumerate(tm.encoder.layers) if ly.mlp is mlp]
                if len(idx) != 1:
                    self._use_native = False
                    break
                self._layer_index[name] = idx[0]
        if not self._use_native:
            self._ensure_hook_weights()

    def _ensure_hook_weights(self):
        if not self._hook_weights_set:
            for name, mlp, _ in self._mlps:
                self.accs[name].set_weights(mlp.fc1.weight, mlp.fc1.bias)
            self._hook_weights_set = True

    def _native_encoder(self, n_tokens: int, n_captions: int):
        nat = self._native
        if nat is None or nat.max_tokens < n_tokens or nat.max_captions < n_captions:
            if nat is not None:
                torch.cuda.synchronize(nat.device)
                nat.close()
            cap_tokens = max(n_tokens, nat.max_tokens if nat else 0, int(self.capacity_hint))
            # captions are at least one token long: a caption capacity of cap_tokens never has to grow (it sizes nothing)
            self._native = nat = clip_forward.NativeClipTextEncoder(self.model, cap_tokens, max(n_captions, cap_tokens))
        return nat

    def _hook(self, name, stop):
        def pre_hook(_module, args):
            self.accs[name].add(args[0], self._mask)
            if stop:
                raise nethook.StopForward()
        return pre_hook

    def _max_positions(self) -> int:
        return getattr(self.model, "text_model", self.model).embeddings.position_embedding.weight.shape[0]

    def _run_native(self, ids, pos, cu, S, T):
        nat = self._native_encoder(T, S)
        if self.final_hidden:
            nat.forward_final(ids, pos, cu, S, T, acc=self.accs[LAST_HIDDEN])
        else:
            names = list(self._layer_index)
            nat.forward_stats(ids, pos, cu, S, T, [self._layer_index[n] for n in names], [self.accs[n] for n in names])

    @torch.no_grad()
    def run_batch(self, batch: Dict[str, torch.Tensor]) -> None:
        """batch: input_ids / position_ids / attention_mask [B, L] (host or device tensors), or an already
        packed block (packed_ids / packed_pos / cu_seqlens, see stat_dataset.packed_collation)."""
        if "packed_ids" in batch:
            cu = batch["cu_seqlens"]
            S, T = cu.numel() - 1, batch["packed_ids"].numel()
            if T == 0:
                return
            if self._use_native and (cu.is_cuda or int((cu[1:] - cu[:-1]).max()) <= self._max_positions()):
                self._run_native(batch["packed_ids"], batch["packed_pos"], cu, S, T)
                return
            batch = unpack_to_padded(batch)
        if self._use_native:
            packed = clip_forward.pack_batch(batch, self._max_positions())
            if packed is not None:
                ids, pos, cu, S, T = packed
                if T:
                    self._run_native(ids, pos, cu, S, T)
                return
            self.fallback_blocks += 1   # not a right-padding mask: keep the HF forward for this block
            if not self.final_hidden:
                self._ensure_hook_weights()
        device = next(self.model.parameters()).device
        batch = {k: v.to(device, non_blocking=True) for k, v in batch.items()}
        if self.final_hidden:
            # reference emcid/layer_stats.py:415-426: feats = flatten_masked_batch(last_hidden_state, mask)
            tm = getattr(self.model, "text_model", self.model)
            hidden = tm(**batch)[0] if tm is not self.model else self.model(**batch)[0]
            self.accs[LAST_HIDDEN].add(hidden, batch["attention_mask"])
            return
        self._mask = batch["attention_mask"]
        handles = [mlp.register_forward_pre_hook(self._hook(name, stop)) for name, mlp, stop in self._mlps]
        try:
            tm = getattr(self.model, "text_model", None)
            (tm if type(self.model).__name__ == "CLIPModel" and tm is not None else self.model)(**batch)
        except nethook.StopForward:
            pass
        finally:
            for h in handles:
                h.remove()
            self._mask = None

    def finalize(self):
        """{layer_name: (mom2 [d,d] fp32 on device, count 0-d int64 on device)}."""
        return {name: acc.finalize() for name, acc in self.accs.items()}

    def launches(self) -> int:
        """Kernels of this library launched so far by the pass (accumulators + native forward)."""
        n = sum(int(acc.get_profile()["launches"]) for acc in self.accs.values() if hasattr(acc, "get_profile"))
        return n + (self._native.launches() if self._native is not None else 0)

    def close(self):
        for acc in self.accs.values():
            acc.close()
        if self._native is not None:
            self._native.close()
            self._native = None


class _SecondMomentAccumulator:
    """Accumulator interface over runningstats.SecondMoment (emcid_gemm3x_nt lower GEMM) for feature rows that already
    exist on the device: the last_hidden_state fallback when the native forward does not cover the model."""

    def __init__(self, device, d):
        self.device, self.d = device, d
        self.stat = SecondMoment()

    def add(self, feats, mask=None):
        flat = feats.reshape(-1, feats.shape[-1]).float()
        if mask is not None:
            flat = flat[mask.reshape(-1).nonzero()[:, 0]]
        self.stat.add(flat.contiguous())

    def finalize(self):
        m = self.stat.mom2
        if m is None:
            m = torch.zeros(self.d, self.d, dtype=torch.float32, device=self.device)
        return m, torch.tensor(self.stat.count, dtype=torch.int64, device=self.device)

    def close(self):
        pass


def _dist_info(distributed):
    import torch.distributed as dist

    if distributed is False or not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(), dist.get_world_size()


def _nccl_comm(dist, device) -> Optional[int]:
    """Address of the ncclComm_t torch's default process group uses on `device`, or None when the job does not run on
    NCCL (gloo host-logic tests) or this torch build does not expose it."""
    try:
        if dist.get_backend() != "nccl":
            return None
        pg = dist.distributed_c10d._get_default_group()
        backend = pg._get_backend(torch.device(device))
        if not hasattr(backend, "_comm_ptr"):
            return None
        try:
            ptr = int(backend._comm_ptr())
        except Exception:
            ptr = 0
        if not ptr:    # communicators are created lazily: a first collective brings this one up
            dist.all_reduce(torch.zeros(1, device=device))
            ptr = int(backend._comm_ptr())
        return ptr or None
    except Exception:
        return None


def _pinned_copy(t: torch.Tensor) -> torch.Tensor:
    """Asynchronous D2H into pinned memory; the caller synchronises the stream once for all copies."""
    if not t.is_cuda:
        return t
    host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host


def _reduce_results(runner, todo, dist, rank, world, device, broadcast, keep_on_device):
    """The exchange step + hand-over: {layer: (mom2 or None, count)}; mom2 is present on the layer's root rank (layer i
    -> rank i mod world) and, with `broadcast`, everywhere.  CPU tensors unless keep_on_device."""
    from . import _lib

    roots = {n: i % world for i, n in enumerate(todo)}
    out: Dict[str, tuple] = {}
    if dist is None or world == 1:
        fin = runner.finalize()
        for n in todo:
            mom2, count = fin[n]
            out[n] = [mom2 if keep_on_device else _pinned_copy(mom2), count]
        if device.type == "cuda":
            torch.cuda.current_stream(device).synchronize()
        return {n: (m, int(c.item())) for n, (m, c) in out.items()}, roots
    comm = _nccl_comm(dist, device) if all(hasattr(runner.accs[n], "reduce") for n in todo) else None
    LAST_PASS_INFO["exchange"] = "emcid_mom2_reduce (ncclReduce, lower-packed fp32)" if comm else "torch.distributed"
    counts = torch.zeros(len(todo), dtype=torch.int64, device=device)
    mats: Dict[str, Optional[torch.Tensor]] = {}
    if comm:
        for n in todo:
            runner.accs[n].reduce(comm, roots[n])
        for i, n in enumerate(todo):
            if rank == roots[n]:
                mats[n], c = runner.accs[n].finalize()
                counts[i] = c
            else:
                mats[n] = None
        dist.all_reduce(counts)                 # roots contribute the reduced counts, everybody else zeros
        if broadcast:
            lib = _lib.lib()
            for n in todo:
                d = runner.accs[n].d
                if mats[n] is None:
                    mats[n] = torch.empty(d, d, dtype=torch.float32, device=device)
                with torch.cuda.device(device):
                    _lib.check(lib.emcid_mom2_broadcast(_lib.ptr(mats[n]), None, d, comm, roots[n], _lib.current_stream_ptr()))
    else:   # host-logic path (gloo, test accumulators): the same exchange through torch.distributed
        fin = runner.finalize()
        for i, n in enumerate(todo):
            mom2, c = fin[n]
            if broadcast:
                dist.all_reduce(mom2)
            else:
                dist.reduce(mom2, dst=roots[n])
            mats[n] = mom2 if (broadcast or rank == roots[n]) else None
            counts[i] = c
        dist.all_reduce(counts)
    for n in todo:
        m = mats[n]
        out[n] = m if (m is None or keep_on_device or not m.is_cuda) else _pinned_copy(m)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()
    counts = counts.cpu().tolist()
    return {n: (out[n], int(counts[i])) for i, n in enumerate(todo)}, roots


# ---------------------------------------------------------------------------------------------------------
# resumable pass: checkpoint of the accumulators beside the npz files
# ---------------------------------------------------------------------------------------------------------
class _Checkpointer:
    """Every `every` device blocks: fold + export the packed fp64 lower triangles and counts of all layers
    (emcid_mom2_export_state), copy them to pinned memory and write them — in a background thread — to ONE npz per rank
    next to the statistics files, together with the cursor (captions of this rank's shard consumed so far).  A pass that
    finds a matching file starts from it.  Checkpoints cut the pass at block boundaries, and blocks are a pure function
    of the caption order (stat_dataset.PackedReblocker), so a resumed pass sees the same blocks and folds at the same
    points as an uninterrupted pass with the same `every`."""

    def __init__(self, path: Path, meta: Dict[str, object], every: int, device=None):
        self.path, self.meta, self.every, self.device = Path(path), dict(meta), int(every), device
        self._thread: Optional[threading.Thread] = None
        self._host: Dict[str, torch.Tensor] = {}
        self.blocks = 0
        self.written = 0

    @staticmethod
    def file_for(stats_dir, model_name, ds_name, precision, batch_tokens, sample_size, layer_names, rank, world) -> Path:
        tag = zlib.crc32(",".join(layer_names).encode()) & 0xFFFFFFFF
        return Path(stats_dir) / (f"{model_name}/{ds_name}_stats/.resume_{precision}_t{batch_tokens}_{sample_size}_"
                                  f"{tag:08x}_r{rank}of{world}.npz")

    def load(self, runner, names) -> int:
        """Captions already consumed (0 = start from scratch); imports the accumulator states when the file matches."""
        try:
            with np.load(self.path) as dat:
                for k, v in self.meta.items():
                    if k not in dat.files or str(dat[k]) != str(v):
                        return 0
                done = int(dat["captions_done"])
                states = [(dat[f"state.{i}"], int(dat[f"count.{i}"])) for i in range(len(names))]
        except (FileNotFoundError, OSError, ValueError, KeyError, EOFError):
            return 0
        for n, (packed, count) in zip(names, states):
            runner.accs[n].import_state(torch.from_numpy(packed), count)
        return done

    def after_block(self, runner, names, captions_done: int) -> None:
        self.blocks += 1
        if self.every <= 0 or self.blocks % self.every:
            return
        self.wait()                                    # the pinned buffers are free again
        payload = {}
        for i, n in enumerate(names):
            packed, count = runner.accs[n].export_state()
            if packed.is_cuda:
                host = self._host.get(n)
                if host is None:
                    host = self._host[n] = torch.empty(packed.shape, dtype=packed.dtype, device="cpu", pin_memory=True)
                host.copy_(packed, non_blocking=True)
                packed = host
            payload[f"state.{i}"] = packed
            payload[f"count.{i}"] = count
        if self.device is not None and torch.device(self.device).type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        payload = {k: (v.numpy() if k.startswith("state.") else np.int64(v.item())) for k, v in payload.items()}
        payload.update(self.meta, captions_done=np.int64(captions_done))
        self._thread = threading.Thread(target=self._write, args=(payload,), daemon=True)
        self._thread.start()

    def _write(self, payload) -> None:
        self.path.parent.mkdir(parents=True, exist_ok=True)
        tmp = f"{self.path}.{os.getpid()}.tmp"
        with open(tmp, "wb") as f:              # a file object: numpy must not append ".npz" to the temporary name
            np.savez(f, **payload)
        os.replace(tmp, self.path)
        self.written += 1

    def _sweep_temporaries(self) -> None:
        """Half-written files of workers that died inside _write."""
        for stale in self.path.parent.glob(self.path.name + ".*.tmp"):
            try:
                os.remove(stale)
            except OSError:
                pass

    def wait(self) -> None:
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def discard(self) -> None:
        self.wait()
        self._sweep_temporaries()
        try:
            os.remove(self.path)
        except FileNotFoundError:
            pass


def _delegate_to_reference(fn_name: str, what: str, *args, **kwargs):
    """Requests outside the accelerated path (mean / norm_mean statistics, float64 accumulation) go to the reference's
    own function when SilentView/EMCID is importable in this process, unchanged; otherwise they are refused."""
    try:
        import emcid.layer_stats as ref_ls       # the reference package, if it is installed next to this one
    except Exception:
        raise NotImplementedError(
            f"{what} is not on the B200 path (every EMCID edit passes to_collect=['mom2'], precision='float32': "
            "emcid/emcid_main.py:2192,2227,2267) and the reference package `emcid` is not importable to delegate to") from None
    return getattr(ref_ls, fn_name)(*args, **kwargs)


def _run_pass(model, tokenizer, todo: List[str], files, stats, args, *, sample_size, device, progress, force_recompute,
              captions_per_batch, num_workers, slab_tokens, distributed, keep_on_device, block_tokens, broadcast,
              checkpoint, accumulator_factory, write_also=None):
    """One pass over this rank's caption shard for the layers in `todo` (all MLP layers, or [LAST_HIDDEN])."""
    t_start = time.perf_counter()
    dist, rank, world = _dist_info(distributed)
    ds = get_ccs_filtered_ds(tokenizer=tokenizer)
    indices = subset_indices(len(ds), sample_size, random_sample=1)  # tally(..., random_sample=1), reference :204
    my_indices = indices[rank::world]

    runner = TextEncoderMom2Pass(model, todo, slab_tokens=slab_tokens, accumulator_factory=accumulator_factory)
    ckpt, done = None, 0
    if checkpoint is not None and all(hasattr(a, "export_state") for a in runner.accs.values()):
        ckpt_file, every, meta = checkpoint
        meta = dict(meta, dataset_len=len(ds), shard_len=len(my_indices), rank=rank, world=world,
                    block_tokens=int(block_tokens or 0), captions_per_batch=int(captions_per_batch), layers=",".join(todo))
        ckpt = _Checkpointer(ckpt_file(rank, world), meta, every, device)
        done = ckpt.load(runner, todo)
        LAST_PASS_INFO["resumed_from_caption"] = done
    # The first HEAD_CAPTIONS captions of the shard are collated right here, in order, while the loader's worker
    # processes are forked and collate their first batches; the workers serve the rest.  Caption order — and with it the
    # checkpoint cursor — is what it was.
    rest = my_indices[done:]
    head_n = min(len(rest), HEAD_CAPTIONS) if num_workers > 0 else 0
    collate = packed_collation() if (device.type == "cuda" and accumulator_factory is None) else fixed_width_collation()
    loader = torch.utils.data.DataLoader(
        ds, sampler=FixedSubsetSampler(rest[head_n:]), batch_size=captions_per_batch, collate_fn=collate,
        num_workers=num_workers if len(rest) > head_n else 0,      # nothing left for workers: do not fork any
        pin_memory=(device.type == "cuda"))
    batch_count = -(-head_n // captions_per_batch) + -(-(len(rest) - head_n) // captions_per_batch)
    if progress is None:
        progress = lambda x, total=None: x
    reblock = PackedReblocker(block_tokens) if (block_tokens and runner._use_native) else None
    if reblock is not None:
        runner.capacity_hint = min(int(block_tokens), len(my_indices) * int(runner._max_positions()))
    t_loop = time.perf_counter()
    consumed = done
    # Start-up order: the weights go to the library first, the first in-process batches follow (the device has work from
    # here on), and only then are the loader's worker processes forked — two forks of a process with a CUDA context are
    # 60 ms during which nothing else happens on the host; behind the first three device blocks they are free.  (Forked
    # before anything else, start-up was 0.1 s of every pass: a sixth of configs[1]'s 0.6 s per rank at 8 GPUs.)
    if reblock is not None and runner.capacity_hint > 0 and batch_count > 0:
        runner._native_encoder(int(runner.capacity_hint), 1)
    fork_after = 6 * captions_per_batch

    def all_batches():
        tail_batches = None
        for a in range(0, head_n, captions_per_batch):
            if tail_batches is None and a >= fork_after:
                tail_batches = iter(loader)
            yield collate([ds[i] for i in rest[a: min(a + captions_per_batch, head_n)]])
        yield from (tail_batches if tail_batches is not None else iter(loader))

    def feed(block):
        nonlocal consumed
        runner.run_batch(block)
        consumed += (block["cu_seqlens"].numel() - 1) if "cu_seqlens" in block else block["input_ids"].shape[0]
        if ckpt is not None:
            ckpt.after_block(runner, todo, consumed)

    try:
        t_wait = t_run = 0.0
        t_prev = time.perf_counter()
        for batch in progress(all_batches(), total=batch_count):
            t_got = time.perf_counter()
            t_wait += t_got - t_prev
            if ("packed_ids" in batch and batch["packed_ids"].numel() == 0) or \
                    ("input_ids" in batch and batch["input_ids"].numel() == 0):
                t_prev = time.perf_counter()
                continue
            if reblock is not None and "packed_ids" in batch:
                # loader batches count captions, device blocks count tokens (stat_dataset.PackedReblocker)
                for block in reblock.push(batch):
                    feed(block)
            else:
                feed(batch)  # host tensors: packed on the host, then one pinned H2D copy per field
            t_prev = time.perf_counter()
            t_run += t_prev - t_got
        if reblock is not None:
            for block in reblock.flush():
                feed(block)
        t_fin = time.perf_counter()
        results, roots = _reduce_results(runner, todo, dist, rank, world, device, broadcast, keep_on_device)
        LAST_PASS_INFO.update(native_forward=runner._native is not None, launches=runner.launches(),
                              fallback_blocks=runner.fallback_blocks, captions=len(my_indices))
        to_save = []
        for n in todo:
            mom2, count = results[n]
            for name in ([n] if write_also is None else write_also):
                sm = stats[name].mom2
                sm.count = count
                sm.mom2 = mom2
                if not force_recompute and rank == roots[n]:
                    to_save.append(name)
        # the statistics files of this rank's layers, written side by side: one 37.7 MB npz is 60 ms of CRC-32 and file
        # write (both outside the interpreter lock), five of them in a row 0.3 s behind a 4.6 s pass
        if len(to_save) > 1:
            from concurrent.futures import ThreadPoolExecutor

            with ThreadPoolExecutor(max_workers=min(len(to_save), 8)) as pool:
                list(pool.map(lambda name: save_cached_state(files[name], stats[name], args), to_save))
        else:
            for name in to_save:
                save_cached_state(files[name], stats[name], args)
        if ckpt is not None:
            ckpt.discard()                       # the pass is complete: its statistics files are the durable result
        if dist is not None and world > 1:
            dist.barrier()
        # host-side timeline of the pass (no extra synchronisation: the loop time includes whatever the host waited for)
        t_end = time.perf_counter()
        LAST_PASS_INFO["timing"] = {"setup_s": t_loop - t_start, "loop_s": t_fin - t_loop, "finalize_s": t_end - t_fin,
                                    "loader_wait_s": t_wait, "run_batch_s": t_run}
    finally:
        if ckpt is not None:
            ckpt.wait()
        runner.close()


def layer_stats_text_encoder_multi(
    model,
    tokenizer,
    layer_names: Sequence[str],
    stats_dir="data/stats",
    ds_name="ccs_filtered",
    to_collect=["mom2"],
    model_name="text_encoder",
    sample_size=None,
    precision=None,
    batch_tokens=3 * 1024,
    download=False,
    progress=tqdm,
    force_recompute=False,
    captions_per_batch: int = 256,
    num_workers: int = 2,
    slab_tokens: int = 0,
    distributed: Optional[bool] = None,
    keep_on_device: bool = False,
    block_tokens: int = DEFAULT_BLOCK_TOKENS,
    broadcast: bool = False,
    checkpoint_every: Optional[int] = None,
    _accumulator_factory: Optional[Callable] = None,
) -> Dict[str, CombinedStat]:
    """All `layer_names` in one pass.  Arguments up to `force_recompute` mean exactly what they mean
    in the reference's layer_stats_text_encoder; the rest tune the B200 driver:

    broadcast         under torch.distributed the reduced matrix of layer i lives on rank i mod world (which writes its
                      npz); False leaves `.mom2.mom2 = None` on the other ranks (counts are set everywhere), True sends
                      every matrix to every rank (what the single-layer reference-shaped call does).
    checkpoint_every  device blocks between checkpoints of the accumulators (0 = never; default: env
                      EMCID_STATS_CHECKPOINT_BLOCKS or 0).  A pass killed after a checkpoint resumes from it."""
    device = model.device
    if precision is None:
        precision = "float64"  # reference default (:161-162); every EMCID caller passes "float32"
    if precision != "float32":
        raise NotImplementedError(
            f"precision={precision!r}: emcid_b200 accumulates mom2 with fp32-class products (hparams.mom2_dtype == 'float32', "
            "the only precision an EMCID edit asks for); pass precision='float32'")
    if sorted(to_collect) != ["mom2"]:
        raise NotImplementedError(
            f"only to_collect=['mom2'] is on the accelerated path (got {to_collect}); "
            "mean / norm_mean are not used by the EMCID edit")
    layer_names = list(layer_names)
    stats_dir = Path(stats_dir)
    stats_dir.mkdir(exist_ok=True, parents=True)
    files = {n: stats_filename(stats_dir, model_name, ds_name, n, precision, to_collect, batch_tokens, sample_size)
             for n in layer_names}
    for f in files.values():
        if not f.exists() and download:
            raise NotImplementedError("Downloading stats from remote is not implemented yet.")  # reference :176-178

    args = {"sample_size": sample_size}  # tally forwards it even when None (boxed as a null NaN in the npz)
    stats: Dict[str, CombinedStat] = {}
    todo: List[str] = []
    for n in layer_names:
        stat = CombinedStat(**{k: STAT_TYPES[k]() for k in to_collect})
        cached = None if force_recompute else load_cached_state(files[n], args)
        if cached is not None:
            stat.load_state_dict(cached)
            if keep_on_device:
                stat.to_(device)
        else:
            todo.append(n)
        stats[n] = stat
    if not todo:
        return stats
    if checkpoint_every is None:
        checkpoint_every = int(os.environ.get("EMCID_STATS_CHECKPOINT_BLOCKS", "0"))
    checkpoint = None
    if checkpoint_every > 0 and not force_recompute:
        checkpoint = (lambda rank, world: _Checkpointer.file_for(stats_dir, model_name, ds_name, precision, batch_tokens,
                                                                 sample_size, todo, rank, world),
                      checkpoint_every, {"sample_size": str(sample_size), "seed": 1})
    _run_pass(model, tokenizer, todo, files, stats, args, sample_size=sample_size, device=device, progress=progress,
              force_recompute=force_recompute, captions_per_batch=captions_per_batch, num_workers=num_workers,
              slab_tokens=slab_tokens, distributed=distributed, keep_on_device=keep_on_device, block_tokens=block_tokens,
              broadcast=broadcast, checkpoint=checkpoint, accumulator_factory=_accumulator_factory)
    return stats


def layer_stats_text_encoder(
    model,
    tokenizer,
    layer_name,
    stats_dir="data/stats",
    ds_name="ccs_filtered",
    to_collect=["mom2"],
    model_name="text_encoder",
    sample_size=None,
    precision=None,
    batch_tokens=3 * 1024,
    download=False,
    progress=tqdm,
    force_recompute=False,
    **b200_options,
) -> CombinedStat:
    """Function to load or compute cached stats (signature of emcid/layer_stats.py:140-154).  Every caller gets the
    statistics (under torch.distributed: broadcast from the rank that reduced them)."""
    if any(k in REFERENCE_ONLY_STATS for k in to_collect) or (precision or "float64") != "float32":
        # STAT_TYPES mean / norm_mean and float64 accumulation (:26-30, :161-162) are not on the edit path
        return _delegate_to_reference(
            "layer_stats_text_encoder", f"to_collect={to_collect}, precision={precision!r}", model, tokenizer, layer_name,
            stats_dir=stats_dir, ds_name=ds_name, to_collect=to_collect, model_name=model_name, sample_size=sample_size,
            precision=precision, batch_tokens=batch_tokens, download=download, progress=progress,
            force_recompute=force_recompute)
    b200_options.setdefault("broadcast", True)
    return layer_stats_text_encoder_multi(
        model, tokenizer, [layer_name], stats_dir=stats_dir, ds_name=ds_name, to_collect=to_collect,
        model_name=model_name, sample_size=sample_size, precision=precision, batch_tokens=batch_tokens,
        download=download, progress=progress, force_recompute=force_recompute, **b200_options)[layer_name]


# ---------------------------------------------------------------------------------------------------------
# UNet cross-attention K/V modules: statistics of the text encoder's output
# ---------------------------------------------------------------------------------------------------------
CROSS_ATTN_KV_TEMPLATES = {   # util/globals.py:37-38 (UNET_EDIT_TEMPLATES["cross-k" / "cross-v"])
    "cross-k": "{}.{}.attentions.{}.transformer_blocks.0.attn2.to_k",
    "cross-v": "{}.{}.attentions.{}.transformer_blocks.0.attn2.to_v",
}


def get_all_cross_attn_kv_layer_names(pipe) -> List[str]:
    """Names of the attn2.to_k / attn2.to_v projections present in pipe.unet, in the reference's order
    (emcid/layer_stats.py:470-495: block types, block index, template, attention index; the mid block has no block index)."""
    names = []
    for block_type, count in (("down_blocks", 4), ("up_blocks", 4), ("mid_block", 1)):
        for idx in range(count):
            for template in CROSS_ATTN_KV_TEMPLATES.values():
                for sub_idx in (0, 1, 2):
                    name = template.format(block_type, idx, sub_idx)
                    if block_type == "mid_block":
                        name = name.replace(f"mid_block.{idx}.", "mid_block.")
                    obj = pipe.unet
                    try:
                        for part in name.split("."):
                            obj = getattr(obj, part)
                    except AttributeError:
                        continue
                    names.append(name)
    return names


def layer_stats_cross_attn_kv(
    pipe,
    layer_name,
    stats_dir="data/stats",
    ds_name="ccs_filtered",
    to_collect=["mom2"],
    model_name="unet",
    sample_size=None,
    precision=None,
    batch_tokens=3 * 1024,
    download=False,
    progress=tqdm,
    force_recompute=False,
    share_with: Optional[Sequence[str]] = None,
    **b200_options,
) -> CombinedStat:
    """Statistics of the input of one UNet cross-attention K/V projection (signature of emcid/layer_stats.py:333-346).
    That input is the text encoder's last_hidden_state whatever the module, so the pass runs the text encoder only
    (the reference pushes dummy latents through the UNet up to the traced module, :397-424) and — `share_with`, default:
    every K/V module of pipe.unet — its result is written under all those modules' file names at once, which turns the
    32 passes of compute_cross_attn_kv_stats (:429-467) into one."""
    if any(k in REFERENCE_ONLY_STATS for k in to_collect) or (precision or "float64") != "float32":
        return _delegate_to_reference(
            "layer_stats_cross_attn_kv", f"to_collect={to_collect}, precision={precision!r}", pipe, layer_name,
            stats_dir=stats_dir, ds_name=ds_name, to_collect=to_collect, model_name=model_name, sample_size=sample_size,
            precision=precision, batch_tokens=batch_tokens, download=download, progress=progress,
            force_recompute=force_recompute)
    model = pipe.text_encoder
    device = model.device
    stats_dir = Path(stats_dir)
    stats_dir.mkdir(exist_ok=True, parents=True)
    nethook.get_module(pipe.unet, layer_name)                    # LookupError for unknown names
    if share_with is None:
        share_with = get_all_cross_attn_kv_layer_names(pipe) if hasattr(pipe, "unet") else []
    names = [layer_name] + [n for n in share_with if n != layer_name]
    files = {n: stats_filename(stats_dir, model_name, ds_name, n, precision, to_collect, batch_tokens, sample_size)
             for n in names}
    print(files[layer_name])                                      # reference :362
    if not files[layer_name].exists() and download:
        raise NotImplementedError("Downloading stats from remote is not implemented yet.")
    args = {"sample_size": sample_size}
    stats = {n: CombinedStat(**{k: STAT_TYPES[k]() for k in to_collect}) for n in names}
    cached = None if force_recompute else load_cached_state(files[layer_name], args)
    if cached is not None:
        stats[layer_name].load_state_dict(cached)
        return stats[layer_name]
    b200_options.setdefault("broadcast", True)
    opts = dict(captions_per_batch=256, num_workers=2, slab_tokens=0, distributed=None, keep_on_device=False,
                block_tokens=DEFAULT_BLOCK_TOKENS, broadcast=True)
    opts.update(b200_options)
    # the pass keys its result by LAST_HIDDEN and stores it under every module name in `names`
    _run_pass(model, pipe.tokenizer, [LAST_HIDDEN], files, stats, args, sample_size=sample_size, device=device,
              progress=progress, force_recompute=force_recompute, checkpoint=None, accumulator_factory=None,
              write_also=names, **opts)
    return stats[layer_name]


def compute_cross_attn_kv_stats(pipe, dataset="ccs_filtered", to_collect=["mom2"], sample_size=100000,
                                batch_tokens=3 * 1024, precision="float32", stats_dir="data/stats", download=0,
                                force_recompute=False, **b200_options):
    """All cross-attention K/V statistics "in one go" (emcid/layer_stats.py:429-467) — here literally one pass; takes
    the pipeline instead of loading CompVis/stable-diffusion-v1-4 itself.  Returns {layer name: CombinedStat}."""
    names = get_all_cross_attn_kv_layer_names(pipe)
    out = {}
    for n in names:   # the first call computes and writes every file, the others are cache hits
        out[n] = layer_stats_cross_attn_kv(pipe, n, stats_dir=stats_dir, ds_name=dataset, to_collect=to_collect,
                                           sample_size=sample_size, precision=precision, batch_tokens=batch_tokens,
                                           download=download, force_recompute=force_recompute and n == names[0],
                                           share_with=names, **b200_options)
    return out


# ---------------------------------------------------------------------------------------------------------
# command line: pre-cache the statistics of a text encoder (emcid/layer_stats.py:34-134)
# ---------------------------------------------------------------------------------------------------------
def _load_text_encoder(model_name: str, device, synthetic: int):
    """(model, tokenizer, dataset or None).  The reference loads the SD / SDXL pipelines with diffusers (:59-109); in an
    environment without them (`--synthetic N`) a random-init encoder of the same shape and N synthetic captions stand in."""
    if synthetic:
        from . import synth

        model = synth.make_text_encoder(model_name, seed=0).to(device)
        return model, None, synth.CaptionMatrixDataset(synth.make_caption_matrix(synthetic, seed=7))
    try:
        if model_name == "sd-text":
            from diffusers import StableDiffusionPipeline

            pipe = StableDiffusionPipeline.from_pretrained("CompVis/stable-diffusion-v1-4", torch_dtype=torch.float32,
                                                           safety_checker=None, requires_safety_checker=False)
            model, tokenizer = pipe.text_encoder, pipe.tokenizer
        else:
            from diffusers import StableDiffusionXLPipeline

            pipe = StableDiffusionXLPipeline.from_pretrained("stabilityai/stable-diffusion-xl-base-1.0",
                                                             torch_dtype=torch.float32, safety_checker=None,
                                                             requires_safety_checker=False, use_safetensors=True, variant="fp16")
            model, tokenizer = ((pipe.text_encoder, pipe.tokenizer) if model_name == "sdxl-text1" else
                                (pipe.text_encoder_2, pipe.tokenizer_2))
    except ImportError as e:
        raise SystemExit(f"loading {model_name} needs diffusers ({e}); use --synthetic N for a random-init encoder") from None
    model = model.float().eval().to(device)
    for p in model.parameters():
        p.requires_grad_(False)
    return model, tokenizer, None


def main(argv=None):
    """python -m emcid_b200.layer_stats --model_name sd-text --layers 7,8,9,10,11 [--sample_size 100000]

    Pre-caches the statistics files of ALL requested layers in one pass (the reference's CLI runs one pass per layer,
    :112-134).  Under `torchrun --nproc-per-node N` the captions are sharded over the N GPUs of the node."""
    import argparse

    from .globals import STATS_DIR

    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--model_name", default="sd-text", choices=["sd-text", "sdxl-text1", "sdxl-text2"])
    ap.add_argument("--dataset", default="ccs_filtered", choices=["ccs_filtered"])
    ap.add_argument("--layers", default="12", help="N = layers 0..N-1 like the reference's flag, or a comma-separated list")
    ap.add_argument("--to_collect", default=["mom2"], type=lambda x: x.split(","))
    ap.add_argument("--sample_size", default=100000, type=lambda x: None if x == "all" else int(x))
    ap.add_argument("--batch_tokens", default=3 * 1024, type=lambda x: None if x == "any" else int(x))
    ap.add_argument("--precision", default="float32", choices=["float64", "float32", "float16"])
    ap.add_argument("--stats_dir", default=STATS_DIR)
    ap.add_argument("--download", default=0, type=int, choices=[0, 1])
    ap.add_argument("--device", default=None, help="default: cuda:LOCAL_RANK")
    ap.add_argument("--force_recompute", action="store_true")
    ap.add_argument("--checkpoint_every", type=int, default=None, help="device blocks between resumable checkpoints")
    ap.add_argument("--synthetic", type=int, default=0, metavar="N",
                    help="random-init encoder of the named shape and N synthetic 77-token captions (no checkpoints needed)")
    a = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device(a.device or f"cuda:{local}")
    if device.type == "cuda":
        torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    layers = [int(x) for x in a.layers.split(",")] if "," in a.layers else list(range(int(a.layers)))
    model, tokenizer, synthetic_ds = _load_text_encoder(a.model_name, device, a.synthetic)
    if synthetic_ds is not None:
        global get_ccs_filtered_ds
        get_ccs_filtered_ds = lambda tokenizer: synthetic_ds
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in layers]
    print(f"Computing stats for layers {layers} of {a.model_name} over {a.sample_size or 'all'} samples of {a.dataset} "
          "in one pass. Note, the statistics are collected over the inputs to the second MLP layer, or equivalently the "
          "outputs of the first MLP layer.")
    t0 = time.perf_counter()
    stats = layer_stats_text_encoder_multi(model, tokenizer, names, a.stats_dir, a.dataset, a.to_collect,
                                           sample_size=a.sample_size, precision=a.precision, batch_tokens=a.batch_tokens,
                                           download=bool(a.download), force_recompute=a.force_recompute,
                                           checkpoint_every=a.checkpoint_every)
    if int(os.environ.get("RANK", "0")) == 0:
        counts = {n: stats[n].mom2.count for n in names}
        print(f"done in {time.perf_counter() - t0:.2f} s: {counts}")
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
