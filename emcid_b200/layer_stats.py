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
  * Captions are sharded `subset[rank::world]` across the GPUs of a torch.distributed job and the
    per-rank sums meet in one NCCL reduce per layer; counts are reduced as int64 (bit exact).

There is no CPU path: a model that is not on a CUDA (sm_100) device is an error.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import clip_forward, nethook
from .runningstats import (CombinedStat, FixedSubsetSampler, Mean, NormMean, SecondMoment, load_cached_state,
                           save_cached_state, subset_indices)
from .stat_dataset import (DEFAULT_BLOCK_TOKENS, PackedReblocker, TokenizedDataset, dict_to_, fixed_width_collation,
                           packed_collation, unpack_to_padded)

try:  # the reference's default progress bar
    from tqdm.auto import tqdm
except Exception:  # pragma: no cover
    tqdm = None

STAT_TYPES = {"mom2": SecondMoment, "mean": Mean, "norm_mean": NormMean}

# what the last computed pass did (tests and bench.py read it): {"native_forward": bool, "launches": int}
LAST_PASS_INFO: Dict[str, object] = {}


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
    """Streams padded caption blocks through the HF CLIP text model and accumulates mom2/count for
    a set of MLP layers in the same pass."""

    def __init__(self, model, layer_names: Sequence[str], slab_tokens: int = 0, accumulator_factory=None,
                 native: Optional[bool] = None):
        from .mom2 import Mom2Accumulator

        if native is None:
            native = os.environ.get("EMCID_NATIVE_FORWARD", "1") != "0"
        self._use_native = bool(native) and accumulator_factory is None and clip_forward.supports(model)
        self._native = None
        self.model = model
        self.layer_names = list(layer_names)
        mlps = {name: _mlp_of(model, name) for name in self.layer_names}  # LookupError before anything else
        device = next(model.parameters()).device
        if accumulator_factory is None:
            if device.type != "cuda":
                raise RuntimeError(
                    "emcid_b200.layer_stats needs the model on a CUDA sm_100 device (no CPU fallback); "
                    f"model is on {device}")
            accumulator_factory = lambda d, h, act: Mom2Accumulator(device, d, h, act, slab_tokens=slab_tokens)
        act = getattr(model.config, "hidden_act", None)
        if act not in ("quick_gelu", "gelu"):
            raise NotImplementedError(f"unsupported CLIP MLP activation {act!r}")
        self.accs: Dict[str, object] = {}
        self._mlps = []
        order = {n: i for i, (n, _) in enumerate(model.named_modules())}
        deepest = max(self.layer_names, key=lambda n: order[n])
        for name in self.layer_names:
            mlp = mlps[name]
            W1, b1 = mlp.fc1.weight, mlp.fc1.bias
            if W1.dtype != torch.float32:
                raise NotImplementedError("the statistics pass runs on fp32 weights (reference precision float32)")
            acc = accumulator_factory(W1.shape[0], W1.shape[1], act)
            self.accs[name] = acc
            self._mlps.append((name, mlp, name == deepest))
        self._mask = None
        self._hook_weights_set = False
        self._layer_index = {}
        self.capacity_hint = 0     # tokens per block the caller is going to send (sizes the native encoder once)
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

    @torch.no_grad()
    def run_batch(self, batch: Dict[str, torch.Tensor]) -> None:
        """batch: input_ids / position_ids / attention_mask [B, L] (host or device tensors), or an already
        packed block (packed_ids / packed_pos / cu_seqlens, see stat_dataset.packed_collation)."""
        if "packed_ids" in batch:
            max_pos = getattr(self.model, "text_model", self.model).embeddings.position_embedding.weight.shape[0]
            cu = batch["cu_seqlens"]
            S, T = cu.numel() - 1, batch["packed_ids"].numel()
            if T == 0:
                return
            if self._use_native and (cu.is_cuda or int((cu[1:] - cu[:-1]).max()) <= max_pos):
                names = list(self._layer_index)
                self._native_encoder(T, S).forward_stats(batch["packed_ids"], batch["packed_pos"], cu, S, T,
                                                         [self._layer_index[n] for n in names],
                                                         [self.accs[n] for n in names])
                return
            batch = unpack_to_padded(batch)
        if self._use_native:
            max_pos = getattr(self.model, "text_model", self.model).embeddings.position_embedding.weight.shape[0]
            packed = clip_forward.pack_batch(batch, max_pos)
            if packed is not None:
                ids, pos, cu, S, T = packed
                if T == 0:
                    return
                names = list(self._layer_index)
                self._native_encoder(T, S).forward_stats(ids, pos, cu, S, T, [self._layer_index[n] for n in names],
                                                         [self.accs[n] for n in names])
                return
            self._ensure_hook_weights()  # not a right-padding mask: keep the HF forward for this block
        device = next(self.model.parameters()).device
        batch = {k: v.to(device, non_blocking=True) for k, v in batch.items()}
        self._mask = batch["attention_mask"]
        handles = [mlp.register_forward_pre_hook(self._hook(name, stop)) for name, mlp, stop in self._mlps]
        try:
            self.model(**batch)
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


def _dist_info(distributed):
    import torch.distributed as dist

    if distributed is False or not (dist.is_available() and dist.is_initialized()):
        return None, 0, 1
    return dist, dist.get_rank(), dist.get_world_size()


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
    _accumulator_factory: Optional[Callable] = None,
) -> Dict[str, CombinedStat]:
    """All `layer_names` in one pass.  Arguments up to `force_recompute` mean exactly what they mean
    in the reference's layer_stats_text_encoder; the rest tune the B200 driver."""
    device = model.device
    if precision is None:
        precision = "float64"  # reference default (:161-162); every EMCID caller passes "float32"
    if precision != "float32":
        raise NotImplementedError("emcid_b200 accumulates mom2 in fp32 (hparams.mom2_dtype == 'float32')")
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

    import time
    t_start = time.perf_counter()
    dist, rank, world = _dist_info(distributed)
    ds = get_ccs_filtered_ds(tokenizer=tokenizer)
    indices = subset_indices(len(ds), sample_size, random_sample=1)  # tally(..., random_sample=1), reference :204
    my_indices = indices[rank::world]
    loader = torch.utils.data.DataLoader(
        ds, sampler=FixedSubsetSampler(my_indices), batch_size=captions_per_batch,
        collate_fn=packed_collation() if (device.type == "cuda" and _accumulator_factory is None) else fixed_width_collation(),
        num_workers=num_workers, pin_memory=(device.type == "cuda"))
    batch_count = -(-len(my_indices) // captions_per_batch)
    if progress is None:
        progress = lambda x, total=None: x

    runner = TextEncoderMom2Pass(model, todo, slab_tokens=slab_tokens, accumulator_factory=_accumulator_factory)
    reblock = PackedReblocker(block_tokens) if (block_tokens and runner._use_native) else None
    if reblock is not None:
        max_pos = getattr(model, "text_model", model).embeddings.position_embedding.weight.shape[0]
        runner.capacity_hint = min(int(block_tokens), len(my_indices) * int(max_pos))
    t_loop = time.perf_counter()
    try:
        t_wait = t_run = 0.0
        t_prev = time.perf_counter()
        for batch in progress(loader, total=batch_count):
            t_got = time.perf_counter()
            t_wait += t_got - t_prev
            if ("packed_ids" in batch and batch["packed_ids"].numel() == 0) or \
                    ("input_ids" in batch and batch["input_ids"].numel() == 0):
                t_prev = time.perf_counter()
                continue
            if reblock is not None and "packed_ids" in batch:
                # loader batches count captions, device blocks count tokens (stat_dataset.PackedReblocker)
                for block in reblock.push(batch):
                    runner.run_batch(block)
            else:
                runner.run_batch(batch)  # host tensors: packed on the host, then one pinned H2D copy per field
            t_prev = time.perf_counter()
            t_run += t_prev - t_got
        if reblock is not None:
            for block in reblock.flush():
                runner.run_batch(block)
        t_fin = time.perf_counter()
        results = runner.finalize()
        LAST_PASS_INFO.update(native_forward=runner._native is not None, launches=runner.launches())
        for i, n in enumerate(todo):
            mom2, count = results[n]
            root = i % world
            if dist is not None and world > 1:
                # the one exchange step of the pass: per-rank partial sums -> one matrix per layer
                # (every rank returns the full statistics, like the single-process reference call, so this is an
                # all-reduce; rank `root` additionally writes that layer's npz)
                dist.all_reduce(mom2, op=dist.ReduceOp.SUM)
                dist.all_reduce(count, op=dist.ReduceOp.SUM)
            sm = stats[n].mom2
            sm.count = int(count.item())
            sm.mom2 = mom2 if keep_on_device else mom2.to("cpu")
            if not force_recompute and rank == root:
                save_cached_state(files[n], stats[n], args)
        if dist is not None and world > 1:
            dist.barrier()
        # host-side timeline of the pass (no extra synchronisation: the loop time includes whatever the host waited for)
        t_end = time.perf_counter()
        LAST_PASS_INFO["timing"] = {"setup_s": t_loop - t_start, "loop_s": t_fin - t_loop, "finalize_s": t_end - t_fin,
                                    "loader_wait_s": t_wait, "run_batch_s": t_run}
    finally:
        runner.close()
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
    """Function to load or compute cached stats (signature of emcid/layer_stats.py:140-154)."""
    return layer_stats_text_encoder_multi(
        model, tokenizer, [layer_name], stats_dir=stats_dir, ds_name=ds_name, to_collect=to_collect,
        model_name=model_name, sample_size=sample_size, precision=precision, batch_tokens=batch_tokens,
        download=download, progress=progress, force_recompute=force_recompute, **b200_options)[layer_name]
