"""The EMCID text-encoder edit on B200: drop-in for the hot-path entry points of
emcid/emcid_main.py in SilentView/EMCID —

    get_cov_text_encoder                 :2239-2276
    execute_emcid_text_encoder           :818-1082     (stage 2, the insert loop :980-1073)
    apply_emcid_to_text_encoder          :769-815
    execute_emcid_sd_xl_text_encoders    :1085-1425
    apply_emcid_to_sdxl_text_encoders    :38-106
    execute_emcid_clip / apply_emcid_to_clip                 :109-311   (the text tower of a whole CLIPModel)
    execute_emcid_cross_attn / apply_emcid_to_cross_attn     :314-547   (UNet cross-attention to_k / to_v)
    get_cov_cross_attn                   :2203-2236
    upd_matrix_match_shape               :2279-2298

Same signatures, same return values (adj_k / resid as fp64 CPU tensors, weights restored on return
of execute_*, fc2 weights updated in place by apply_*).  The statistics come from
`emcid_b200.layer_stats` and the per-layer solve runs in `libemcid_b200.so` (csrc/solve.cuh):
3xTF32 tcgen05 Cholesky + TRSM with fp64-residual refinement instead of fp64 LU.

Stage 1 (the v* optimisation through UNet/VAE, emcid/compute_z.py) is outside this path: requests
must come with their `v_star` cached at `cache_name + "source_{src}_dest_{dst}.npz"` exactly as the
reference caches them (:873-969); a miss raises NotImplementedError.
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict
from copy import deepcopy
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import nethook
from .compute_ks import get_module_input_output_at_words, prefetch_keys, prepare_lookup
from .globals import STATS_DIR, XL_STATS_DIR1, XL_STATS_DIR2
from . import _lib
from .layer_stats import (get_all_cross_attn_kv_layer_names, layer_stats_cross_attn_kv, layer_stats_text_encoder,
                          layer_stats_text_encoder_multi)
from .solve import DEFAULT_REFINE_STEPS, CachedFactor, solve_layers

COV_CACHE: Dict[Tuple[str, str], torch.Tensor] = {}

# Factorisations of mom2_update_weight * C * (1 - edit_weight) / 0.5, kept on the device next to COV_CACHE for edits that
# come back with the same covariance and few concepts (sequential editing, experiments/sequential_editing.py:98-171, where
# the reference re-runs torch.linalg.solve on a fresh d x d matrix per edit and layer, emcid_main.py:1045-1047).
# Key: (model name, layer name, mom2_update_weight, edit_weight); value: (the COV_CACHE tensor it was built from, factor).
# An entry is rebuilt when COV_CACHE holds a different tensor (force_recompute, cleared cache).  20 d^2 bytes per entry
# (189 MB at d = 3072); least recently used entries beyond FACTOR_CACHE_MAX are dropped.  EMCID_FACTOR_CACHE=0 disables.
FACTOR_CACHE: "OrderedDict[Tuple[str, str, float, float], Tuple[torch.Tensor, CachedFactor]]" = OrderedDict()
FACTOR_CACHE_MAX = 16
# the cached path wins while the n_pad x n_pad system of the push-through identity stays small next to d x d
# (measured on B200, ms cached / direct, profiles/round1/r03e_probe_factor.json: d = 3072: n = 100 2.9 / 6.3, 300 4.8 / 6.7,
#  700 6.9 / 7.2, 1000 13.3 / 7.4;  d = 5120: n = 100 9.5 / 14.2, 1000 50 / 20)
FACTOR_CACHE_MAX_FRACTION = 6


# which solver produced each layer's update in the last edit ("cached_factor", "direct", "fp64_lu"); tests / bench read it
LAST_SOLVE_PATHS: List[str] = []


def _fp64_lu_on_device(cov_raw, layer_ks, sources_t, mom2_update_weight, ew, layers_left):
    """The reference's own arithmetic (emcid_main.py:1037-1050: fp64 M, torch.linalg.solve = LU with partial pivoting)
    on the device — the last resort for systems the fp32-class Cholesky cannot factor or does not contract on."""
    scale = (ew / 0.5) ** 0.5
    Ks = layer_ks.double().T * scale                                   # [d, n]
    Ss = sources_t.double().T * scale                                  # [h, n]
    M = mom2_update_weight * (cov_raw * (1 - ew) / 0.5).double() + Ks @ Ks.T
    adj_k = torch.linalg.solve(M, Ks)
    resid = Ss / layers_left
    return adj_k, resid, (resid @ adj_k.T).float()


def _solve_one_layer(text_encoder, module_name: str, cov_raw: torch.Tensor, layer_ks: torch.Tensor,
                     sources_t: torch.Tensor, mom2_update_weight: float, ew: float, layers_left: int, refine_steps: int):
    """adj_k, resid, dW of one layer (emcid_main.py:1037-1050): through the cached factor of lambda * C32 when the edit is
    narrow (n_pad <= d / FACTOR_CACHE_MAX_FRACTION), else the direct batched solver.

    The reference solves with fp64 LU, which does not care about conditioning; a Cholesky factorisation in fp32-class
    arithmetic breaks down, or stops contracting under refinement, once the DIAGONALLY SCALED condition number of the
    matrix approaches 1 / eps_fp32 (diagonal equilibration would change nothing: with power-of-two scales the scaled
    factorisation is bit for bit the scaled factor).  A breakdown or a refinement that misses its target therefore falls
    through: cached factor of lambda*C  ->  direct factorisation of lambda*C + K K^T (better conditioned)  ->  the
    reference's fp64 LU on the device (torch.linalg.solve), with a RuntimeWarning naming the layer."""
    d = cov_raw.shape[0]
    n_pad = -(-layer_ks.shape[0] // 128) * 128
    scale = (ew / 0.5) ** 0.5
    if os.environ.get("EMCID_FACTOR_CACHE", "1") != "0" and n_pad * FACTOR_CACHE_MAX_FRACTION <= d:
        key = (text_encoder.config._name_or_path.replace("/", "_"), module_name, float(mom2_update_weight), float(ew))
        try:
            entry = FACTOR_CACHE.get(key)
            if entry is None or entry[0] is not cov_raw:
                if entry is not None:
                    FACTOR_CACHE.pop(key)[1].close()
                entry = (cov_raw, CachedFactor(cov_raw * (1 - ew) / 0.5, mom2_update_weight))      # fp32 scaling, :1037
                FACTOR_CACHE[key] = entry
                while len(FACTOR_CACHE) > FACTOR_CACHE_MAX:
                    FACTOR_CACHE.popitem(last=False)[1][1].close()
            FACTOR_CACHE.move_to_end(key)
            out = entry[1].solve(layer_ks.float(), sources_t.float(), scale, layers_left, refine_steps=refine_steps,
                                 strict=True)
            LAST_SOLVE_PATHS.append("cached_factor")
            return out
        except _lib.EmcidError as e:
            if e.code != -4:
                raise
            hit = FACTOR_CACHE.pop(key, None)
            if hit is not None:
                hit[1].close()
    try:
        adj_k, resid, dW = solve_layers(cov_raw * (1 - ew) / 0.5, layer_ks.float(), sources_t.float(), mom2_update_weight,
                                        scale, [layers_left], refine_steps=refine_steps, strict=True)
        LAST_SOLVE_PATHS.append("direct")
        return adj_k[0], resid[0], dW[0]
    except _lib.EmcidError as e:
        if e.code != -4:
            raise
        import warnings

        warnings.warn(f"emcid_b200: {module_name}: {e}; solving this layer with fp64 LU on the device "
                      "(the reference's arithmetic) instead", RuntimeWarning)
    LAST_SOLVE_PATHS.append("fp64_lu")
    return _fp64_lu_on_device(cov_raw, layer_ks, sources_t, mom2_update_weight, ew, layers_left)


def _copy_requests(requests):
    """The reference deep-copies the request list on entry (emcid_main.py:850) so that nothing it does can reach the
    caller's objects; nothing here writes to a request, so a copy of the list and of each request dict gives the same
    isolation (the deep copy of 1000 requests was 1.7 ms of every edit call)."""
    return [dict(r) for r in requests]


def clear_factor_cache() -> None:
    while FACTOR_CACHE:
        FACTOR_CACHE.popitem()[1][1].close()

# Set TIMING = True to have the edit loop synchronise the device between its stages and leave their wall-clock
# milliseconds in LAST_EDIT_TIMING (bench.py does, for one extra repetition; off by default: no extra syncs).
TIMING = False
LAST_EDIT_TIMING: Dict[str, float] = {}


def _tick(device, key: str, t0: float) -> float:
    import time

    if not TIMING:
        return t0
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)
    t1 = time.perf_counter()
    LAST_EDIT_TIMING[key] = LAST_EDIT_TIMING.get(key, 0.0) + 1e3 * (t1 - t0)
    return t1


def get_cov_text_encoder(model, tok, layer_name: str, mom2_dataset: str, mom2_n_samples: int, mom2_dtype: str,
                         inv: bool = False, force_recompute: bool = False, verbose: bool = True,
                         stat_dir: str = STATS_DIR) -> torch.Tensor:
    """C = mom2 / count as fp32 on model.device; cached per (model name, layer) like the reference,
    but the cached copy already lives on the device (no 37.7 MB H2D per layer per edit)."""
    model_name = model.config._name_or_path.replace("/", "_")
    key = (model_name, layer_name)
    if verbose:
        print(f"Retrieving covariance statistics for {model_name} @ {layer_name}.")
    if key not in COV_CACHE or force_recompute:
        stat = layer_stats_text_encoder(model, tok, layer_name, stat_dir, mom2_dataset, to_collect=["mom2"],
                                        sample_size=mom2_n_samples, precision=mom2_dtype,
                                        force_recompute=force_recompute)
        COV_CACHE[key] = stat.mom2.moment().float().to(model.device)
    cov = COV_CACHE[key].to(model.device)
    return torch.inverse(cov) if inv else cov


def _to_host_async(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if not t.is_cuda:
        return t.cpu()
    host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    host.copy_(t, non_blocking=True)
    return host


def upd_matrix_match_shape(matrix: torch.Tensor, shape: torch.Size) -> torch.Tensor:
    if matrix.shape == shape:
        return matrix
    if matrix.T.shape == shape:
        return matrix.T
    if len(matrix.shape) == 2 and len(shape) == 4:
        return matrix.reshape(shape[0], shape[1], *shape[2:])
    print(f"matrix shape: {matrix.shape}")
    print(f"desired shape: {shape}")
    raise ValueError("Update matrix computed by EMCIDdoes not match original weight shape. "
                     "Check for bugs in the code?")


_NPY_HEADERS: Dict[bytes, Tuple[np.dtype, tuple, int]] = {}   # npy header bytes -> (dtype, shape, element count)


def _parse_npy_header(text: bytes):
    """(dtype, shape, count) of an npy header dict; every v* file of an edit carries the same header bytes, so the
    parse (ast.literal_eval: 50 us) is memoised on them."""
    import ast
    import math

    hit = _NPY_HEADERS.get(text)
    if hit is None:
        header = ast.literal_eval(text.decode("latin1"))
        dtype = np.dtype(header["descr"])
        if header["fortran_order"] or dtype.hasobject:
            raise ValueError
        shape = tuple(int(x) for x in header["shape"])
        hit = (dtype, shape, math.prod(shape))
        if len(_NPY_HEADERS) < 64:
            _NPY_HEADERS[text] = hit
    return hit


def _read_npz_array(path, key: str) -> np.ndarray:
    """`np.load(path)[key]` for the v* cache files (emcid_main.py:886-901: `np.savez(file, v_star=...)`, one small stored
    array per file).  np.load spends ~90 us per file in zipfile / NpzFile machinery, which is most of a 1000-concept edit's
    host time; the stored member can be located from the local file header alone.  Anything unexpected (compression, another
    member first, object arrays, fortran order) falls back to np.load.  A missing file raises FileNotFoundError."""
    import struct

    with open(path, "rb") as f:
        raw = f.read()
    try:
        if raw[:4] != b"PK\x03\x04":
            raise ValueError
        method, = struct.unpack_from("<H", raw, 8)
        n_name, n_extra = struct.unpack_from("<HH", raw, 26)
        name = raw[30:30 + n_name].decode()
        if method != 0 or name != key + ".npy":
            raise ValueError
        off = 30 + n_name + n_extra
        if raw[off:off + 6] != b"\x93NUMPY":
            raise ValueError
        major = raw[off + 6]
        if major == 1:
            hlen, = struct.unpack_from("<H", raw, off + 8)
            hstart = off + 10
        else:
            hlen, = struct.unpack_from("<I", raw, off + 8)
            hstart = off + 12
        dtype, shape, count = _parse_npy_header(raw[hstart:hstart + hlen])
        data = hstart + hlen
        return np.frombuffer(raw, dtype=dtype, count=count, offset=data).reshape(shape).copy()
    except (ValueError, KeyError, SyntaxError, TypeError, struct.error, IndexError):
        return np.load(path)[key]


def _vstar_stem(request, idx, hparams, naming: str) -> str:
    """File stem of a request's cached v* (emcid_main.py:873-890; the SDXL loop only ever uses the source/dest form,
    :1157-1166)."""
    if naming == "sd":
        if "esd" in getattr(hparams, "objective", ""):
            return f"source_{request['source']}"
        if getattr(hparams, "sld_supervision", False):
            return f"source_{request['source_cat']}_{idx}"
    return f"source_{request['source']}_dest_{request['dest']}"


def _vstar_path_check(requests, hparams, cache_name, suffix="", naming="sd"):
    first = cache_name + _vstar_stem(requests[0], 0, hparams, naming) + suffix + ".npz"
    if not os.path.exists(first):
        raise NotImplementedError(
            f"v_star cache miss for request {requests[0]['source']!r} ({first}): stage 1 (compute_z, UNet/VAE "
            "optimisation) is outside the B200 hot path — precompute v_star with the reference")


def _load_vstars(requests, hparams, cache_name, device, suffix="", naming="sd"):
    """zs [h, n] on `device` from the per-request cache files.  A missing file and an unreadable one are both cache
    misses (the reference logs the read error and recomputes, :892-907); recomputing v* is stage 1, outside this path, so
    either is a NotImplementedError that names the file."""
    import zipfile

    zs = []
    for idx, request in enumerate(requests):
        path = cache_name + _vstar_stem(request, idx, hparams, naming) + suffix + ".npz" if cache_name is not None else None
        try:
            if path is None:
                raise FileNotFoundError
            zs.append(torch.from_numpy(_read_npz_array(path, "v_star")))      # no stat() first: open() finds out
        except FileNotFoundError:
            raise NotImplementedError(
                f"v_star cache miss for request {request['source']!r} ({path}): stage 1 (compute_z, UNet/VAE "
                "optimisation) is outside the B200 hot path — precompute v_star with the reference") from None
        except (zipfile.BadZipFile, OSError, KeyError, ValueError, EOFError) as e:
            raise NotImplementedError(
                f"v_star cache miss for request {request['source']!r}: {path} is unreadable ({type(e).__name__}: {e}); "
                "stage 1 (compute_z) is outside the B200 hot path — recompute v_star with the reference") from None
    # stacked on the host, one H2D copy (the reference moves every v* separately, :892-901)
    if getattr(hparams, "use_new_compute_z", False):
        z = torch.stack(zs, dim=0).to(device)           # [rq, num, h]
        return z.permute(2, 0, 1).reshape(z.shape[2], -1)
    return torch.stack(zs, dim=1).to(device)            # [h, n]


class _VstarPrefetch:
    """The n v* files of an edit, read by the library (`emcid_read_npz_f32`) on a helper thread that holds no interpreter
    lock, started before the prompts are tokenised and collected (`__call__`) after the first key extraction has been
    launched.  The first file is read here, on the calling thread: it gives the shape every other file must have, and a
    missing one fails the edit before any device work (the reference would recompute v*, :892-907 — stage 1, outside this
    path).  Anything the C reader does not take (another dtype or member order, a compressed archive, a missing or
    damaged file) goes to `_load_vstars`, which reads the general layout and words the errors."""

    def __init__(self, requests, hparams, cache_name, device, suffix="", naming="sd"):
        import threading

        self.args = (requests, hparams, cache_name, device, suffix, naming)
        self.thread = None
        if cache_name is None:
            return
        _vstar_path_check(requests, hparams, cache_name, suffix, naming)
        paths = [cache_name + _vstar_stem(r, i, hparams, naming) + suffix + ".npz" for i, r in enumerate(requests)]
        try:
            first = _read_npz_array(paths[0], "v_star")
        except Exception:
            return                                                   # _load_vstars words the error at collection time
        if first.dtype != np.float32 or first.size == 0:
            return
        pin = torch.device(device).type == "cuda"
        self.host = torch.empty((len(paths),) + tuple(first.shape), dtype=torch.float32, pin_memory=pin)
        self.rc = None
        arr = (ctypes.c_char_p * len(paths))(*[os.fsencode(q) for q in paths])
        bad = ctypes.c_int(-1)

        def work():
            self.rc = _lib.lib().emcid_read_npz_f32(arr, len(paths), b"v_star", self.host.data_ptr(), first.size,
                                                    ctypes.byref(bad))

        self.thread = threading.Thread(target=work, name="emcid-vstar-read", daemon=True)
        self.thread.start()

    def __call__(self):
        requests, hparams, cache_name, device, suffix, naming = self.args
        if self.thread is None:
            return _load_vstars(requests, hparams, cache_name, device, suffix, naming)
        self.thread.join()
        if self.rc != 0:
            return _load_vstars(requests, hparams, cache_name, device, suffix, naming)
        z = self.host.to(device, non_blocking=True)                  # one H2D copy (the reference: one per request)
        if getattr(hparams, "use_new_compute_z", False):
            return z.permute(2, 0, 1).reshape(z.shape[2], -1)        # [rq, num, h] -> [h, rq * num]
        return z.T                                                   # [h, n]


# fp32 [h, d] updates of the last insert loop, by weight name, still on the device: apply_* adds them in place instead of
# re-forming adj_k @ resid^T from the host copies (reference :802-809; SURVEY.md §8 a11 "or skip: dW already known")
_DEVICE_UPDATES: Dict[int, Dict[str, torch.Tensor]] = {}


def _prefetch_covariances(text_encoder, tokenizer, hparams, layers, stat_dir, verbose):
    """Cold COV_CACHE: every missing layer's statistics file is loaded — or, on a cold disk cache too, computed — by ONE
    multi-layer pass (the reference's get_cov_text_encoder runs one pass per layer, :2263-2272)."""
    model_name = text_encoder.config._name_or_path.replace("/", "_")
    names = [hparams.rewrite_module_tmp.format(l) for l in layers]
    missing = [n for n in names if (model_name, n) not in COV_CACHE]
    if len(missing) < 2 or not all(n.endswith(".fc2") for n in missing):
        return
    stats = layer_stats_text_encoder_multi(text_encoder, tokenizer, missing, stat_dir, hparams.mom2_dataset,
                                           to_collect=["mom2"], sample_size=hparams.mom2_n_samples,
                                           precision=hparams.mom2_dtype, broadcast=True,
                                           progress=None if not verbose else layer_stats_progress())
    for n in missing:
        COV_CACHE[(model_name, n)] = stats[n].mom2.moment().float().to(text_encoder.device)


def layer_stats_progress():
    from .layer_stats import tqdm

    return tqdm


def _insert_loop(text_encoder, tokenizer, requests, hparams, layers, zs, mom2_update_weight, stat_dir, verbose,
                 refine_steps, host_deltas: bool = True):
    """The stage-2 loop of the reference (:980-1073) for one encoder.  `zs`: the [h, n] targets, or a callable producing
    them — called right after the first layer's key extraction has been LAUNCHED, so that the host-side reading of the
    1000 small v* files (10 ms of Python) runs while the device works through the forward of the prompts."""
    device = text_encoder.device
    names = [f"{hparams.rewrite_module_tmp.format(l)}.weight" for l in layers]
    weights = {n: nethook.get_parameter(text_encoder, n) for n in names}
    weights_copy = {n: w.detach().clone() for n, w in weights.items()}
    deltas = {}
    ew = hparams.edit_weight
    # prompts, token ids and last-subject-token positions are the same for every layer: tokenise / look up once
    import time

    t = time.perf_counter()
    # ... and the device starts on the layers below the first edited one as soon as the token ids exist
    first_module = hparams.rewrite_module_tmp.format(layers[0])
    prepared = prepare_lookup(tokenizer, requests, hparams.num_edit_tokens, device,
                              after_tokenise=lambda enc, serial: prefetch_keys(text_encoder, enc, first_module, serial))
    t = _tick(device, "tokenise_lookup_ms", t)
    _prefetch_covariances(text_encoder, tokenizer, hparams, layers, stat_dir, verbose)
    updates = _DEVICE_UPDATES[id(text_encoder)] = {}
    try:
        with torch.no_grad():
            for i, layer in enumerate(layers):
                module_name = hparams.rewrite_module_tmp.format(layer)
                # one forward gives both the keys (:987-996) and the current outputs (:1004-1014)
                layer_ks, cur_zs = get_module_input_output_at_words(
                    text_encoder, tokenizer, requests, module_name, num_fact_token=hparams.num_edit_tokens,
                    prepared=prepared)
                if callable(zs):
                    t_read = time.perf_counter()
                    zs = zs()                                                     # host work beside the queued forward
                    LAST_EDIT_TIMING["vstar_npz_read_host_ms"] = 1e3 * (time.perf_counter() - t_read)
                t = _tick(device, "keys_ms", t)
                if hparams.num_edit_tokens > 1:
                    layer_ks = layer_ks.reshape(-1, layer_ks.shape[-1])
                    cur_zs = cur_zs.reshape(-1, cur_zs.shape[-1])
                if verbose:
                    print(f"\n\nLAYER {layer}\n")
                    print(f"Writing {layer_ks.size(0)} key/value pair(s) into layer {layer}")
                sources_t = zs.T.to(cur_zs.dtype) - cur_zs                       # (zs - cur_zs)^T, [n, h]
                if verbose:
                    print("z error", torch.linalg.norm(sources_t, dim=1).mean())
                cov = get_cov_text_encoder(text_encoder, tokenizer, module_name, hparams.mom2_dataset,
                                           hparams.mom2_n_samples, hparams.mom2_dtype, stat_dir=stat_dir,
                                           force_recompute=False, verbose=verbose)
                adj_k, resid, dW = _solve_one_layer(text_encoder, module_name, cov, layer_ks, sources_t,
                                                    mom2_update_weight, ew, len(layers) - i, refine_steps)
                t = _tick(device, "solve_ms", t)
                name = names[i]
                upd = upd_matrix_match_shape(dW, weights[name].shape)
                if verbose:
                    print("orig norm", torch.linalg.norm(weights[name]))
                    print("upd norm", torch.linalg.norm(upd))
                weights[name][...] = weights_copy[name] + upd                     # :1061 — next layer sees it
                updates[name] = upd
                # the reference returns CPU tensors (:1062-1065); copy asynchronously into pinned memory so the transfer
                # of layer i (30 MB of fp64) overlaps the key extraction and the solve of layer i + 1
                deltas[name] = (_to_host_async(adj_k), _to_host_async(resid)) if host_deltas else (adj_k, resid)
                t = _tick(device, "write_and_d2h_ms", t)
        if torch.device(device).type == "cuda":
            torch.cuda.current_stream(device).synchronize()      # all delta copies have landed
    finally:
        with torch.no_grad():
            for n, w in weights.items():
                w[...] = weights_copy[n]                                          # :1075-1078
    return deltas


def execute_emcid_text_encoder(pipe, requests: List[Dict], hparams, cache_name: Optional[str] = None,
                               mom2_weight: Optional[int] = None, edit_weight: Optional[float] = None,
                               verbose: bool = True, stat_dir=STATS_DIR,
                               refine_steps: int = DEFAULT_REFINE_STEPS) -> Dict[str, Tuple[torch.Tensor, torch.Tensor]]:
    """Executes the EMCID update algorithm; invariant: model weights at return == at entry."""
    device = pipe.device
    hparams.mom2_update_weight = mom2_weight if mom2_weight is not None else hparams.mom2_update_weight
    hparams.edit_weight = edit_weight if edit_weight is not None else hparams.edit_weight
    requests = _copy_requests(requests)
    if verbose:
        for request in requests:
            print(f"EMCID request sample: [{request['source']}] -> [{request['dest']}]")
    return _execute_text_encoder(pipe, requests, hparams, cache_name, verbose, stat_dir, refine_steps, host_deltas=True)


def _execute_text_encoder(pipe, requests, hparams, cache_name, verbose, stat_dir, refine_steps, host_deltas):
    device = pipe.device
    LAST_EDIT_TIMING.clear()
    del LAST_SOLVE_PATHS[:]
    # a cache miss fails here, before any device work (the first file stands for all); the rest is read beside the
    # tokenisation and the first forward
    deltas = _insert_loop(pipe.text_encoder, pipe.tokenizer, requests, hparams, hparams.layers,
                          _VstarPrefetch(requests, hparams, cache_name, device),
                          hparams.mom2_update_weight, stat_dir, verbose, refine_steps, host_deltas=host_deltas)
    print(f"Deltas successfully computed for {list(deltas.keys())}")
    return deltas


def _apply_deltas(model, deltas, device):
    """w += (adj_k @ resid^T)^T.float() for every edited weight (reference :802-809).  The insert loop that produced
    `deltas` left exactly that matrix — float(resid @ adj_k^T), formed in fp64 by the solver — on the device: it is
    added in place.  Deltas that did not come from the last insert loop on this model (a caller replaying saved
    deltas) take the reference's route: the fp64 product, on the library's DMMA GEMM."""
    updates = _DEVICE_UPDATES.pop(id(model), {})
    with torch.no_grad():
        for w_name, (key_mat, val_mat) in deltas.items():
            w = nethook.get_parameter(model, w_name)
            upd = updates.get(w_name)
            if upd is None or upd.device != w.device:
                upd = _update_from_deltas(key_mat.to(w.device), val_mat.to(w.device))
            # `upd` = float(resid @ adj_k^T) is [out, in]; the reference forms adj_k @ resid^T = upd^T ([in, out]) and lets
            # upd_matrix_match_shape pick the orientation (:805-807).  Going through upd^T keeps its behaviour for SQUARE
            # weights too, where the shapes already "match" and the reference adds the product untransposed.
            w[...] += upd_matrix_match_shape(upd.T, w.shape).to(w.dtype)


def _update_from_deltas(adj_k: torch.Tensor, resid: torch.Tensor) -> torch.Tensor:
    """float(resid @ adj_k^T) [h, d] from fp64 CUDA deltas (adj_k [d, n], resid [h, n]) through emcid_delta_update."""
    adj_k, resid = adj_k.double().contiguous(), resid.double().contiguous()
    d, n = adj_k.shape
    h = resid.shape[0]
    out = torch.empty(h, d, dtype=torch.float32, device=adj_k.device)
    with torch.cuda.device(adj_k.device):
        _lib.check(_lib.lib().emcid_delta_update(h, d, n, _lib.ptr(resid), _lib.ptr(adj_k), _lib.ptr(out),
                                                 _lib.current_stream_ptr()))
    return out


def apply_emcid_to_text_encoder(pipe, requests: List[Dict], hparams, device: str, mom2_weight: Optional[int] = None,
                                edit_weight: Optional[float] = None, return_orig_text_encoder=False,
                                cache_name: Optional[str] = None, stats_dir: Optional[str] = STATS_DIR,
                                verbose: bool = True):
    """Returns (pipe with edited text encoder, original text encoder or None)."""
    origin_text_encoder = None
    if return_orig_text_encoder:
        origin_text_encoder = deepcopy(pipe.text_encoder).to("cpu")
    hparams.mom2_update_weight = mom2_weight if mom2_weight is not None else hparams.mom2_update_weight
    hparams.edit_weight = edit_weight if edit_weight is not None else hparams.edit_weight
    requests = _copy_requests(requests)
    if verbose:
        for request in requests:
            print(f"EMCID request sample: [{request['source']}] -> [{request['dest']}]")
    # the deltas stay on the device: nothing here reads them on the host (execute_emcid_text_encoder returns CPU copies)
    deltas = _execute_text_encoder(pipe, requests, hparams, cache_name, verbose, stats_dir, DEFAULT_REFINE_STEPS,
                                   host_deltas=False)
    _apply_deltas(pipe.text_encoder, deltas, device)
    print(f"New weights successfully inserted into {list(deltas.keys())}")
    if return_orig_text_encoder:
        origin_text_encoder = origin_text_encoder.to(device)
    return pipe, origin_text_encoder


def _execute_sdxl(pipe, requests, hparams, cache_name, verbose, stat_dir, stat_dir_2, refine_steps, host_deltas):
    device = pipe.device
    LAST_EDIT_TIMING.clear()
    del LAST_SOLVE_PATHS[:]
    zs = _load_vstars(requests, hparams, cache_name, device, naming="sdxl")
    zs_2 = _load_vstars(requests, hparams, cache_name, device, suffix="_2", naming="sdxl")
    deltas = _insert_loop(pipe.text_encoder, pipe.tokenizer, requests, hparams, hparams.layers, zs,
                          hparams.mom2_update_weight, stat_dir, verbose, refine_steps, host_deltas=host_deltas)
    deltas_2 = _insert_loop(pipe.text_encoder_2, pipe.tokenizer_2, requests, hparams, hparams.layers_2, zs_2,
                            hparams.mom2_update_weight_2, stat_dir_2, verbose, refine_steps, host_deltas=host_deltas)
    print(f"Deltas successfully computed for {list(deltas.keys())} and {list(deltas_2.keys())}")
    return deltas, deltas_2


def _set_sdxl_hparams(hparams, mom2_weight, mom2_weight_2, edit_weight):
    hparams.mom2_update_weight = mom2_weight if mom2_weight is not None else hparams.mom2_update_weight
    hparams.mom2_update_weight_2 = mom2_weight_2 if mom2_weight_2 is not None else hparams.mom2_update_weight_2
    hparams.edit_weight = edit_weight if edit_weight is not None else hparams.edit_weight


def execute_emcid_sd_xl_text_encoders(pipe, requests: List[Dict], hparams, cache_name: Optional[str] = None,
                                      mom2_weight: Optional[int] = None, mom2_weight_2: Optional[int] = None,
                                      edit_weight: Optional[float] = None, verbose: bool = True,
                                      stat_dir="data/stats/sdxl/text1", stat_dir_2="data/stats/sdxl/text2",
                                      refine_steps: int = DEFAULT_REFINE_STEPS):
    """Both SDXL text encoders: `layers` / `mom2_update_weight` / stat_dir for text_encoder,
    `layers_2` / `mom2_update_weight_2` / stat_dir_2 for text_encoder_2 (v_star files end in _2)."""
    _set_sdxl_hparams(hparams, mom2_weight, mom2_weight_2, edit_weight)
    return _execute_sdxl(pipe, _copy_requests(requests), hparams, cache_name, verbose, stat_dir, stat_dir_2, refine_steps, True)


def apply_emcid_to_sdxl_text_encoders(pipe, requests: List[Dict], hparams, device: str,
                                      mom2_weight: Optional[int] = None, mom2_weight_2: Optional[int] = None,
                                      edit_weight: Optional[float] = None, return_orig_text_encoder=False,
                                      cache_name: Optional[str] = None, stat_dir: Optional[str] = XL_STATS_DIR1,
                                      stat_dir_2: Optional[str] = XL_STATS_DIR2, verbose: bool = True):
    origin_1 = origin_2 = None
    if return_orig_text_encoder:
        origin_1 = deepcopy(pipe.text_encoder).to("cpu")
        origin_2 = deepcopy(pipe.text_encoder_2).to("cpu")
    _set_sdxl_hparams(hparams, mom2_weight, mom2_weight_2, edit_weight)
    deltas, deltas_2 = _execute_sdxl(pipe, _copy_requests(requests), hparams, cache_name, verbose, stat_dir, stat_dir_2,
                                     DEFAULT_REFINE_STEPS, False)
    _apply_deltas(pipe.text_encoder, deltas, device)
    _apply_deltas(pipe.text_encoder_2, deltas_2, device)
    print(f"New weights successfully inserted into {list(deltas.keys())}")
    if return_orig_text_encoder:
        origin_1, origin_2 = origin_1.to(device), origin_2.to(device)
    return pipe, origin_1, origin_2


# ---------------------------------------------------------------------------------------------------------
# the text tower of a whole CLIPModel (reference :109-311)
# ---------------------------------------------------------------------------------------------------------
def execute_emcid_clip(model, processor, requests: List[Dict], hparams, cache_name: Optional[str] = None,
                       verbose: bool = True, stat_dir=STATS_DIR, refine_steps: int = DEFAULT_REFINE_STEPS):
    """execute_emcid_clip (:151-311): the same stage-2 loop on `model` (a transformers CLIPModel whose
    hparams.rewrite_module_tmp names text_model.encoder.layers.{}.mlp.fc2), tokenizer = processor.tokenizer; v* files
    named source_{source}_dest_{dest}.npz.  Invariant: model weights at return == at entry."""
    requests = _copy_requests(requests)
    for request in requests:
        print(f"EMCID request sample: [{request['source']}] -> [{request['dest']}]")
    pipe = _ClipAsPipe(model, processor)
    LAST_EDIT_TIMING.clear()
    del LAST_SOLVE_PATHS[:]
    zs = _load_vstars(requests, hparams, cache_name, pipe.device, naming="clip")
    deltas = _insert_loop(model, pipe.tokenizer, requests, hparams, hparams.layers, zs, hparams.mom2_update_weight, stat_dir,
                          verbose, refine_steps)
    print(f"Deltas successfully computed for {list(deltas.keys())}")
    return deltas


class _ClipAsPipe:
    def __init__(self, model, processor):
        self.text_encoder, self.tokenizer, self.device = model, processor.tokenizer, model.device


def apply_emcid_to_clip(model, processor, requests: List[Dict], hparams, device: str, mom2_weight: Optional[int] = None,
                        edit_weight: Optional[float] = None, return_orig_text_model=False,
                        cache_name: Optional[str] = None, stat_dir=STATS_DIR):
    """Returns (the updated model, the original model or None) — apply_emcid_to_clip (:109-148)."""
    hparams.mom2_update_weight = mom2_weight if mom2_weight is not None else hparams.mom2_update_weight
    hparams.edit_weight = edit_weight if edit_weight is not None else hparams.edit_weight
    origin = deepcopy(model).to("cpu") if return_orig_text_model else None
    deltas = execute_emcid_clip(model, processor, requests, hparams, cache_name=cache_name, stat_dir=stat_dir)
    _apply_deltas(model, deltas, device)
    print(f"New weights successfully inserted into {list(deltas.keys())}")
    if origin is not None:
        origin = origin.to(device)
    return model, origin


# ---------------------------------------------------------------------------------------------------------
# UNet cross-attention to_k / to_v (reference :314-547)
# ---------------------------------------------------------------------------------------------------------
def get_cov_cross_attn(pipe, layer_name: str, mom2_dataset: str, sample_size: int, mom2_dtype: str, inv: bool = False,
                       force_recompute: bool = False, verbose: bool = False, stat_dir: str = STATS_DIR) -> torch.Tensor:
    """C = mom2 / count of the module's input (the text encoder's last_hidden_state), fp32 on pipe.device, cached per
    (unet name, layer) like the reference (:2203-2236)."""
    model_name = pipe.unet.config._name_or_path.replace("/", "_")
    key = (model_name, layer_name)
    if verbose:
        print(f"Retrieving covariance statistics for {model_name} @ {layer_name}.")
    if key not in COV_CACHE or force_recompute:
        stat = layer_stats_cross_attn_kv(pipe, layer_name, stat_dir, mom2_dataset, to_collect=["mom2"],
                                         sample_size=sample_size, precision=mom2_dtype, force_recompute=force_recompute)
        COV_CACHE[key] = stat.mom2.moment().float().to(pipe.device)
    cov = COV_CACHE[key].to(pipe.device)
    return torch.inverse(cov) if inv else cov


def _cross_attn_keys(pipe, requests):
    """K [n, hidden]: last_hidden_state at the last subject token of every source prompt, averaged over each request's
    prompts (emcid/compute_ks.py:52-139 reads it as the traced input of every attn2.to_k / to_v — the same tensor for all
    of them, so it is taken once, from the text encoder, and the UNet is not run)."""
    from .compute_ks import last_hidden_at_words

    return last_hidden_at_words(pipe.text_encoder, pipe.tokenizer, requests)


def _load_cross_attn_vstars(requests, cache_name, layer_names, device):
    """{layer: zs [out_features, n]} from source_{source}.npz files holding one pickled {"v_star": array} per layer
    (:365-396).  A miss is an error: computing them (compute_z_unet_x_kv) is stage 1."""
    per_layer = {n: [] for n in layer_names}
    for request in requests:
        path = cache_name + f"source_{request['source']}.npz" if cache_name is not None else None
        try:
            if path is None:
                raise FileNotFoundError
            data = np.load(path, allow_pickle=True)
            for n in layer_names:
                per_layer[n].append(torch.from_numpy(np.asarray(data[n].item()["v_star"])))
        except Exception as e:
            raise NotImplementedError(
                f"v_star cache miss for request {request['source']!r} ({path}: {type(e).__name__}): stage 1 "
                "(compute_z_unet_x_kv) is outside the B200 hot path — precompute v_star with the reference") from None
    return {n: torch.stack(v, dim=1).to(device) for n, v in per_layer.items()}


def execute_emcid_cross_attn(pipe, requests: List[Dict], hparams, cache_name: Optional[str] = None,
                             mom2_weight: Optional[int] = None, edit_weight: Optional[float] = None, verbose: bool = True,
                             stat_dir: str = STATS_DIR, refine_steps: int = DEFAULT_REFINE_STEPS):
    """execute_emcid_cross_attn (:314-508): closed-form update of every attn2.to_k / attn2.to_v of pipe.unet.

    Unlike the text-encoder edit the modules are INDEPENDENT (no module sees another's update, `resid = sources`,
    :470) and all of them read the same input, so there is one key matrix K and one system matrix
    M = lambda * C32 + Ks Ks^T for all 32 of them; only the right-hand sides (zs_l - W_l K) and the output widths differ.
    The modules are grouped by output width and each group goes through ONE batched library solve."""
    device = pipe.device
    hparams.mom2_update_weight = mom2_weight if mom2_weight is not None else hparams.mom2_update_weight
    hparams.edit_weight = edit_weight if edit_weight is not None else hparams.edit_weight
    requests = _copy_requests(requests)
    for request in requests:
        if "dest" in request:
            print(f"EMCID request sample: [{request['source']}] -> [{request['dest']}]")
        else:
            print(f"EMCID request sample: erasing [{request['source']}]")
    names = get_all_cross_attn_kv_layer_names(pipe)
    weights = {f"{n}.weight": nethook.get_parameter(pipe.unet, f"{n}.weight") for n in names}
    zs = _load_cross_attn_vstars(requests, cache_name, names, device)
    ew = hparams.edit_weight
    scale = (ew / 0.5) ** 0.5
    deltas = {}
    updates = _DEVICE_UPDATES[id(pipe.unet)] = {}
    del LAST_SOLVE_PATHS[:]
    with torch.no_grad():
        K = _cross_attn_keys(pipe, requests).float().contiguous()                  # [n, hidden]
        groups: Dict[int, List[str]] = {}
        for n in names:
            groups.setdefault(weights[f"{n}.weight"].shape[0], []).append(n)
        for out_features, members in groups.items():
            covs, raw_covs, sources = [], [], []
            for n in members:
                w = weights[f"{n}.weight"]
                bias = getattr(nethook.get_module(pipe.unet, n), "bias", None)
                cur = _lib.gemm3x_nt(K, w.detach().float().contiguous())           # W_l K: the module's current output, [n, out]
                if bias is not None:
                    cur = cur + bias
                sources.append(zs[n].T.float() - cur)                              # (zs - cur_zs)^T, [n, out]
                if verbose:
                    print(f"Writing {K.shape[0]} key/value pair(s) into layer {n}")
                    print("z error", torch.linalg.norm(sources[-1], dim=1).mean())
                raw_covs.append(get_cov_cross_attn(pipe, n, hparams.mom2_dataset, hparams.mom2_n_samples, hparams.mom2_dtype,
                                                   verbose=verbose, stat_dir=stat_dir))
                covs.append(raw_covs[-1] * (1 - ew) / 0.5)                         # fp32 scaling, as the reference forms it (:455)
            B = len(members)
            try:
                adj_k, resid, dW = solve_layers(torch.stack(covs), K[None].expand(B, -1, -1), torch.stack(sources),
                                                hparams.mom2_update_weight, scale, [1] * B, refine_steps=refine_steps,
                                                strict=True)
                LAST_SOLVE_PATHS.extend(["direct"] * B)
                adj_k, resid, dW = list(adj_k), list(resid), list(dW)
            except _lib.EmcidError as e:
                if e.code != -4:
                    raise
                import warnings

                warnings.warn(f"emcid_b200: cross-attention modules of width {out_features}: {e}; solving with fp64 LU on "
                              "the device instead", RuntimeWarning)
                trip = [_fp64_lu_on_device(c, K, s_, hparams.mom2_update_weight, ew, 1) for c, s_ in zip(raw_covs, sources)]
                adj_k, resid, dW = zip(*trip)
                LAST_SOLVE_PATHS.extend(["fp64_lu"] * B)
            for i, n in enumerate(members):
                name = f"{n}.weight"
                upd = upd_matrix_match_shape(dW[i], weights[name].shape)
                if verbose:
                    print("orig norm", torch.linalg.norm(weights[name]))
                    print("upd norm", torch.linalg.norm(upd))
                updates[name] = upd
                deltas[name] = (_to_host_async(adj_k[i]), _to_host_async(resid[i]))
        if torch.device(device).type == "cuda":
            torch.cuda.current_stream(device).synchronize()
    deltas = {f"{n}.weight": deltas[f"{n}.weight"] for n in names}                 # the reference's order
    print(f"Deltas successfully computed for {list(deltas.keys())}")
    return deltas


def apply_emcid_to_cross_attn(pipe, requests: List[Dict], hparams, device: str, mom2_weight: Optional[int] = None,
                              edit_weight: Optional[float] = None, return_orig_text_model=False,
                              cache_name: Optional[str] = None, stat_dir: str = STATS_DIR):
    """Returns (pipe with edited UNet cross-attention K/V projections, original unet or None) — :511-547."""
    orig_unet = deepcopy(pipe.unet) if return_orig_text_model else None
    deltas = execute_emcid_cross_attn(pipe, requests, hparams, cache_name=cache_name, mom2_weight=mom2_weight,
                                      edit_weight=edit_weight, stat_dir=stat_dir)
    _apply_deltas(pipe.unet, deltas, device)
    print(f"New weights successfully inserted into {list(deltas.keys())}")
    return pipe, orig_unet
