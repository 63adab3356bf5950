"""The EMCID text-encoder edit on B200: drop-in for the hot-path entry points of
emcid/emcid_main.py in SilentView/EMCID —

    get_cov_text_encoder                 :2239-2276
    execute_emcid_text_encoder           :818-1082     (stage 2, the insert loop :980-1073)
    apply_emcid_to_text_encoder          :769-815
    execute_emcid_sd_xl_text_encoders    :1085-1425
    apply_emcid_to_sdxl_text_encoders    :38-106
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

import os
from collections import OrderedDict
from copy import deepcopy
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import nethook
from .compute_ks import get_module_input_output_at_words, prepare_lookup
from .globals import STATS_DIR, XL_STATS_DIR1, XL_STATS_DIR2
from .layer_stats import layer_stats_text_encoder
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
# (measured on B200, ms cached / direct, profiles/r03e_probe_factor.json: d = 3072: n = 100 2.9 / 6.3, 300 4.8 / 6.7,
#  700 6.9 / 7.2, 1000 13.3 / 7.4;  d = 5120: n = 100 9.5 / 14.2, 1000 50 / 20)
FACTOR_CACHE_MAX_FRACTION = 6


def _solve_one_layer(text_encoder, module_name: str, cov_raw: torch.Tensor, layer_ks: torch.Tensor,
                     sources_t: torch.Tensor, mom2_update_weight: float, ew: float, layers_left: int, refine_steps: int):
    """adj_k, resid, dW of one layer (emcid_main.py:1037-1050): through the cached factor of lambda * C32 when the edit is
    narrow (n_pad <= d / FACTOR_CACHE_MAX_FRACTION), else the direct batched solver."""
    d = cov_raw.shape[0]
    n_pad = -(-layer_ks.shape[0] // 128) * 128
    scale = (ew / 0.5) ** 0.5
    if os.environ.get("EMCID_FACTOR_CACHE", "1") != "0" and n_pad * FACTOR_CACHE_MAX_FRACTION <= d:
        key = (text_encoder.config._name_or_path.replace("/", "_"), module_name, float(mom2_update_weight), float(ew))
        entry = FACTOR_CACHE.get(key)
        if entry is None or entry[0] is not cov_raw:
            if entry is not None:
                entry[1].close()
            entry = (cov_raw, CachedFactor(cov_raw * (1 - ew) / 0.5, mom2_update_weight))      # fp32 scaling, :1037
            FACTOR_CACHE[key] = entry
            while len(FACTOR_CACHE) > FACTOR_CACHE_MAX:
                FACTOR_CACHE.popitem(last=False)[1][1].close()
        FACTOR_CACHE.move_to_end(key)
        return entry[1].solve(layer_ks.float(), sources_t.float(), scale, layers_left, refine_steps=refine_steps)
    adj_k, resid, dW = solve_layers(cov_raw * (1 - ew) / 0.5, layer_ks.float(), sources_t.float(), mom2_update_weight,
                                    scale, [layers_left], refine_steps=refine_steps)
    return adj_k[0], resid[0], dW[0]


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


def _load_vstars(requests, hparams, cache_name, device, suffix=""):
    zs = []
    for idx, request in enumerate(requests):
        if "esd" in getattr(hparams, "objective", ""):
            stem = f"source_{request['source']}"
        elif getattr(hparams, "sld_supervision", False):
            stem = f"source_{request['source_cat']}_{idx}"
        else:
            stem = f"source_{request['source']}_dest_{request['dest']}"
        path = cache_name + stem + suffix + ".npz" if cache_name is not None else None
        try:
            if path is None:
                raise FileNotFoundError
            zs.append(torch.from_numpy(_read_npz_array(path, "v_star")))      # no stat() first: open() finds out
        except FileNotFoundError:
            raise NotImplementedError(
                f"v_star cache miss for request {request['source']!r} ({path}): stage 1 (compute_z, UNet/VAE "
                "optimisation) is outside the B200 hot path — precompute v_star with the reference") from None
    # stacked on the host, one H2D copy (the reference moves every v* separately, :892-901)
    if getattr(hparams, "use_new_compute_z", False):
        z = torch.stack(zs, dim=0).to(device)           # [rq, num, h]
        return z.permute(2, 0, 1).reshape(z.shape[2], -1)
    return torch.stack(zs, dim=1).to(device)            # [h, n]


def _insert_loop(text_encoder, tokenizer, requests, hparams, layers, zs, mom2_update_weight, stat_dir, verbose,
                 refine_steps):
    """The stage-2 loop of the reference (:980-1073) for one encoder."""
    device = text_encoder.device
    names = [f"{hparams.rewrite_module_tmp.format(l)}.weight" for l in layers]
    weights = {n: nethook.get_parameter(text_encoder, n) for n in names}
    weights_copy = {n: w.detach().clone() for n, w in weights.items()}
    deltas = {}
    ew = hparams.edit_weight
    # prompts, token ids and last-subject-token positions are the same for every layer: tokenise / look up once
    import time

    t = time.perf_counter()
    prepared = prepare_lookup(tokenizer, requests, hparams.num_edit_tokens, device)
    t = _tick(device, "tokenise_lookup_ms", t)
    try:
        with torch.no_grad():
            for i, layer in enumerate(layers):
                module_name = hparams.rewrite_module_tmp.format(layer)
                # one forward gives both the keys (:987-996) and the current outputs (:1004-1014)
                layer_ks, cur_zs = get_module_input_output_at_words(
                    text_encoder, tokenizer, requests, module_name, num_fact_token=hparams.num_edit_tokens,
                    prepared=prepared)
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
                # the reference returns CPU tensors (:1062-1065); copy asynchronously into pinned memory so the transfer
                # of layer i (30 MB of fp64) overlaps the key extraction and the solve of layer i + 1
                deltas[name] = (_to_host_async(adj_k), _to_host_async(resid))
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
    requests = deepcopy(requests)
    if verbose:
        for request in requests:
            print(f"EMCID request sample: [{request['source']}] -> [{request['dest']}]")
    import time

    LAST_EDIT_TIMING.clear()
    t = time.perf_counter()
    zs = _load_vstars(requests, hparams, cache_name, device)
    _tick(device, "vstar_npz_read_ms", t)
    deltas = _insert_loop(pipe.text_encoder, pipe.tokenizer, requests, hparams, hparams.layers, zs,
                          hparams.mom2_update_weight, stat_dir, verbose, refine_steps)
    print(f"Deltas successfully computed for {list(deltas.keys())}")
    return deltas


def _apply_deltas(model, deltas, device):
    with torch.no_grad():
        for w_name, (key_mat, val_mat) in deltas.items():
            key_mat, val_mat = key_mat.to(device), val_mat.to(device)
            upd_matrix = key_mat @ val_mat.T                                      # :805, fp64 on the device
            w = nethook.get_parameter(model, w_name)
            w[...] += upd_matrix_match_shape(upd_matrix, w.shape).float()


def apply_emcid_to_text_encoder(pipe, requests: List[Dict], hparams, device: str, mom2_weight: Optional[int] = None,
                                edit_weight: Optional[float] = None, return_orig_text_encoder=False,
                                cache_name: Optional[str] = None, stats_dir: Optional[str] = STATS_DIR,
                                verbose: bool = True):
    """Returns (pipe with edited text encoder, original text encoder or None)."""
    origin_text_encoder = None
    if return_orig_text_encoder:
        origin_text_encoder = deepcopy(pipe.text_encoder).to("cpu")
    deltas = execute_emcid_text_encoder(pipe, requests, hparams, cache_name=cache_name, mom2_weight=mom2_weight,
                                        edit_weight=edit_weight, verbose=verbose, stat_dir=stats_dir)
    _apply_deltas(pipe.text_encoder, deltas, device)
    print(f"New weights successfully inserted into {list(deltas.keys())}")
    if return_orig_text_encoder:
        origin_text_encoder = origin_text_encoder.to(device)
    return pipe, origin_text_encoder


def execute_emcid_sd_xl_text_encoders(pipe, requests: List[Dict], hparams, cache_name: Optional[str] = None,
                                      mom2_weight: Optional[int] = None, mom2_weight_2: Optional[int] = None,
                                      edit_weight: Optional[float] = None, verbose: bool = True,
                                      stat_dir="data/stats/sdxl/text1", stat_dir_2="data/stats/sdxl/text2",
                                      refine_steps: int = DEFAULT_REFINE_STEPS):
    """Both SDXL text encoders: `layers` / `mom2_update_weight` / stat_dir for text_encoder,
    `layers_2` / `mom2_update_weight_2` / stat_dir_2 for text_encoder_2 (v_star files end in _2)."""
    device = pipe.device
    hparams.mom2_update_weight = mom2_weight if mom2_weight is not None else hparams.mom2_update_weight
    hparams.mom2_update_weight_2 = mom2_weight_2 if mom2_weight_2 is not None else hparams.mom2_update_weight_2
    hparams.edit_weight = edit_weight if edit_weight is not None else hparams.edit_weight
    requests = deepcopy(requests)
    zs = _load_vstars(requests, hparams, cache_name, device)
    zs_2 = _load_vstars(requests, hparams, cache_name, device, suffix="_2")
    deltas = _insert_loop(pipe.text_encoder, pipe.tokenizer, requests, hparams, hparams.layers, zs,
                          hparams.mom2_update_weight, stat_dir, verbose, refine_steps)
    deltas_2 = _insert_loop(pipe.text_encoder_2, pipe.tokenizer_2, requests, hparams, hparams.layers_2, zs_2,
                            hparams.mom2_update_weight_2, stat_dir_2, verbose, refine_steps)
    print(f"Deltas successfully computed for {list(deltas.keys())} and {list(deltas_2.keys())}")
    return deltas, deltas_2


def apply_emcid_to_sdxl_text_encoders(pipe, requests: List[Dict], hparams, device: str,
                                      mom2_weight: Optional[int] = None, mom2_weight_2: Optional[int] = None,
                                      edit_weight: Optional[float] = None, return_orig_text_encoder=False,
                                      cache_name: Optional[str] = None, stat_dir: Optional[str] = XL_STATS_DIR1,
                                      stat_dir_2: Optional[str] = XL_STATS_DIR2, verbose: bool = True):
    origin_1 = origin_2 = None
    if return_orig_text_encoder:
        origin_1 = deepcopy(pipe.text_encoder).to("cpu")
        origin_2 = deepcopy(pipe.text_encoder_2).to("cpu")
    deltas, deltas_2 = execute_emcid_sd_xl_text_encoders(
        pipe, requests, hparams, cache_name=cache_name, mom2_weight=mom2_weight, mom2_weight_2=mom2_weight_2,
        edit_weight=edit_weight, verbose=verbose, stat_dir=stat_dir, stat_dir_2=stat_dir_2)
    _apply_deltas(pipe.text_encoder, deltas, device)
    _apply_deltas(pipe.text_encoder_2, deltas_2, device)
    print(f"New weights successfully inserted into {list(deltas.keys())}")
    if return_orig_text_encoder:
        origin_1, origin_2 = origin_1.to(device), origin_2.to(device)
    return pipe, origin_1, origin_2
