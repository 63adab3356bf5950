"""Caller side of the update: keys and current outputs at the edited words.

Behavioural mirror of emcid/compute_z.py::get_module_input_output_at_words (:2252-2384),
tokenize_prompts (:56-74), emcid/compute_ks.py::compute_ks_text_encoder (:21-41) and
experiments/causal_trace.py::find_token_range (:1057-1103).  One forward yields both the fc2 input and output, so callers that need both (the edit loop does,
reference emcid_main.py:987-1014) pay for one pass instead of two.  For a plain fp32 HF CLIP text tower on a
CUDA device — CLIPTextModel(WithProjection) or the text half of a whole CLIPModel — that forward runs in the library
(emcid_clip_forward_keys: packed prompts, 3xFP16 tcgen05 GEMMs, fc2 evaluated on the looked-up rows only); the edit
loop's in-place weight writes reach the library through `Tensor._version` and a per-edit content checksum
(clip_forward.NativeClipTextEncoder.sync_weights).  Everything else (several edit tokens per prompt, non-right-padded
masks, look-ups beyond a prompt's valid tokens, EMCID_NATIVE_KEYS=0) keeps the traced HF forward.
"""
from __future__ import annotations

import itertools
import re
import unicodedata
from typing import Dict, List, Optional, Tuple

import torch

from . import clip_forward, nethook

_SERIAL = itertools.count(1)   # identity of a prepared prompt set (continuation token of the native key extraction)
_FC2_NAME = re.compile(r"^(?:text_model\.)?encoder\.layers\.(\d+)\.mlp\.fc2$")
LAST_PATH = {"native": False}   # which path served the last call (tests / bench read it)
_PACKED: Dict[str, tuple] = {}  # packed prompts + looked-up rows of the last prepared prompt set


def _native_keys(text_encoder, enc: Dict[str, torch.Tensor], lookup, module_name: str, token=None):
    """(fc2 input [P, d], fc2 output [P, h]) at lookup[p][0] of every prompt through the library, or None when
    this model / module / mask is not covered.  `token` identifies the prompt set: consecutive calls with the same
    token at increasing layers, between which only fc2 of the previous layer changed (the edit loop), continue from
    the previous call's state instead of re-running the layers below.  `lookup=None` only LAUNCHES the forward up to
    act(fc1) of the layer (see prefetch_keys): the call that follows with the looked-up positions, same token and layer,
    gathers its rows from the state this one leaves behind."""
    m = _FC2_NAME.match(module_name)
    if m is None or not text_encoder.device.type == "cuda":
        return None
    layer = int(m.group(1))
    tm = getattr(text_encoder, "text_model", text_encoder)
    if not hasattr(tm, "encoder") or layer >= len(tm.encoder.layers):
        return None
    if not clip_forward.supports(text_encoder):
        return None
    max_pos = tm.embeddings.position_embedding.weight.shape[0]
    # the packed form of the prompts and the looked-up rows depend on the prompt set only: once per edit, not per layer
    hit = _PACKED.get("last")
    if token is None or hit is None or hit[0] != token or hit[1] is not enc["input_ids"]:
        packed = clip_forward.pack_batch({"input_ids": enc["input_ids"], "attention_mask": enc["attention_mask"]}, max_pos)
        hit = _PACKED["last"] = [token, enc["input_ids"], packed, None, False]       # ..., rows, rows looked up
    packed = hit[2]
    if packed is None or packed[4] <= 0:
        return None
    ids, pos, cu, S, T = packed
    if lookup is None:
        rows = torch.zeros(1, dtype=torch.int32, device=cu.device)
    else:
        if not hit[4]:
            first = torch.tensor([row[0] for row in lookup], dtype=torch.int32, device=cu.device)
            if bool((first < cu[1:] - cu[:-1]).all()):
                hit[3] = cu[:-1] + first
            # else: a looked-up position beyond a prompt's valid tokens (subjects "", " ", "[EOS]" with ragged prompts: the
            # LAST COLUMN of the padded batch, causal_trace.py:1063-1064): that pad row does not exist in the packed forward
            hit[4] = True
        rows = hit[3]
        if rows is None:
            return None
    prev_token = getattr(_native_keys, "_last_token", None)
    # once per edit (a new prompt set): compare content checksums too — writes through `.data` do not bump `_version`
    native = clip_forward.key_encoder(text_encoder, T, S, layer, verify=token is None or token != prev_token)
    _native_keys._last_token = token
    if native is None:
        return None
    resume = -1
    prev = native.keys_token
    if token is not None and prev is not None and prev[0] == token and prev[1] <= layer and prev[2] == T:
        changed = native.last_sync
        # fc2.weight / fc2.bias are source tensors 14 and 15 of a layer (NativeClipTextEncoder._layer_tensors)
        # the saved state depends on the embeddings, on layers below prev[1] and on everything of layer prev[1] but its fc2
        if all(isinstance(k, int) and (k > prev[1] or (k == prev[1] and set(v) <= {14, 15})) for k, v in changed.items()):
            resume = prev[1]
    if lookup is not None:
        LAST_PATH["resumed_from"] = resume
    out = native.forward_keys(ids, pos, cu, S, T, layer, rows, resume_layer=resume)
    native.keys_token = (token, layer, T) if token is not None else None
    return out


def prefetch_keys(text_encoder, enc, module_name: str, token) -> bool:
    """Launches the library forward of the tokenised prompts up to act(fc1) of `module_name`'s layer BEFORE the positions
    of the subject tokens are known: the host-side search for them (find_token_range over every prompt, 7 ms per 1000
    requests) then runs beside the device's pass through the layers below the first edited one instead of in front of it.
    The keys call that follows (same token and layer) only gathers its rows.  False when the library does not cover
    this model / module / mask — nothing was launched and the caller's path is unchanged."""
    with torch.no_grad():
        return _native_keys(text_encoder, enc, None, module_name, token=token) is not None


def last_hidden_at_words(text_encoder, tok, requests: List[Dict]) -> torch.Tensor:
    """[n, hidden]: the text encoder's last_hidden_state at the last subject token of every source prompt, averaged over
    each request's prompts — the input every UNet cross-attention to_k / to_v sees at that token
    (emcid/compute_ks.py:52-139, where it is read back through a traced UNet forward).  Library forward for a plain fp32
    CLIP text tower on a CUDA device, HF forward otherwise."""
    device = text_encoder.device
    enc, lookup, counts, _serial = prepare_lookup(tok, requests, 1, device)
    assert len(set(counts)) == 1, "All the requests should have the same number of prompts."       # reference :66-67, :78-79
    rows_out = None
    if device.type == "cuda" and clip_forward.supports(text_encoder):
        tm = getattr(text_encoder, "text_model", text_encoder)
        packed = clip_forward.pack_batch({"input_ids": enc["input_ids"], "attention_mask": enc["attention_mask"]},
                                         tm.embeddings.position_embedding.weight.shape[0])
        if packed is not None and packed[4] > 0:
            ids, pos, cu, S, T = packed
            first = torch.tensor([row[0] for row in lookup], dtype=torch.int32, device=cu.device)
            native = clip_forward.key_encoder(text_encoder, T, S, 10 ** 6, verify=True)
            if native is not None and native.has_final_norm and bool((first < cu[1:] - cu[:-1]).all()):
                rows_out = native.forward_final(ids, pos, cu, S, T, rows=cu[:-1] + first)
    LAST_PATH["native"] = rows_out is not None
    if rows_out is None:
        with torch.no_grad():
            hidden = text_encoder(**enc)[0]
        rows_out = hidden[torch.arange(len(lookup), device=device), torch.tensor([r[0] for r in lookup], device=device)]
    c = counts[0]
    return rows_out.reshape(len(counts), c, -1).mean(1)


def _tokenize_on_host(prompts, tokenizer, padding_length=None):
    if padding_length is None:
        return tokenizer(prompts, return_tensors="pt", padding=True, truncation=True)
    return tokenizer(prompts, return_tensors="pt", padding="max_length", truncation=True, max_length=padding_length)


def tokenize_prompts(prompts, tokenizer, device, padding_length=None):
    return {k: v.to(device) for k, v in _tokenize_on_host(prompts, tokenizer, padding_length).items()}


def find_token_range(tokenizer, token_array, substring_orig: str, piece_cache: Optional[Dict[int, str]] = None) -> Tuple[int, int]:
    """[start, end) of the tokens that spell `substring_orig` inside `token_array`.  `piece_cache` memoises the
    per-token decodes across the prompts of one edit (thousands of prompts share a few hundred distinct tokens)."""
    if substring_orig == "[CLS]":
        return (0, 1)
    if substring_orig in ("[EOS]", "", " "):
        return (len(token_array) - 1, len(token_array))
    needle = substring_orig.replace(" ", "").lower()
    if piece_cache is None:
        pieces = [tokenizer.decode([t]) for t in token_array]
        haystack = tokenizer.decode(token_array).replace(" ", "")
    else:
        pieces = []
        for t in token_array:
            t = int(t)
            piece = piece_cache.get(t)
            if piece is None:
                piece = piece_cache[t] = tokenizer.decode([t])
            pieces.append(piece)
        joined = "".join(pieces)
        # The reference decodes the whole array (causal_trace.py:1072-1075).  When every token decodes to ASCII text the
        # concatenated per-token decodes are the same string once the spaces are dropped (byte-level BPE only differs
        # where a multi-byte character is split across tokens), so the second, whole-array decode per prompt — the
        # bulk of this function's time with a Python tokenizer — is only paid for non-ASCII prompts.
        haystack = joined.replace(" ", "") if joined.isascii() else tokenizer.decode(token_array).replace(" ", "")
    if "’" in needle:
        haystack = haystack.replace("'", "’")
    if not (haystack.isascii() and needle.isascii()):     # NFKC is the identity on ASCII
        haystack = unicodedata.normalize("NFKC", haystack)
        needle = unicodedata.normalize("NFKC", needle)
    try:
        char_loc = haystack.index(needle)
    except ValueError:
        print("Cannot find substring in tokens")
        print("substring: ", needle)
        print("whole string: ", haystack)
        raise ValueError
    seen, start, end = 0, None, None
    quirk = "ń" in needle                                   # reference quirk: 2 tokens, 1 char
    char_end = char_loc + len(needle)
    for i, piece in enumerate(pieces):
        if not (quirk and int(token_array[i]) == 78):
            seen += len(piece)
        if start is None and seen > char_loc:
            start = i
        if seen >= char_end:
            end = i + 1
            break
    return (start, end)


def _prompts_and_subjects(requests: List[Dict]):
    key = "source_prompts" if "source_prompts" in requests[0] else "prompts"
    if key == "source_prompts":
        prompts = [p for r in requests for p in r["source_prompts"]]
    else:
        prompts = [p.format(r["source"]) for r in requests for p in r["prompts"]]
    subjects = [r["source"] for r in requests for _ in r[key]]
    counts = [len(r[key]) for r in requests]
    return prompts, subjects, counts


def prepare_lookup(tok, requests: List[Dict], num_fact_token: int, device, after_tokenise=None):
    """Tokenised source prompts, the looked-up token positions of every prompt and the prompts-per-request counts
    (compute_z.py:2284-2300).  Independent of the layer: the edit loop computes it once per edit.  `after_tokenise(enc,
    serial)` (one looked-up token per prompt only) is called between the tokenisation and the look-up, see prefetch_keys."""
    prompts, subjects, counts = _prompts_and_subjects(requests)
    host = _tokenize_on_host(prompts, tok)
    enc = {k: v.to(device) for k, v in host.items()}
    serial = next(_SERIAL)
    if after_tokenise is not None and num_fact_token == 1:
        after_tokenise(enc, serial)
    # find_token_range walks the token ids one by one: host lists, taken from the tokenizer's own host tensors (reading
    # them back from the device would wait for whatever after_tokenise has just launched there)
    pieces: Dict[int, str] = {}
    if num_fact_token == 1:
        lookup = [[find_token_range(tok, ids, w, pieces)[-1] - 1] for ids, w in zip(host["input_ids"].tolist(), subjects)]
    else:
        extra = num_fact_token - 2
        host = _tokenize_on_host(prompts, tok, padding_length=len(host["input_ids"][0]) + extra)
        enc = {k: v.to(device) for k, v in host.items()}
        lookup = []
        for ids, w, n_valid in zip(host["input_ids"].tolist(), subjects, host["attention_mask"].sum(1).tolist()):
            eos = int(n_valid) - 1
            lookup.append([find_token_range(tok, ids, w, pieces)[-1] - 1] + list(range(eos, eos + extra + 1)))
    return enc, lookup, counts, serial


def get_module_input_output_at_words(text_encoder, tok, requests: List[Dict], module_name: str,
                                     num_fact_token: int = 1, prepared=None):
    """(input, output) of `module_name` at the last subject token of every source prompt, averaged
    over each request's prompts: [n, d] and [n, h]  (num_fact_token == 1) or with an extra
    token dimension [n, num_fact_token, ·] (last subject token, EOS, then padding positions)."""
    device = text_encoder.device
    enc, lookup, counts, serial = prepared if prepared is not None else prepare_lookup(tok, requests, num_fact_token, device)
    assert len(enc["input_ids"]) == len(lookup)
    native = None
    if num_fact_token == 1:
        with torch.no_grad():
            native = _native_keys(text_encoder, enc, lookup, module_name,
                                  token=serial if prepared is not None else None)
    LAST_PATH["native"] = native is not None
    if native is not None:
        l_in, l_out = native[0][:, None, :], native[1][:, None, :]      # [P, 1, d], [P, 1, h]
    else:
        with torch.no_grad(), nethook.TraceDict(text_encoder, [module_name], retain_input=True, retain_output=True) as td:
            if type(text_encoder).__name__ == "CLIPModel":
                text_encoder.get_text_features(**enc)
            else:
                text_encoder(**enc)
            idx = torch.tensor(lookup, device=device)                        # [P, F]
            rows = torch.arange(len(lookup), device=device)[:, None]
            l_in = td[module_name].input[rows, idx].detach().clone()         # [P, F, d]
            l_out = td[module_name].output[rows, idx].detach().clone()       # [P, F, h]
    if len(set(counts)) == 1:
        # every request has the same number of prompts (ICEB / artist templates): one reduction instead of n
        c = counts[0]
        ins = l_in.reshape(len(counts), c, *l_in.shape[1:]).mean(1)
        outs = l_out.reshape(len(counts), c, *l_out.shape[1:]).mean(1)
    else:
        ins, outs, at = [], [], 0
        for c in counts:
            ins.append(l_in[at: at + c].mean(0))
            outs.append(l_out[at: at + c].mean(0))
            at += c
        ins, outs = torch.stack(ins, 0), torch.stack(outs, 0)
    if num_fact_token == 1:
        ins, outs = ins[:, 0], outs[:, 0]
    return ins, outs


def compute_ks_text_encoder(model, tok, requests: List[Dict], hparams, layer: int):
    return get_module_input_output_at_words(model, tok, requests, hparams.rewrite_module_tmp.format(layer),
                                            num_fact_token=hparams.num_edit_tokens)[0]
