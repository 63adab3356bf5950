"""Caller side of the update: keys and current outputs at the edited words.

Behavioural mirror of emcid/compute_z.py::get_module_input_output_at_words (:2252-2384),
tokenize_prompts (:56-74), emcid/compute_ks.py::compute_ks_text_encoder (:21-41) and
experiments/causal_trace.py::find_token_range (:1057-1103).  The forward itself stays in PyTorch
(HF CLIP); SURVEY.md §8(a9) keeps this producer on the caller side of the hot path.  One forward
yields both the fc2 input and output, so callers that need both (the edit loop does, reference
emcid_main.py:987-1014) pay for one pass instead of two.
"""
from __future__ import annotations

import unicodedata
from typing import Dict, List, Tuple

import torch

from . import nethook


def tokenize_prompts(prompts, tokenizer, device, padding_length=None):
    if padding_length is None:
        enc = tokenizer(prompts, return_tensors="pt", padding=True, truncation=True)
    else:
        enc = tokenizer(prompts, return_tensors="pt", padding="max_length", truncation=True,
                        max_length=padding_length)
    return {k: v.to(device) for k, v in enc.items()}


def find_token_range(tokenizer, token_array, substring_orig: str) -> Tuple[int, int]:
    """[start, end) of the tokens that spell `substring_orig` inside `token_array`."""
    if substring_orig == "[CLS]":
        return (0, 1)
    if substring_orig in ("[EOS]", "", " "):
        return (len(token_array) - 1, len(token_array))
    needle = substring_orig.replace(" ", "").lower()
    pieces = [tokenizer.decode([t]) for t in token_array]
    haystack = tokenizer.decode(token_array).replace(" ", "")
    if "’" in needle:
        haystack = haystack.replace("'", "’")
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
    for i, piece in enumerate(pieces):
        if not ("ń" in needle and int(token_array[i]) == 78):  # reference quirk: 2 tokens, 1 char
            seen += len(piece)
        if start is None and seen > char_loc:
            start = i
        if end is None and seen >= char_loc + len(needle):
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


def get_module_input_output_at_words(text_encoder, tok, requests: List[Dict], module_name: str,
                                     num_fact_token: int = 1):
    """(input, output) of `module_name` at the last subject token of every source prompt, averaged
    over each request's prompts: [n, d] and [n, h]  (num_fact_token == 1) or with an extra
    token dimension [n, num_fact_token, ·] (last subject token, EOS, then padding positions)."""
    device = text_encoder.device
    prompts, subjects, counts = _prompts_and_subjects(requests)
    enc = tokenize_prompts(prompts, tok, device)
    if num_fact_token == 1:
        lookup = [[find_token_range(tok, ids, w)[-1] - 1] for ids, w in zip(enc["input_ids"], subjects)]
    else:
        extra = num_fact_token - 2
        enc = tokenize_prompts(prompts, tok, device, padding_length=len(enc["input_ids"][0]) + extra)
        lookup = []
        for ids, w, mask in zip(enc["input_ids"], subjects, enc["attention_mask"]):
            eos = int(mask.sum()) - 1
            lookup.append([find_token_range(tok, ids, w)[-1] - 1] + list(range(eos, eos + extra + 1)))
    assert len(enc["input_ids"]) == len(lookup)
    with torch.no_grad(), nethook.TraceDict(text_encoder, [module_name], retain_input=True, retain_output=True) as td:
        if type(text_encoder).__name__ == "CLIPModel":
            text_encoder.get_text_features(**enc)
        else:
            text_encoder(**enc)
        idx = torch.tensor(lookup, device=device)                        # [P, F]
        rows = torch.arange(len(lookup), device=device)[:, None]
        l_in = td[module_name].input[rows, idx].detach().clone()         # [P, F, d]
        l_out = td[module_name].output[rows, idx].detach().clone()       # [P, F, h]
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
