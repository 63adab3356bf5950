"""Caption batching for the statistics pass (host logic).

Mirrors the observable behaviour of dsets/stat_dataset.py in SilentView/EMCID — item layout
(:99-110), length-sorted sub-batching (:122-150), right padding with 0 (:153-163) and the masked
flatten (:166-172) — because those four pieces define WHICH token rows enter mom2 and `count`.
The B200 driver itself (`layer_stats.py`) re-batches captions into fixed-width blocks: with right
padding and CLIP's causal attention the fc2 input of a valid token does not depend on how captions
are batched, so only the *set* of valid tokens has to match, and it does.
"""
from __future__ import annotations

import json
import os
from typing import Dict, List, Sequence

import torch
from torch.utils.data import Dataset


class TokenizedDataset(Dataset):
    """Caption json -> {input_ids, position_ids, attention_mask} 1-D int64 tensors per item."""

    def __init__(self, data_path, tokenizer=None, maxlen=None):
        if not os.path.exists(data_path):
            raise FileNotFoundError(
                f"{data_path} not found (the reference downloads ccs_filtered.json here; this build has no network)")
        with open(data_path, "r") as f:
            records = json.load(f)
        self.data = [rec["caption"] for rec in records]
        self.tokenizer = tokenizer
        self.maxlen = maxlen

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        ids = self.tokenizer.encode(self.data[idx], truncation=True, max_length=self.maxlen)
        n = len(ids)
        return dict(input_ids=torch.tensor(ids), position_ids=torch.arange(n),
                    attention_mask=torch.ones(n, dtype=torch.long))


def dict_to_(data: Dict[str, torch.Tensor], device):
    for k in data:
        data[k] = data[k].to(device)
    return data


def make_padded_batch(items: Sequence[Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """Right-pad every field with zeros to the longest sequence; empty sequences are dropped."""
    keep = [it for it in items if len(it["input_ids"])]
    if not keep:
        return {k: torch.zeros((0, 0), dtype=torch.long) for k in items[0]}
    width = max(len(it["input_ids"]) for it in keep)
    out = {}
    for k in items[0]:
        buf = torch.zeros((len(keep), width), dtype=keep[0][k].dtype)
        for r, it in enumerate(keep):
            buf[r, : len(it[k])] = it[k]
        out[k] = buf
    return out


def length_collation(token_size: int):
    """collate_fn factory: sort by decreasing length, cut a new sub-batch whenever
    width * (rows + 1) would exceed token_size, pad each sub-batch (reference :122-150)."""

    def collate_fn(items):
        ordered = sorted(items, key=lambda it: -len(it["input_ids"]))
        groups: List[List[Dict[str, torch.Tensor]]] = []
        cur: List[Dict[str, torch.Tensor]] = []
        width = 0
        for it in ordered:
            n = len(it["input_ids"])
            if n == 0:
                break
            if width * (len(cur) + 1) > token_size:
                groups.append(cur)
                cur, width = [], 0
            if not cur:
                width = n
            cur.append(it)
        if cur:
            groups.append(cur)
        return [make_padded_batch(g) for g in groups]

    return collate_fn


def flatten_masked_batch(data: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Rows of data.view(-1, d) whose attention mask is non-zero (reference :166-172).  The fused
    kernels never call this (they mask on the h-wide MLP input); it exists for API parity."""
    flat = data.reshape(-1, data.size(-1))
    return flat[mask.reshape(-1).nonzero()[:, 0]]


def fixed_width_collation(width: int = 0):
    """collate_fn of the B200 driver: one right-padded block per DataLoader batch (`width` = 0 pads
    to the longest caption of the block)."""

    def collate_fn(items):
        batch = make_padded_batch(items)
        if width and batch["input_ids"].shape[1] < width:
            pad = width - batch["input_ids"].shape[1]
            batch = {k: torch.nn.functional.pad(v, (0, pad)) for k, v in batch.items()}
        return batch

    return collate_fn


def packed_collation():
    """collate_fn of the native-forward driver: no padding at all.  Captions are concatenated into packed
    int32 `packed_ids` / `packed_pos` with `cu_seqlens` prefix sums (what csrc/clip.cuh consumes); empty
    captions are dropped like make_padded_batch drops them.  Items whose attention mask is not all ones
    (never produced by TokenizedDataset, reference :99-110) make the block fall back to a padded batch."""

    def collate_fn(items):
        keep = [it for it in items if len(it["input_ids"])]
        if not keep:
            return make_padded_batch(items)
        masks = torch.cat([it["attention_mask"] for it in keep])
        if not bool((masks != 0).all()):
            return make_padded_batch(items)
        lens = torch.tensor([len(it["input_ids"]) for it in keep], dtype=torch.int32)
        cu = torch.zeros(len(keep) + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(lens, 0)
        return {"packed_ids": torch.cat([it["input_ids"] for it in keep]).to(torch.int32),
                "packed_pos": torch.cat([it["position_ids"] for it in keep]).to(torch.int32),
                "cu_seqlens": cu}

    return collate_fn


def unpack_to_padded(batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """packed block -> right-padded int64 input_ids / position_ids / attention_mask [B, L]."""
    cu = batch["cu_seqlens"].to(torch.int64).cpu()
    lens = cu[1:] - cu[:-1]
    B, L = len(lens), int(lens.max()) if len(lens) else 0
    cols = torch.arange(L)[None, :]
    mask = cols < lens[:, None]
    out = {}
    for src, dst in (("packed_ids", "input_ids"), ("packed_pos", "position_ids")):
        buf = torch.zeros(B, L, dtype=torch.int64)
        buf[mask] = batch[src].cpu().to(torch.int64)
        out[dst] = buf
    out["attention_mask"] = mask.to(torch.int64)
    return out


# 74 CTA pairs x 256-row tiles x 2: with this many packed tokens per block every linear layer of the forward fills
# whole waves of 256 x 256 pair tiles on a 148-SM B200 (37 888 / 256 = 148 row tiles; x 3, 9, 12 column tiles of
# CLIP-L or x 5, 15, 20 of OpenCLIP bigG are all multiples of 74).  Measured with 512 full-length captions
# (39 424 tokens, 154 row tiles): the N = 768 products (out projection, fc2) ran 7 waves for 6.24 waves of work.
DEFAULT_BLOCK_TOKENS = 37888


class PackedReblocker:
    """Regroups packed caption batches (packed_collation) into device blocks of at most `budget` tokens, cutting
    only between captions and keeping their order.  DataLoader batches are sized in captions; the native forward
    wants blocks sized in TOKENS (real captions are ragged: 256 of them can be 3 000 or 19 712 tokens)."""

    def __init__(self, budget: int = DEFAULT_BLOCK_TOKENS):
        self.budget = int(budget)
        self._ids: List[torch.Tensor] = []
        self._pos: List[torch.Tensor] = []
        self._lens: List[torch.Tensor] = []
        self._tokens = 0

    def _emit(self) -> Dict[str, torch.Tensor]:
        lens = torch.cat(self._lens)
        cu = torch.zeros(lens.numel() + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(lens, 0)
        out = {"packed_ids": torch.cat(self._ids), "packed_pos": torch.cat(self._pos), "cu_seqlens": cu}
        self._ids, self._pos, self._lens, self._tokens = [], [], [], 0
        return out

    def push(self, batch: Dict[str, torch.Tensor]):
        """Yields every block completed by `batch`."""
        ids, pos, cu = batch["packed_ids"], batch["packed_pos"], batch["cu_seqlens"].to(torch.int64)
        lens = (cu[1:] - cu[:-1]).to(torch.int32)
        start, n = 0, lens.numel()
        while start < n:
            room = self.budget - self._tokens
            # captions [start, stop) fit: largest stop with cu[stop] - cu[start] <= room
            stop = int(torch.searchsorted(cu, cu[start] + room, right=True)) - 1
            stop = min(stop, n)
            if stop <= start:
                if self._tokens:          # the block is full
                    yield self._emit()
                    continue
                stop = start + 1          # a single caption longer than the budget travels alone
            a, b = int(cu[start]), int(cu[stop])
            self._ids.append(ids[a:b]); self._pos.append(pos[a:b]); self._lens.append(lens[start:stop])
            self._tokens += b - a
            start = stop
            if self._tokens >= self.budget:
                yield self._emit()

    def flush(self):
        if self._tokens:
            yield self._emit()
