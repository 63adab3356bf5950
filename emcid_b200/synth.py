"""Synthetic inputs of the shapes BASELINE.json names: random-init HF CLIP text
encoders and seeded token-id captions.  There is no network in the build/bench environment, so
neither SD checkpoints nor ccs_filtered.json exist; every benchmark says "synthetic"."""
from __future__ import annotations

from typing import List

import torch

ENCODERS = {
    # SD-v1.4 text encoder = SDXL text encoder 1 (CLIP ViT-L/14 text tower)
    "sd-text": dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                    num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu"),
    # SDXL text encoder 2 (OpenCLIP bigG text tower)
    "sdxl-text2": dict(vocab_size=49408, hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                       num_attention_heads=20, max_position_embeddings=77, hidden_act="gelu"),
}
ENCODERS["sdxl-text1"] = ENCODERS["sd-text"]


def make_text_encoder(kind: str = "sd-text", seed: int = 0):
    from transformers import CLIPTextConfig, CLIPTextModel

    cfg = dict(ENCODERS[kind])
    cfg["bos_token_id"] = cfg["vocab_size"] - 2
    cfg["eos_token_id"] = cfg["vocab_size"] - 1
    torch.manual_seed(seed)
    model = CLIPTextModel(CLIPTextConfig(**cfg)).eval()
    for p in model.parameters():
        p.requires_grad_(False)
    model.config._name_or_path = f"synthetic/{kind}-seed{seed}"
    return model


def make_caption_ids(n: int, vocab: int = 49408, seed: int = 0, width: int = 77, full: bool = True,
                     min_len: int = 8) -> List[torch.Tensor]:
    """BOS first, EOS last, ids ~ U{0..vocab-3}; `full` = every caption is `width` tokens (the
    throughput set), otherwise len ~ U{min_len..width} (the parity / masking set)."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.full((n,), width) if full else torch.randint(min_len, width + 1, (n,), generator=g)
    ids = torch.randint(0, vocab - 2, (n, width), generator=g)
    ids[:, 0] = vocab - 2
    out = []
    for i in range(n):
        L = int(lens[i])
        row = ids[i, :L].clone()
        row[L - 1] = vocab - 1
        out.append(row)
    return out


class CaptionIdDataset(torch.utils.data.Dataset):
    """Items laid out like dsets/stat_dataset.py::TokenizedDataset (:99-110) of the reference."""

    def __init__(self, captions):
        self.captions = captions
        self._pos, self._ones = {}, {}      # per-length position ids / masks are shared (collation only reads them)

    def __len__(self):
        return len(self.captions)

    def __getitem__(self, i):
        ids = self.captions[i]
        n = len(ids)
        if n not in self._pos:
            self._pos[n] = torch.arange(n)
            self._ones[n] = torch.ones(n, dtype=torch.long)
        return dict(input_ids=ids, position_ids=self._pos[n], attention_mask=self._ones[n])


class CaptionMatrixDataset(torch.utils.data.Dataset):
    """Full-width captions held as one [n, width] int64 matrix (the throughput set of SURVEY.md §8d): same item
    layout as CaptionIdDataset without materialising n row tensors up front."""

    def __init__(self, ids: torch.Tensor):
        assert ids.dim() == 2 and ids.dtype == torch.int64
        self.ids = ids
        n = ids.shape[1]
        self._pos, self._ones = torch.arange(n), torch.ones(n, dtype=torch.long)

    def __len__(self):
        return self.ids.shape[0]

    def __getitem__(self, i):
        return dict(input_ids=self.ids[i], position_ids=self._pos, attention_mask=self._ones)


def make_caption_matrix(n: int, vocab: int = 49408, seed: int = 0, width: int = 77) -> torch.Tensor:
    """[n, width] int64: BOS first, EOS last, ids ~ U{0..vocab-3} (every caption full width)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab - 2, (n, width), generator=g)
    ids[:, 0] = vocab - 2
    ids[:, -1] = vocab - 1
    return ids


# ---------------------------------------------------------------------------------------------- edit workload (config 3)
class WordHashTokenizer:
    """Offline stand-in for the CLIP tokenizer (no vocabulary files exist in the build/bench environment): whitespace
    words hashed into the vocabulary, BOS first, EOS last, right padding with the EOS id like CLIP's tokenizer.  Implements
    what tokenize_prompts / find_token_range call: __call__(prompts, return_tensors, padding, truncation[, max_length]),
    decode(ids), model_max_length."""

    model_max_length = 77

    def __init__(self, vocab_size: int = 49408):
        self.vocab_size = vocab_size
        self.bos, self.eos = vocab_size - 2, vocab_size - 1
        self._words = {}   # id -> word
        self._ids = {}     # word -> id

    def _id(self, word: str) -> int:
        import zlib

        i = self._ids.get(word)
        if i is None:
            i = zlib.crc32(word.encode()) % (self.vocab_size - 2)
            while i in self._words:          # linear probing: ids stay unique, so decode() inverts encode()
                i = (i + 1) % (self.vocab_size - 2)
            self._words[i] = word
            self._ids[word] = i
        return i

    def encode(self, text, truncation=True, max_length=None):
        ids = [self.bos] + [self._id(w) for w in text.split()] + [self.eos]
        limit = max_length or self.model_max_length
        return ids if len(ids) <= limit else ids[: limit - 1] + [self.eos]

    def __call__(self, prompts, return_tensors="pt", padding=True, truncation=True, max_length=None):
        prompts = [prompts] if isinstance(prompts, str) else prompts
        enc = [self.encode(p, max_length=max_length if truncation else None) for p in prompts]
        width = max(len(e) for e in enc) if padding is True else (max_length or self.model_max_length)
        enc = [e[:width] for e in enc]
        ids = torch.tensor([e + [self.eos] * (width - len(e)) for e in enc], dtype=torch.long)
        mask = torch.tensor([[1] * len(e) + [0] * (width - len(e)) for e in enc], dtype=torch.long)
        return {"input_ids": ids, "attention_mask": mask}

    def decode(self, ids):
        if torch.is_tensor(ids):
            ids = ids.tolist()
        ids = [ids] if isinstance(ids, int) else ids
        return " ".join(self._piece(int(i)) for i in ids)

    def _piece(self, i: int) -> str:
        if i == self.bos:
            return "<|startoftext|>"
        if i == self.eos:
            return "<|endoftext|>"
        return self._words.get(i) or f"w{i}"


def make_edit_requests(n: int):
    """n requests with the three ICEB edit templates (dsets/iceb_dataset.py:325-329), two-word subjects."""
    return [{"source": f"artist{i} name{i}", "dest": "art", "seed_train": 0,
             "prompts": ["An image of {}", "A photo of {}", "{}"]} for i in range(n)]


def make_edit_hparams(layers, mom2_n_samples: int, mom2_update_weight: float = 4000, edit_weight: float = 0.5):
    """The fields of emcid/emcid_hparams.py::EMCIDHyperParams (:55-163) the stage-2 loop reads."""
    from types import SimpleNamespace

    return SimpleNamespace(layers=list(layers), mom2_update_weight=mom2_update_weight, edit_weight=edit_weight,
                           rewrite_module_tmp="text_model.encoder.layers.{}.mlp.fc2", mom2_dataset="ccs_filtered",
                           mom2_n_samples=mom2_n_samples, mom2_dtype="float32", num_edit_tokens=1, objective="ori",
                           sld_supervision=False, use_new_compute_z=False)


def write_vstar_cache(cache_name: str, requests, h: int, seed: int = 2):
    """v* ~ N(0, 1) for every request, in the reference's cache layout (emcid_main.py:886-901)."""
    import os

    import numpy as np

    g = torch.Generator().manual_seed(seed)
    os.makedirs(os.path.dirname(cache_name) or ".", exist_ok=True)
    for r in requests:
        np.savez(cache_name + f"source_{r['source']}_dest_{r['dest']}.npz", v_star=torch.randn(h, generator=g).numpy())
