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
