"""Shared helpers for the test-suite (test infrastructure: may import oracle/)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def unpack_captions(g):
    flat, offs = g["cap_flat"], g["cap_offs"]
    return [torch.from_numpy(flat[offs[i]: offs[i + 1]].copy()) for i in range(len(offs) - 1)]


def model_from_golden(g):
    """Rebuild the tiny CLIP text model stored in a fixture (weights included)."""
    model = rh.make_clip_text_model(str(g["kind"]), seed=int(g["seed_model"]))
    sd = model.state_dict()
    for k in g.files:
        if k.startswith("w."):
            sd[k[2:]] = torch.from_numpy(g[k])
    model.load_state_dict(sd)
    return model


def weight_checksum(model):
    return np.array([float(p.detach().double().abs().sum()) for p in model.parameters()][:32])
