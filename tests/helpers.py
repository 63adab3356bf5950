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


def padded_batch(caps):
    """Right-padded input_ids / position_ids / attention_mask of a caption list (dsets/stat_dataset.py:153-163)."""
    B, L = len(caps), max(len(c) for c in caps)
    ids = torch.zeros(B, L, dtype=torch.long)
    pos = torch.zeros(B, L, dtype=torch.long)
    mask = torch.zeros(B, L, dtype=torch.long)
    for i, c in enumerate(caps):
        ids[i, :len(c)] = c
        pos[i, :len(c)] = torch.arange(len(c))
        mask[i, :len(c)] = 1
    return {"input_ids": ids, "position_ids": pos, "attention_mask": mask}


def fp64_gram_reference(model, caps, layers, dev, chunk=512, probe=None):
    """Ground truth for the statistics pass at sizes the CPU oracle cannot reach: an fp64 copy of the HF model runs the
    reference's forward (emcid/layer_stats.py:208-219) on `dev`, the fc2 inputs of `layers` are flattened by the attention
    mask (dsets/stat_dataset.py:166-172) and accumulated as G = sum a a^T in fp64 (util/runningstats.py:493 without the fp32
    rounding).  Returns ({layer: G [d, d] fp64 on dev}, count); with `probe` = {layer: v [d] fp64} only G v is formed
    ({layer: [d] fp64}), which needs no d x d product."""
    import copy

    m64 = copy.deepcopy(model).double().to(dev)
    feats = {}
    hooks = [m64.text_model.encoder.layers[l].mlp.fc2.register_forward_pre_hook(
        lambda m, a, l=l: feats.__setitem__(l, a[0])) for l in layers]
    out, count = {}, 0
    try:
        with torch.no_grad():
            for c0 in range(0, len(caps), chunk):
                batch = {k: v.to(dev) for k, v in padded_batch(caps[c0: c0 + chunk]).items()}
                keep = batch["attention_mask"].bool()
                m64(**batch)
                count += int(keep.sum())
                for l in layers:
                    a = feats.pop(l)[keep]
                    if probe is not None:
                        g = a.T @ (a @ probe[l].to(dev))
                    else:
                        g = a.T @ a
                    out[l] = g if l not in out else out[l] + g
                    del a, g
    finally:
        for h in hooks:
            h.remove()
    del m64
    return out, count
