"""TEST INFRASTRUCTURE ONLY — fixtures for the modules SURVEY.md §8 f4 names, produced by the UNMODIFIED reference
(/root/reference) on CPU in the build container:

    python oracle/gen_golden_f4.py

  tests/golden/tiny_cross_attn.npz   layer_stats_cross_attn_kv (emcid/layer_stats.py:333-427) for two K/V modules of a
                                     miniature UNet, then execute_emcid_cross_attn + apply_emcid_to_cross_attn
                                     (emcid/emcid_main.py:314-547) on all of its K/V modules
  tests/golden/tiny_clip_model.npz   execute_emcid_clip + apply_emcid_to_clip (emcid/emcid_main.py:109-311) on a tiny
                                     transformers.CLIPModel
The stand-in models are rebuilt from their seeds at test time (oracle/ref_harness.py); the fixtures carry the weights
that matter so that a changed RNG stream is detected."""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402
from oracle.gen_golden import write_stats_npz  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def kv_names(pipe):
    ref = rh.import_reference()
    return ref.layer_stats.get_all_cross_attn_kv_layer_names(pipe)


def get(obj, name):
    for part in name.split("."):
        obj = getattr(obj, part)
    return obj


def tiny_cross_attn(out, n_caps=90, sample_size=64, n_req=6, lam=4000, edit_weight=0.5):
    pipe = rh.make_cross_attn_pipe(seed=0)
    caps = rh.make_captions(n_caps, 1000, seed=13)
    names = kv_names(pipe)
    tmp = tempfile.mkdtemp()
    try:
        stats_dir = os.path.join(tmp, "stats")
        d = dict(names=np.array(names), sample_size=sample_size, n_caps=n_caps, seed_caps=13, n_req=n_req, lam=lam,
                 edit_weight=edit_weight)
        # statistics: the reference runs one pass per module; two of them are enough to show they are the same matrix
        for n in (names[0], names[-1]):
            st = rh.run_reference_cross_attn_stats(pipe, caps, n, stats_dir, sample_size)
            d[f"mom2.{n}"] = st.mom2.mom2.numpy().copy()
            d[f"count.{n}"] = st.mom2.count
        files = sorted(os.path.relpath(os.path.join(dp, f), stats_dir) for dp, _, fs in os.walk(stats_dir) for f in fs)
        d["stat_files"] = np.array(files)
        reqs = rh.make_requests(n_req)
        cache = os.path.join(tmp, "v", "c_")
        rh.write_cross_attn_vstar_cache(cache, reqs, pipe, names, seed=2)
        hp = rh.make_hparams([0], sample_size, mom2_update_weight=lam, edit_weight=edit_weight)
        for n in names:
            d[f"w_before.{n}"] = get(pipe.unet, n).weight.detach().numpy().copy()
        deltas = rh.run_reference_cross_attn_edit(pipe, reqs, hp, cache, stats_dir, apply=False)
        assert list(deltas) == [f"{n}.weight" for n in names]
        for n in names:
            assert np.array_equal(get(pipe.unet, n).weight.numpy(), d[f"w_before.{n}"])          # restored
            d[f"adj_k.{n}"] = deltas[f"{n}.weight"][0].numpy()
            d[f"resid.{n}"] = deltas[f"{n}.weight"][1].numpy()
        rh.run_reference_cross_attn_edit(pipe, reqs, hp, cache, stats_dir, apply=True)
        for n in names:
            d[f"w_after.{n}"] = get(pipe.unet, n).weight.detach().numpy().copy()
        d["text_fc2_checksum"] = float(pipe.text_encoder.text_model.encoder.layers[1].mlp.fc2.weight.double().abs().sum())
        np.savez_compressed(os.path.join(GOLD, out), **d)
        print(out, len(names), "modules; stats files:", files[:2], "...")
    finally:
        shutil.rmtree(tmp)


def tiny_clip_model(out, n_caps=150, sample_size=120, n_req=8, layers=(0, 1), lam=4000, edit_weight=0.5):
    model = rh.make_clip_model(seed=5)
    tower = rh.text_tower_of(model)
    caps = rh.make_captions(n_caps, 1000, seed=3)
    tok = rh.FakeTokenizer(1000)
    reqs = rh.make_requests(n_req)
    tmp = tempfile.mkdtemp()
    try:
        stats_dir = os.path.join(tmp, "stats")
        d = dict(layers=np.array(layers), sample_size=sample_size, n_caps=n_caps, seed_caps=3, n_req=n_req, lam=lam,
                 edit_weight=edit_weight)
        for l in layers:
            st = rh.run_reference_layer_stats(tower, caps, l, stats_dir, sample_size)
            d[f"mom2.{l}"] = st.mom2.mom2.numpy().copy()
            d[f"count.{l}"] = st.mom2.count
            d[f"w_before.{l}"] = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().numpy().copy()
        cache = os.path.join(tmp, "v", "c_")
        d["zs"] = rh.write_vstar_cache(cache, reqs, 64, seed=2).numpy()
        hp = rh.make_hparams(layers, sample_size, mom2_update_weight=lam, edit_weight=edit_weight)
        deltas = rh.run_reference_clip_edit(model, tok, reqs, hp, cache, stats_dir, apply=False)
        for l in layers:
            name = f"text_model.encoder.layers.{l}.mlp.fc2.weight"
            assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), d[f"w_before.{l}"])
            d[f"adj_k.{l}"] = deltas[name][0].numpy()
            d[f"resid.{l}"] = deltas[name][1].numpy()
        rh.run_reference_clip_edit(model, tok, reqs, hp, cache, stats_dir, apply=True)
        for l in layers:
            d[f"w_after.{l}"] = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().numpy().copy()
        np.savez_compressed(os.path.join(GOLD, out), **d)
        print(out, "layers", layers)
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    assert rh.reference_available(), "run in the build container (needs /root/reference)"
    torch.set_num_threads(os.cpu_count())
    tiny_cross_attn("tiny_cross_attn.npz")
    tiny_clip_model("tiny_clip_model.npz")
