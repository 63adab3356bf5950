"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, SilentView/EMCID) on CPU in the build container.  Re-run with
    python oracle/gen_golden.py
Fixtures are small (tiny CLIP configs carry their weights; CLIP-L fixtures are digests and carry a
weight checksum instead, the model is rebuilt from its seed)."""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def pack_captions(caps):
    flat = np.concatenate([c.numpy() for c in caps]).astype(np.int64)
    offs = np.cumsum([0] + [len(c) for c in caps]).astype(np.int64)
    return flat, offs


def weights_of(model):
    return {"w." + k: v.detach().numpy() for k, v in model.state_dict().items() if v.dtype.is_floating_point}


def checksum(model):
    return np.array([float(p.detach().double().abs().sum()) for p in model.parameters()][:32])


def tiny_stats(kind, layer, n_caps, sample_size, batch_tokens, seed_model, seed_caps, out):
    model = rh.make_clip_text_model(kind, seed=seed_model)
    caps = rh.make_captions(n_caps, model.config.vocab_size, seed=seed_caps)
    tmp = tempfile.mkdtemp()
    try:
        stat = rh.run_reference_layer_stats(model, caps, layer, tmp, sample_size, batch_tokens=batch_tokens)
        files = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp) for f in fs]
        assert len(files) == 1
        saved = np.load(files[0])
        flat, offs = pack_captions(caps)
        np.savez_compressed(
            os.path.join(GOLD, out), kind=kind, layer=layer, sample_size=(-1 if sample_size is None else sample_size),
            batch_tokens=batch_tokens,
            seed_model=seed_model, cap_flat=flat, cap_offs=offs, rel_path=os.path.relpath(files[0], tmp),
            **{"npz." + k: saved[k] for k in saved.files}, **weights_of(model))
        print(out, "count", int(saved["mom2.count"]), "keys", saved.files)
    finally:
        shutil.rmtree(tmp)
    return model, caps


def write_stats_npz(stats_dir, layer, mom2, count, sample_size):
    name = f"text_model.encoder.layers.{layer}.mlp.fc2"
    path = os.path.join(stats_dir, "text_encoder", "ccs_filtered_stats", f"{name}_float32_mom2_t3072_{sample_size}.npz")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, **{"mom2.constructor": "util.runningstats.SecondMoment()", "mom2.count": count,
                      "mom2.mom2": mom2, "sample_size": sample_size})


def tiny_solve(out, edit_weight, lam, n_req=10, layers=(0, 1), sample_size=120):
    """Reference execute_emcid_text_encoder + apply_emcid_to_text_encoder on the tiny model."""
    model = rh.make_clip_text_model("tiny", seed=0)
    caps = rh.make_captions(150, model.config.vocab_size, seed=3)
    tok = rh.FakeTokenizer(model.config.vocab_size)
    reqs = rh.make_requests(n_req)
    tmp = tempfile.mkdtemp()
    try:
        stats_dir = os.path.join(tmp, "stats")
        moms = {}
        for l in layers:
            st = rh.run_reference_layer_stats(model, caps, l, stats_dir, sample_size)
            moms[l] = (st.mom2.mom2.numpy().copy(), st.mom2.count)
        cache = os.path.join(tmp, "vstar", "c_")
        zs = rh.write_vstar_cache(cache, reqs, model.config.hidden_size, seed=2)
        hp = rh.make_hparams(layers, sample_size, mom2_update_weight=lam, edit_weight=edit_weight)
        cap = []
        w0 = {l: model.text_model.encoder.layers[l].mlp.fc2.weight.detach().clone() for l in layers}
        deltas = rh.run_reference_execute(model, tok, reqs, hp, cache, stats_dir, capture=cap)
        for l in layers:  # invariant: weights restored
            assert torch.equal(w0[l], model.text_model.encoder.layers[l].mlp.fc2.weight)
        rh.run_reference_apply(model, tok, reqs, hp, cache, stats_dir)
        d = dict(edit_weight=edit_weight, lam=lam, n_req=n_req, layers=np.array(layers), sample_size=sample_size,
                 zs=zs.numpy())
        for i, l in enumerate(layers):
            name = f"text_model.encoder.layers.{l}.mlp.fc2.weight"
            d[f"adj_k.{l}"] = deltas[name][0].numpy()
            d[f"resid.{l}"] = deltas[name][1].numpy()
            d[f"solveM.{l}"] = cap[i][0].numpy()
            d[f"solveK.{l}"] = cap[i][1].numpy()
            d[f"mom2.{l}"] = moms[l][0]
            d[f"count.{l}"] = moms[l][1]
            d[f"w_before.{l}"] = w0[l].numpy()
            d[f"w_after.{l}"] = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().numpy()
        np.savez_compressed(os.path.join(GOLD, out), **d)
        print(out, "adj_k", d[f"adj_k.{layers[0]}"].shape, "cond(M)", np.linalg.cond(d[f"solveM.{layers[-1]}"]))
    finally:
        shutil.rmtree(tmp)


def tiny_sequential(out, n_edits=3, n_req=5, layers=(0, 1), sample_size=60, lam=4000, edit_weight=0.5):
    """BASELINE configs[4] pattern (experiments/sequential_editing.py:98-171): successive reference
    apply_emcid_to_text_encoder calls on the same model, each solving on the already-edited weights with the same
    cached statistics.  Keeps the fc2 weights after every edit."""
    model = rh.make_clip_text_model("tiny", seed=3)
    caps = rh.make_captions(80, model.config.vocab_size, seed=21)
    tok = rh.FakeTokenizer(model.config.vocab_size)
    tmp = tempfile.mkdtemp()
    try:
        stats_dir = os.path.join(tmp, "stats")
        d = dict(edit_weight=edit_weight, lam=lam, n_req=n_req, n_edits=n_edits, layers=np.array(layers),
                 sample_size=sample_size)
        for l in layers:
            st = rh.run_reference_layer_stats(model, caps, l, stats_dir, sample_size)
            d[f"mom2.{l}"] = st.mom2.mom2.numpy().copy()
            d[f"count.{l}"] = st.mom2.count
            d[f"w_before.{l}"] = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().numpy().copy()
        for e in range(n_edits):
            reqs = [dict(r, source=f"edit{e} {r['source']}") for r in rh.make_requests(n_req)]
            cache = os.path.join(tmp, f"v{e}", "c_")
            d[f"zs.{e}"] = rh.write_vstar_cache(cache, reqs, model.config.hidden_size, seed=10 + e).numpy()
            hp = rh.make_hparams(layers, sample_size, mom2_update_weight=lam, edit_weight=edit_weight)
            rh.run_reference_apply(model, tok, reqs, hp, cache, stats_dir)
            for l in layers:
                d[f"w_after.{e}.{l}"] = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().numpy().copy()
        np.savez_compressed(os.path.join(GOLD, out), **d)
        dw = d[f"w_after.{n_edits - 1}.{layers[-1]}"] - d[f"w_before.{layers[-1]}"]
        print(out, "edits", n_edits, "|dW|/|W| last layer", np.linalg.norm(dw) / np.linalg.norm(d[f"w_before.{layers[-1]}"]))
    finally:
        shutil.rmtree(tmp)


def clipl_stats_digest(out, n_caps=300, sample_size=256, layer=11, ncols=48):
    model = rh.make_clip_text_model("clip-l", seed=0)
    caps = rh.make_captions(n_caps, model.config.vocab_size, seed=5)
    tmp = tempfile.mkdtemp()
    try:
        stat = rh.run_reference_layer_stats(model, caps, layer, tmp, sample_size)
        m = stat.mom2.mom2.numpy()
        rng = np.random.RandomState(0)
        cols = np.sort(rng.choice(m.shape[0], ncols, replace=False))
        v = rng.randn(m.shape[0], 4)
        np.savez_compressed(
            os.path.join(GOLD, out), layer=layer, n_caps=n_caps, sample_size=sample_size, seed_caps=5,
            count=stat.mom2.count, cols=cols, mom2_cols=m[:, cols], diag=np.diag(m).copy(),
            fro=np.linalg.norm(m.astype(np.float64)), probe_v=v, probe_mv=m.astype(np.float64) @ v,
            weight_checksum=checksum(model))
        print(out, "count", stat.mom2.count, "fro", np.linalg.norm(m))
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    assert rh.reference_available(), "run in the build container (needs /root/reference)"
    torch.set_num_threads(os.cpu_count())
    tiny_stats("tiny", 1, 150, 120, 3072, 0, 3, "tiny_stats.npz")
    tiny_stats("tiny-gelu", 0, 90, None, 512, 1, 4, "tiny_gelu_stats.npz")
    tiny_solve("tiny_solve_ew05.npz", 0.5, 4000)
    tiny_solve("tiny_solve_ew06.npz", 0.6, 10000)
    tiny_sequential("tiny_sequential.npz")
    clipl_stats_digest("clipl_stats_digest.npz")
