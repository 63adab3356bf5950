"""Pins oracle/emcid_oracle.py (the CPU restatement) against the reference: committed fixtures made
by the unmodified reference (oracle/gen_golden.py), the reference's own known-answer test, and —
in the build container — the live reference."""
import os

import numpy as np
import pytest
import torch

from helpers import model_from_golden, rel_fro, rh, unpack_captions
from oracle import emcid_oracle as orc


@pytest.mark.parametrize("name", ["tiny_stats.npz", "tiny_gelu_stats.npz"])
def test_stats_oracle_matches_reference_fixture(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name), allow_pickle=False)
    model = model_from_golden(g)
    caps = unpack_captions(g)
    ss = None if int(g["sample_size"]) < 0 else int(g["sample_size"])
    stat = orc.layer_stats_oracle(model, [c.numpy() for c in caps], int(g["layer"]), ss, int(g["batch_tokens"]))
    assert stat.count == int(g["npz.mom2.count"])                      # bit exact
    assert rel_fro(stat.mom2, g["npz.mom2.mom2"]) < 1e-6               # same arithmetic, fp32 noise only
    assert stat.mom2.dtype == np.float32 and g["npz.mom2.mom2"].dtype == np.float32


def test_npz_layout_of_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_stats.npz"))
    assert str(g["npz.mom2.constructor"]) == "util.runningstats.SecondMoment()"
    assert g["npz.mom2.count"].dtype == np.int64 and g["npz.mom2.count"].shape == ()
    assert g["npz.sample_size"].dtype == np.int64
    assert str(g["rel_path"]) == orc.stats_filename("", "text_encoder", "ccs_filtered",
                                                    "text_model.encoder.layers.1.mlp.fc2", "float32", ["mom2"],
                                                    3072, 120).lstrip("/")
    st = orc.SecondMomentOracle()
    st.count, st.mom2 = 5, np.eye(3, dtype=np.float32)
    assert set(orc.combined_state(st, 7)) == {"mom2.constructor", "mom2.count", "mom2.mom2", "sample_size"}


def test_second_moment_known_answer():
    """The reference's own check (util/runningstats.py:1787-1789): moment() == X^T X / n."""
    rng = np.random.RandomState(0)
    data = rng.randn(5000, 24).astype(np.float32)
    st = orc.SecondMomentOracle()
    for i in range(0, 5000, 617):
        st.add(data[i: i + 617])
    st.add(data[:0])
    assert st.count == 5000
    np.testing.assert_allclose(st.moment(), data.T @ data / 5000, rtol=1e-2, atol=1e-4)
    assert rel_fro(st.moment(), data.astype(np.float64).T @ data.astype(np.float64) / 5000) < 1e-6


def test_length_collation_matches_survey_examples():
    # 100 full-length captions -> [(39,77),(39,77),(22,77)]
    subs = orc.length_collation([77] * 100, 3072)
    assert [len(s) for s in subs] == [39, 39, 22]
    lens = [77] * 39 + [49] * 61
    subs = orc.length_collation(lens, 3072)
    assert [(len(s), max(lens[i] for i in s)) for s in subs] == [(39, 77), (61, 49)]
    assert orc.length_collation([0, 0], 3072) == []
    assert orc.length_collation([5000], 3072) == [[0]]  # a single over-long sequence stays alone


@pytest.mark.parametrize("name", ["tiny_solve_ew05.npz", "tiny_solve_ew06.npz"])
def test_solve_block_matches_reference_fixture(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    lam, ew = float(g["lam"]), float(g["edit_weight"])
    layers = [int(l) for l in g["layers"]]
    for i, l in enumerate(layers):
        C = orc.cov_from_state(g[f"mom2.{l}"], int(g[f"count.{l}"]))
        Ks = g[f"solveK.{l}"]                       # already scaled by sqrt(ew/0.5), fp64
        K = (Ks / (ew / 0.5) ** 0.5).astype(np.float32)
        # M as the reference formed it
        C32 = (C * np.float32(1 - ew) / np.float32(0.5)).astype(np.float32)
        M = lam * C32.astype(np.float64) + Ks @ Ks.T
        assert rel_fro(M, g[f"solveM.{l}"]) < 1e-12
        resid = g[f"resid.{l}"]
        S = (resid * (len(layers) - i) / (ew / 0.5) ** 0.5).astype(np.float32)
        adj_k, resid_o, upd = orc.solve_layer(C, K, S, lam, ew, len(layers) - i)
        assert rel_fro(adj_k, g[f"adj_k.{l}"]) < 1e-7
        assert rel_fro(resid_o, resid) < 1e-6
        w_after = orc.apply_delta(g[f"w_before.{l}"], g[f"adj_k.{l}"], resid)
        assert rel_fro(w_after, g[f"w_after.{l}"]) < 1e-7


def test_execute_oracle_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiny_solve_ew05.npz"))
    model = rh.make_clip_text_model("tiny", seed=0)
    layers = [int(l) for l in g["layers"]]
    for l in layers:  # same weights as when the fixture was made
        assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), g[f"w_before.{l}"])
    tok = rh.FakeTokenizer(model.config.vocab_size)
    reqs = rh.make_requests(int(g["n_req"]))
    covs = {l: orc.cov_from_state(g[f"mom2.{l}"], int(g[f"count.{l}"])) for l in layers}
    deltas = orc.execute_oracle(model, tok, reqs, layers, g["zs"], covs, float(g["lam"]), float(g["edit_weight"]))
    for l in layers:
        assert rel_fro(deltas[l][0], g[f"adj_k.{l}"]) < 1e-5
        assert rel_fro(deltas[l][1], g[f"resid.{l}"]) < 1e-5
        upd = deltas[l][1] @ deltas[l][0].T
        assert rel_fro(upd, g[f"resid.{l}"] @ g[f"adj_k.{l}"].T) < 1e-5
        assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), g[f"w_before.{l}"])


def test_sequential_editing_oracle_matches_reference_fixture(golden_dir):
    """BASELINE configs[4]: three successive edits on the same model, generated by the unmodified reference
    (oracle/gen_golden.py::tiny_sequential); the oracle must reproduce the fc2 weights after every edit."""
    g = np.load(os.path.join(golden_dir, "tiny_sequential.npz"))
    layers = [int(l) for l in g["layers"]]
    model = rh.make_clip_text_model("tiny", seed=3)
    for l in layers:
        assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), g[f"w_before.{l}"])
    tok = rh.FakeTokenizer(model.config.vocab_size)
    covs = {l: orc.cov_from_state(g[f"mom2.{l}"], int(g[f"count.{l}"])) for l in layers}
    for e in range(int(g["n_edits"])):
        reqs = [dict(r, source=f"edit{e} {r['source']}") for r in rh.make_requests(int(g["n_req"]))]
        deltas = orc.execute_oracle(model, tok, reqs, layers, g[f"zs.{e}"], covs, float(g["lam"]), float(g["edit_weight"]))
        for l in layers:
            w = model.text_model.encoder.layers[l].mlp.fc2.weight
            with torch.no_grad():
                w[...] = torch.from_numpy(orc.apply_delta(w.numpy(), *deltas[l]))
            got = w.detach().numpy().astype(np.float64) - g[f"w_before.{l}"]
            want = g[f"w_after.{e}.{l}"].astype(np.float64) - g[f"w_before.{l}"]
            assert rel_fro(got, want) < 1e-8, (e, l)        # measured: bit-identical or <= 1.3e-10


def test_execute_oracle_matches_reference_digest_at_clipl_size(golden_dir):
    """The oracle at CLIP-L size against the unmodified reference (oracle/gen_golden_clipl.py): first sequential edit of
    tests/golden/clipl_sequential_digest.npz (100 requests, layers 7-11, template-correlated keys, cond ~ 3e6)."""
    from emcid_b200 import synth
    from helpers import weight_checksum
    g = np.load(os.path.join(golden_dir, "clipl_sequential_digest.npz"))
    layers = [int(x) for x in g["layers"]]
    model = rh.make_clip_text_model("clip-l", seed=0)
    if not np.allclose(weight_checksum(model), g["weight_checksum"], rtol=1e-12):
        pytest.skip("random-init CLIP-L differs from the fixture's (torch RNG stream changed)")
    tok = synth.WordHashTokenizer(49408)
    reqs = [dict(r, source=f"edit0 {r['source']}") for r in rh.make_requests(int(g["n_req"]))]
    g_v = torch.Generator().manual_seed(int(g["seed_vstar0"]))          # ref_harness.write_vstar_cache without the files
    zs = torch.stack([torch.randn(768, generator=g_v) for _ in reqs], dim=1).numpy()
    covs = {l: orc.exact_spd_matrix(3072, 3072 + 1024, seed=l) for l in layers}
    deltas = orc.execute_oracle(model, tok, reqs, layers, zs, covs, float(g["lam"]), float(g["edit_weight"]))
    rng = np.random.RandomState(0)
    P, Q = rng.randn(3072, int(g["n_probe"])), rng.randn(768, int(g["n_probe"]))
    for l in layers:
        w0 = model.text_model.encoder.layers[l].mlp.fc2.weight.numpy()
        dW = orc.apply_delta(w0, *deltas[l]).astype(np.float64) - w0
        assert rel_fro(dW @ P, g[f"cum.0.{l}.MP"]) < 1e-6, l
        assert rel_fro(Q.T @ dW, g[f"cum.0.{l}.QtM"]) < 1e-6, l
        assert abs(np.linalg.norm(dW) / float(g[f"cum.0.{l}.fro"]) - 1) < 1e-6, l


def _get(obj, name):
    for part in name.split("."):
        obj = getattr(obj, part)
    return obj


def test_cross_attn_oracle_matches_reference_fixture(golden_dir):
    """statistics of the cross-attention K/V input and the independent per-module edit, against the
    unmodified reference (oracle/gen_golden_f4.py::tiny_cross_attn)."""
    from emcid_b200 import layer_stats
    g = np.load(os.path.join(golden_dir, "tiny_cross_attn.npz"))
    names = [str(n) for n in g["names"]]
    pipe = rh.make_cross_attn_pipe(seed=0)
    assert float(pipe.text_encoder.text_model.encoder.layers[1].mlp.fc2.weight.double().abs().sum()) == float(g["text_fc2_checksum"])
    assert layer_stats.get_all_cross_attn_kv_layer_names(pipe) == names          # same modules, the reference's order
    for n in names:
        assert np.array_equal(_get(pipe.unet, n).weight.numpy(), g[f"w_before.{n}"])
    caps = [c.numpy() for c in rh.make_captions(int(g["n_caps"]), 1000, seed=int(g["seed_caps"]))]
    stat = orc.cross_attn_kv_stats_oracle(pipe.text_encoder, caps, int(g["sample_size"]))
    for n in (names[0], names[-1]):                                              # one matrix whatever the module
        assert stat.count == int(g[f"count.{n}"])
        assert rel_fro(stat.mom2, g[f"mom2.{n}"]) < 1e-6
    assert [str(f) for f in g["stat_files"]] == sorted(
        orc.stats_filename("", "unet", "ccs_filtered", n, "float32", ["mom2"], 3072, int(g["sample_size"])).lstrip("/")
        for n in (names[0], names[-1]))
    reqs = rh.make_requests(int(g["n_req"]))
    g_v = torch.Generator().manual_seed(2)                                       # ref_harness.write_cross_attn_vstar_cache
    zs = {n: [] for n in names}
    for _ in reqs:
        for n in names:
            zs[n].append(torch.randn(g[f"w_before.{n}"].shape[0], generator=g_v))
    zs = {n: torch.stack(v, dim=1).numpy() for n, v in zs.items()}
    cov = orc.cov_from_state(g[f"mom2.{names[0]}"], int(g[f"count.{names[0]}"]))
    deltas = orc.execute_cross_attn_oracle(pipe.text_encoder, pipe.tokenizer, reqs, {n: g[f"w_before.{n}"] for n in names},
                                           zs, cov, float(g["lam"]), float(g["edit_weight"]))
    for n in names:
        assert rel_fro(deltas[n][0], g[f"adj_k.{n}"]) < 1e-5
        assert rel_fro(deltas[n][1], g[f"resid.{n}"]) < 1e-5
        w_after = orc.apply_delta(g[f"w_before.{n}"], g[f"adj_k.{n}"], g[f"resid.{n}"])
        assert rel_fro(w_after, g[f"w_after.{n}"]) < 1e-7


def test_clip_model_variant_matches_reference_fixture(golden_dir):
    """apply_emcid_to_clip (emcid/emcid_main.py:109-311) runs the text-encoder loop on a whole CLIPModel: the oracle's
    stage-2 restatement on the model's text tower reproduces the reference's deltas and weights."""
    g = np.load(os.path.join(golden_dir, "tiny_clip_model.npz"))
    layers = [int(l) for l in g["layers"]]
    model = rh.make_clip_model(seed=5)
    for l in layers:
        assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), g[f"w_before.{l}"])
    tower = rh.text_tower_of(model)
    tok = rh.FakeTokenizer(1000)
    reqs = rh.make_requests(int(g["n_req"]))
    covs = {l: orc.cov_from_state(g[f"mom2.{l}"], int(g[f"count.{l}"])) for l in layers}
    deltas = orc.execute_oracle(tower, tok, reqs, layers, g["zs"], covs, float(g["lam"]), float(g["edit_weight"]))
    for l in layers:
        assert rel_fro(deltas[l][1] @ deltas[l][0].T, g[f"resid.{l}"] @ g[f"adj_k.{l}"].T) < 1e-5
        assert rel_fro(orc.apply_delta(g[f"w_before.{l}"], g[f"adj_k.{l}"], g[f"resid.{l}"]), g[f"w_after.{l}"]) < 1e-7


def test_exact_spd_matrix_is_reproducible():
    a = orc.exact_spd_matrix(64, 128, seed=3)
    b = orc.exact_spd_matrix(64, 128, seed=3)
    assert np.array_equal(a, b) and np.array_equal(a, a.T)
    assert np.linalg.eigvalsh(a.astype(np.float64)).min() > 0


@pytest.mark.reference
@pytest.mark.skipif(not rh.reference_available(), reason="needs /root/reference")
def test_find_token_range_matches_live_reference():
    import sys
    rh.import_reference()
    from experiments.causal_trace import find_token_range as ref_ftr
    tok = rh.FakeTokenizer(1000)
    for prompt, word in [("An image of artist3 name3", "artist3 name3"), ("artist7 name7", "artist7 name7"),
                         ("A photo of artist1 name1", "artist1 name1")]:
        ids = tok([prompt])["input_ids"][0]
        assert tuple(ref_ftr(tok, ids, word)) == tuple(orc.find_token_range(tok, ids, word))
