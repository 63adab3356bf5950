"""GPU parity of the modules SURVEY.md §8 f4 names — UNet cross-attention K/V statistics and edit, the whole-CLIPModel
variant — against fixtures produced by the UNMODIFIED reference (oracle/gen_golden_f4.py), plus the robustness paths of
this round: resumable statistics (a worker killed mid-pass), deterministic accumulation, weight writes that bypass the
version counter, and the fp64 fallback of the solver on systems an fp32-class Cholesky cannot factor."""
import os
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from helpers import ROOT, fp64_gram_reference, padded_batch, rel_fro, rh
from oracle import emcid_oracle as orc

pytestmark = pytest.mark.gpu

MOM2_TOL = 1e-5
DW_TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    from emcid_b200 import _lib
    _lib.check(_lib.lib().emcid_device_check(0))
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _get(obj, name):
    for part in name.split("."):
        obj = getattr(obj, part)
    return obj


def _write_stats(path, mom2, count, sample_size):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, **{"mom2.constructor": "util.runningstats.SecondMoment()", "mom2.count": int(count), "mom2.mom2": mom2,
                      "sample_size": int(sample_size)})


# ------------------------------------------------------------------------------------------ cross-attention K/V
def test_cross_attn_stats_match_reference_fixture(dev, golden_dir, tmp_path):
    """layer_stats_cross_attn_kv (emcid/layer_stats.py:333-427): count bit exact, mom2 within tolerance of the
    reference's file; ONE pass writes the file of every K/V module (the reference runs one pass per module, :429-467)."""
    from emcid_b200 import layer_stats
    g = np.load(os.path.join(golden_dir, "tiny_cross_attn.npz"))
    names = [str(n) for n in g["names"]]
    pipe = rh.make_cross_attn_pipe(seed=0, device=dev)
    caps = rh.make_captions(int(g["n_caps"]), 1000, seed=int(g["seed_caps"]))
    ss = int(g["sample_size"])
    layer_stats.get_ccs_filtered_ds = lambda tokenizer: rh.SynthTokenDataset(caps)
    stat = layer_stats.layer_stats_cross_attn_kv(pipe, names[0], stats_dir=tmp_path, sample_size=ss, precision="float32",
                                                 progress=None, num_workers=0, captions_per_batch=16)
    assert layer_stats.LAST_PASS_INFO["native_forward"] and layer_stats.LAST_PASS_INFO["launches"] > 0
    assert stat.mom2.count == int(g[f"count.{names[0]}"])
    assert stat.mom2.mom2.device.type == "cpu" and stat.mom2.mom2.dtype == torch.float32
    assert rel_fro(stat.mom2.mom2.numpy(), g[f"mom2.{names[0]}"]) < MOM2_TOL
    assert rel_fro(stat.mom2.mom2.numpy(), g[f"mom2.{names[-1]}"]) < MOM2_TOL       # the same matrix for every module
    for n in names:
        f = layer_stats.stats_filename(tmp_path, "unet", "ccs_filtered", n, "float32", ["mom2"], 3072, ss)
        dat = np.load(f)
        assert int(dat["mom2.count"]) == stat.mom2.count and int(dat["sample_size"]) == ss
        assert np.array_equal(dat["mom2.mom2"], stat.mom2.mom2.numpy())
    assert [str(f) for f in g["stat_files"]][0] == os.path.relpath(
        layer_stats.stats_filename(tmp_path, "unet", "ccs_filtered", names[0], "float32", ["mom2"], 3072, ss), tmp_path)
    layer_stats.get_ccs_filtered_ds = lambda tokenizer: (_ for _ in ()).throw(AssertionError("cache miss"))
    again = layer_stats.layer_stats_cross_attn_kv(pipe, names[-1], stats_dir=tmp_path, sample_size=ss, precision="float32",
                                                  progress=None)
    assert again.mom2.count == stat.mom2.count and torch.equal(again.mom2.mom2, stat.mom2.mom2)
    with pytest.raises(LookupError):
        layer_stats.layer_stats_cross_attn_kv(pipe, "down_blocks.9.nope", stats_dir=tmp_path, sample_size=ss,
                                              precision="float32", progress=None)


def test_cross_attn_stats_at_clipl_width_match_fp64(dev, tmp_path):
    """The same pass at SD-v1.4 size (last_hidden_state of CLIP-L, d = 768, 12 layers + final layer norm) against the
    fp64 copy of the HF model."""
    import copy
    from emcid_b200 import layer_stats, synth
    model = rh.make_clip_text_model("clip-l", seed=0).to(dev)
    unet = rh.TinyUNet(768, seed=3).to(dev)
    pipe = SimpleNamespace(text_encoder=model, unet=unet, tokenizer=None, device=dev)
    caps = synth.make_caption_ids(1500, seed=29, full=False, min_len=4)
    layer_stats.get_ccs_filtered_ds = lambda tokenizer: synth.CaptionIdDataset(caps)
    name = layer_stats.get_all_cross_attn_kv_layer_names(pipe)[0]
    stat = layer_stats.layer_stats_cross_attn_kv(pipe, name, stats_dir=tmp_path, sample_size=len(caps), precision="float32",
                                                 progress=None, num_workers=0, keep_on_device=True)
    assert layer_stats.LAST_PASS_INFO["native_forward"]
    m64 = copy.deepcopy(model).double()
    ref = torch.zeros(768, 768, dtype=torch.float64, device=dev)
    total = 0
    with torch.no_grad():
        for c0 in range(0, len(caps), 500):
            batch = {k: v.to(dev) for k, v in padded_batch(caps[c0:c0 + 500]).items()}
            y = m64(**batch).last_hidden_state[batch["attention_mask"].bool()]
            ref += y.T @ y
            total += y.shape[0]
    assert stat.mom2.count == total == sum(len(c) for c in caps)
    assert float((stat.mom2.mom2.double() - ref).norm() / ref.norm()) < MOM2_TOL


def test_cross_attn_edit_matches_reference_fixture(dev, golden_dir, tmp_path):
    """execute_emcid_cross_attn / apply_emcid_to_cross_attn (emcid/emcid_main.py:314-547) on every K/V module of the
    miniature UNet: the reference's own statistics file, the same pickled v* files, deltas and applied weights compared."""
    from emcid_b200 import clip_forward, compute_ks, emcid_main, layer_stats
    g = np.load(os.path.join(golden_dir, "tiny_cross_attn.npz"))
    names = [str(n) for n in g["names"]]
    ss = int(g["sample_size"])
    pipe = rh.make_cross_attn_pipe(seed=0, device=dev)
    for n in names:
        assert np.array_equal(_get(pipe.unet, n).weight.cpu().numpy(), g[f"w_before.{n}"])
        _write_stats(str(layer_stats.stats_filename(tmp_path, "unet", "ccs_filtered", n, "float32", ["mom2"], 3072, ss)),
                     g[f"mom2.{names[0]}"], g[f"count.{names[0]}"], ss)
    reqs = rh.make_requests(int(g["n_req"]))
    cache = str(tmp_path / "v" / "c_")
    rh.write_cross_attn_vstar_cache(cache, reqs, pipe, names, seed=2)
    hp = rh.make_hparams([0], ss, mom2_update_weight=float(g["lam"]), edit_weight=float(g["edit_weight"]))
    emcid_main.COV_CACHE.clear()
    try:
        deltas = emcid_main.execute_emcid_cross_attn(pipe, reqs, hp, cache_name=cache, verbose=False, stat_dir=str(tmp_path))
        assert compute_ks.LAST_PATH["native"] is True
        assert list(deltas) == [f"{n}.weight" for n in names]
        for n in names:
            adj, resid = deltas[f"{n}.weight"]
            assert adj.dtype == torch.float64 and adj.device.type == "cpu" and tuple(adj.shape) == g[f"adj_k.{n}"].shape
            assert rel_fro(resid.numpy(), g[f"resid.{n}"]) < 1e-5
            assert rel_fro(adj.numpy(), g[f"adj_k.{n}"]) < DW_TOL
            assert rel_fro(resid.numpy() @ adj.numpy().T, g[f"resid.{n}"] @ g[f"adj_k.{n}"].T) < DW_TOL
            assert np.array_equal(_get(pipe.unet, n).weight.cpu().numpy(), g[f"w_before.{n}"])      # unchanged by execute
        emcid_main.apply_emcid_to_cross_attn(pipe, reqs, hp, device=dev, cache_name=cache, stat_dir=str(tmp_path))
        for n in names:
            w0 = g[f"w_before.{n}"].astype(np.float64)
            got = _get(pipe.unet, n).weight.detach().cpu().numpy().astype(np.float64) - w0
            assert rel_fro(got, g[f"w_after.{n}"].astype(np.float64) - w0) < DW_TOL, n
        with pytest.raises(NotImplementedError, match="v_star cache miss"):
            emcid_main.execute_emcid_cross_attn(pipe, [dict(reqs[0], source="nobody")], hp, cache_name=cache, verbose=False,
                                                stat_dir=str(tmp_path))
    finally:
        emcid_main.COV_CACHE.clear()
        clip_forward.release_key_encoders()


# ------------------------------------------------------------------------------------------ whole CLIPModel
def test_clip_model_variant_matches_reference_fixture(dev, golden_dir, tmp_path):
    """execute_emcid_clip / apply_emcid_to_clip (emcid/emcid_main.py:109-311): the stage-2 loop on a transformers.CLIPModel,
    keys from the library forward over its text tower; and the statistics pass itself on the CLIPModel (which the
    reference's model(**batch) cannot run without pixel values) against the reference's statistics of the tower."""
    from emcid_b200 import clip_forward, compute_ks, emcid_main, layer_stats
    g = np.load(os.path.join(golden_dir, "tiny_clip_model.npz"))
    layers = [int(l) for l in g["layers"]]
    ss = int(g["sample_size"])
    model = rh.make_clip_model(seed=5)
    for l in layers:
        assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), g[f"w_before.{l}"])
    model = model.to(dev)
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in layers]
    caps = rh.make_captions(int(g["n_caps"]), 1000, seed=int(g["seed_caps"]))
    layer_stats.get_ccs_filtered_ds = lambda tokenizer: rh.SynthTokenDataset(caps)
    stats = layer_stats.layer_stats_text_encoder_multi(model, None, names, stats_dir=tmp_path / "own", sample_size=ss,
                                                       precision="float32", progress=None, num_workers=0)
    assert layer_stats.LAST_PASS_INFO["native_forward"]
    for l, n in zip(layers, names):
        assert stats[n].mom2.count == int(g[f"count.{l}"])
        assert rel_fro(stats[n].mom2.mom2.numpy(), g[f"mom2.{l}"]) < MOM2_TOL
        _write_stats(str(layer_stats.stats_filename(tmp_path / "ref", "text_encoder", "ccs_filtered", n, "float32", ["mom2"],
                                                    3072, ss)), g[f"mom2.{l}"], g[f"count.{l}"], ss)
    tok = rh.FakeTokenizer(1000)
    processor = SimpleNamespace(tokenizer=tok)
    reqs = rh.make_requests(int(g["n_req"]))
    cache = str(tmp_path / "v" / "c_")
    zs = rh.write_vstar_cache(cache, reqs, 64, seed=2)
    assert np.array_equal(zs.numpy(), g["zs"])
    hp = rh.make_hparams(layers, ss, mom2_update_weight=float(g["lam"]), edit_weight=float(g["edit_weight"]))
    emcid_main.COV_CACHE.clear()
    try:
        deltas = emcid_main.execute_emcid_clip(model, processor, reqs, hp, cache_name=cache, verbose=False,
                                               stat_dir=tmp_path / "ref")
        assert compute_ks.LAST_PATH["native"] is True
        for l, n in zip(layers, names):
            adj, resid = deltas[n + ".weight"]
            assert rel_fro(resid.numpy() @ adj.numpy().T, g[f"resid.{l}"] @ g[f"adj_k.{l}"].T) < DW_TOL
            assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.cpu().numpy(), g[f"w_before.{l}"])
        out, orig = emcid_main.apply_emcid_to_clip(model, processor, reqs, hp, device=dev, cache_name=cache,
                                                   stat_dir=tmp_path / "ref", return_orig_text_model=True)
        assert out is model and orig is not None
        for l in layers:
            w0 = g[f"w_before.{l}"].astype(np.float64)
            got = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().cpu().numpy().astype(np.float64) - w0
            assert rel_fro(got, g[f"w_after.{l}"].astype(np.float64) - w0) < DW_TOL
            assert np.array_equal(orig.text_model.encoder.layers[l].mlp.fc2.weight.cpu().numpy(), g[f"w_before.{l}"])
    finally:
        emcid_main.COV_CACHE.clear()
        clip_forward.release_key_encoders()


# ------------------------------------------------------------------------------------------ robustness
_WORKER = r"""
import os, signal, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch
from helpers import rh
from emcid_b200 import layer_stats, synth
dev = torch.device("cuda:0")
model = rh.make_clip_text_model("tiny", seed=0).to(dev)
caps = synth.make_caption_ids(600, vocab=1000, seed=31, full=False, min_len=4)
layer_stats.get_ccs_filtered_ds = lambda tokenizer: synth.CaptionIdDataset(caps)
names = [f"text_model.encoder.layers.{{l}}.mlp.fc2" for l in (0, 1)]
die_after = int(sys.argv[2])
def progress(loader, total=None):
    for i, b in enumerate(loader):
        if die_after >= 0 and i == die_after:
            os.kill(os.getpid(), signal.SIGKILL)          # no cleanup, no atexit: a crashed worker
        yield b
layer_stats.layer_stats_text_encoder_multi(model, None, names, stats_dir=sys.argv[1], sample_size=500, precision="float32",
                                           progress=progress, num_workers=0, captions_per_batch=20, block_tokens=1024,
                                           checkpoint_every=3)
print("resumed_from", layer_stats.LAST_PASS_INFO.get("resumed_from_caption"))
"""


def test_killed_pass_resumes_to_an_identical_file(dev, tmp_path):
    """A worker process is SIGKILLed in the middle of a statistics pass; the next run of the same command continues from
    the last checkpoint and writes statistics files that are bit-identical to those of an uninterrupted pass
    (EMCID_DETERMINISTIC=1: every tile of the SYRK is accumulated by one CTA pair in a fixed order)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, EMCID_DETERMINISTIC="1")

    def run(out, die_after):
        return subprocess.run([sys.executable, str(script), str(out), str(die_after)], env=env, capture_output=True, text=True,
                              timeout=600)

    whole = run(tmp_path / "a", -1)
    assert whole.returncode == 0, whole.stderr[-2000:]
    killed = run(tmp_path / "b", 12)
    assert killed.returncode == -9
    assert list((tmp_path / "b").rglob(".resume_*.npz")) and not list((tmp_path / "b").rglob("text_model*.npz"))
    resumed = run(tmp_path / "b", -1)
    assert resumed.returncode == 0, resumed.stderr[-2000:]
    assert "resumed_from" in resumed.stdout and "resumed_from 0" not in resumed.stdout and "resumed_from None" not in resumed.stdout
    files = sorted(p.relative_to(tmp_path / "a") for p in (tmp_path / "a").rglob("text_model*.npz"))
    assert len(files) == 2 and not list((tmp_path / "b").rglob(".resume_*.npz"))
    for f in files:
        a, b = np.load(tmp_path / "a" / f), np.load(tmp_path / "b" / f)
        assert int(a["mom2.count"]) == int(b["mom2.count"]) > 0
        assert np.array_equal(a["mom2.mom2"], b["mom2.mom2"]), f                   # bit identical
    again = run(tmp_path / "c", -1)                                                 # and run to run
    for f in files:
        assert np.array_equal(np.load(tmp_path / "a" / f)["mom2.mom2"], np.load(tmp_path / "c" / f)["mom2.mom2"])


def test_deterministic_mode_matches_default_within_rounding(dev, tmp_path, monkeypatch):
    from emcid_b200 import layer_stats, synth
    model = rh.make_clip_text_model("clip-l", seed=0, num_hidden_layers=8).to(dev)
    caps = synth.make_caption_ids(700, seed=37, full=False, min_len=8)
    layer_stats.get_ccs_filtered_ds = lambda tokenizer: synth.CaptionIdDataset(caps)
    name = "text_model.encoder.layers.7.mlp.fc2"
    out = {}
    for mode in ("0", "1", "1"):
        monkeypatch.setenv("EMCID_DETERMINISTIC", mode)
        st = layer_stats.layer_stats_text_encoder(model, None, name, stats_dir=tmp_path, sample_size=len(caps),
                                                  precision="float32", progress=None, num_workers=0, force_recompute=True)
        out.setdefault(mode, []).append(st.mom2.mom2.clone())
    assert torch.equal(out["1"][0], out["1"][1])                                    # fixed accumulation order
    assert rel_fro(out["0"][0].numpy(), out["1"][0].numpy()) < 1e-6                 # same numbers up to fp32 reordering


def test_weight_writes_behind_the_version_counter_are_seen(dev):
    """`param.data.copy_()` does not bump Tensor._version; the key extraction compares content checksums once per edit
    (clip_forward.NativeClipTextEncoder.sync_weights(verify=True)) and must follow such a write."""
    from emcid_b200 import clip_forward, compute_ks
    model = rh.make_clip_text_model("tiny", seed=4).to(dev)
    tok = rh.FakeTokenizer(1000)
    reqs = rh.make_requests(6)
    name = "text_model.encoder.layers.1.mlp.fc2"
    try:
        k0, z0 = compute_ks.get_module_input_output_at_words(model, tok, reqs, name)
        assert compute_ks.LAST_PATH["native"]
        w = model.text_model.encoder.layers[0].mlp.fc1.weight
        v = w._version
        w.data.mul_(1.5)
        assert w._version == v                                                      # the write is invisible to the counter
        k1, z1 = compute_ks.get_module_input_output_at_words(model, tok, reqs, name)
        import copy
        os.environ["EMCID_NATIVE_KEYS"] = "0"
        k_hf, z_hf = compute_ks.get_module_input_output_at_words(copy.deepcopy(model), tok, reqs, name)
        os.environ.pop("EMCID_NATIVE_KEYS")
        assert rel_fro(k1.cpu().numpy(), k0.cpu().numpy()) > 1e-3
        assert rel_fro(k1.cpu().numpy(), k_hf.cpu().numpy()) < 5e-6 and rel_fro(z1.cpu().numpy(), z_hf.cpu().numpy()) < 5e-6
    finally:
        os.environ.pop("EMCID_NATIVE_KEYS", None)
        clip_forward.release_key_encoders()


def test_lookup_beyond_the_prompt_falls_back_to_the_hf_forward(dev):
    """Subject "" looks up the LAST COLUMN of the padded batch (causal_trace.py:1063-1064): for the shorter prompts that is
    a pad position, which the packed forward does not have — the traced HF forward serves the call."""
    from emcid_b200 import clip_forward, compute_ks
    model = rh.make_clip_text_model("tiny", seed=4).to(dev)
    tok = rh.FakeTokenizer(1000)
    reqs = [{"source": "", "dest": "art", "prompts": ["a long prompt about nothing", "short"]}]
    try:
        k, z = compute_ks.get_module_input_output_at_words(model, tok, reqs, "text_model.encoder.layers.1.mlp.fc2")
        assert compute_ks.LAST_PATH["native"] is False and tuple(k.shape) == (1, 256)
    finally:
        clip_forward.release_key_encoders()


@pytest.mark.parametrize("n", [40, 300])
def test_ill_conditioned_system_falls_back_to_fp64_lu(dev, n):
    """A covariance whose spectrum spans 1e12 under a random rotation (no diagonal scaling helps; as an fp32 matrix it is
    not even positive definite any more): the fp32-class Cholesky breaks down; the edit still returns the reference's
    answer, through the fp64 LU fallback, and says which path it took.  A benign system of the same size stays on the
    tensor-core path."""
    import warnings
    from emcid_b200 import emcid_main
    d, h = 768, 64
    g = torch.Generator(device=dev).manual_seed(5 + n)
    Q, _ = torch.linalg.qr(torch.randn(d, d, device=dev, dtype=torch.float64, generator=g))
    enc = SimpleNamespace(config=SimpleNamespace(_name_or_path="synthetic/illcond"))
    K = torch.randn(n, d, device=dev, generator=g) * 0.05
    S = torch.randn(n, h, device=dev, generator=g)
    for span, expect in ((1e12, "fp64_lu"), (1e3, None)):
        spec = torch.logspace(0, -np.log10(span), d, device=dev, dtype=torch.float64)
        cov = ((Q * spec) @ Q.T).float()
        cov = (cov + cov.T) / 2
        lam, ew = 1.0, 0.5
        Ks, Ss = K.double().T, S.double().T
        ref = (Ss / 2) @ torch.linalg.solve(lam * cov.double() + Ks @ Ks.T, Ks).T
        del emcid_main.LAST_SOLVE_PATHS[:]
        emcid_main.clear_factor_cache()
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            adj, resid, dW = emcid_main._solve_one_layer(enc, "layer", cov, K, S, lam, ew, 2, -1)
        err = float((dW.double() - ref).norm() / ref.norm())
        assert err < DW_TOL, (span, emcid_main.LAST_SOLVE_PATHS, err)
        if expect:
            assert emcid_main.LAST_SOLVE_PATHS[-1] == expect and any("fp64 LU" in str(w.message) for w in caught)
        else:
            assert emcid_main.LAST_SOLVE_PATHS[-1] in ("cached_factor", "direct")
    emcid_main.clear_factor_cache()
