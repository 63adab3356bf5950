"""Parity of the CUDA path (through the C ABI) with the reference: committed fixtures produced by
the unmodified reference, the CPU oracle on seeded inputs, and size-independent properties at
CLIP-L / bigG sizes.  Tolerances are BASELINE.json's: token counts bit exact, mom2 <= 1e-5 and
dW <= 1e-4 relative Frobenius error."""
import os

import numpy as np
import pytest
import torch

from helpers import model_from_golden, rel_fro, rh, unpack_captions, weight_checksum
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


def _patch_ds(caps):
    from emcid_b200 import layer_stats
    layer_stats.get_ccs_filtered_ds = lambda tokenizer: rh.SynthTokenDataset(caps)
    return layer_stats


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K,kw", [
    (128, 256, 32, {}), (200, 260, 72, {}), (130, 132, 40, dict(n128=True)),
    (512, 768, 256, dict(chunk=1)), (1024, 1024, 2048, dict(lower=True, streamk=True, beta=1.0)),
])
def test_gemm3x_matches_fp64(dev, M, N, K, kw):
    from emcid_b200 import _lib
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    A = torch.randn(M, K, device=dev, generator=g)
    B = torch.randn(N, K, device=dev, generator=g)
    beta = kw.pop("beta", 0.0)
    C0 = torch.randn(M, N, device=dev, generator=g) if beta else None
    ref = A.double() @ B.double().T + (beta * C0.double() if beta else 0)
    C = _lib.gemm3x_nt(A, B, C0.clone() if beta else None, beta=beta, **kw)
    if kw.get("lower"):
        mask = torch.tril(torch.ones(M, N, device=dev))
        assert float(((C.double() - ref) * mask).norm() / (ref * mask).norm()) < 2e-6
    else:
        assert float((C.double() - ref).norm() / ref.norm()) < 2e-6


# ------------------------------------------------------------------------------------------ mom2 kernel vs oracle
@pytest.mark.parametrize("d,h,T,act,frac", [
    (256, 64, 300, "quick_gelu", 1.0), (256, 64, 1000, "gelu", 0.6), (200, 80, 513, "quick_gelu", 0.5),
    (3072, 768, 2000, "quick_gelu", 0.7), (5120, 1280, 1100, "gelu", 0.9),
])
def test_mom2_accumulator_matches_oracle(dev, d, h, T, act, frac):
    from emcid_b200.mom2 import Mom2Accumulator
    g = torch.Generator().manual_seed(d + T)
    W = torch.randn(d, h, generator=g) * (0.7 / h ** 0.5)
    b = torch.randn(d, generator=g) * 0.1
    X = torch.randn(T, h, generator=g)
    valid = (torch.rand(T, generator=g) < frac) if frac < 1 else None
    acc = Mom2Accumulator(dev, d, h, act)
    acc.set_weights(W.to(dev), b.to(dev))
    acc.add(X[: T // 3].to(dev), None if valid is None else valid[: T // 3].to(dev))
    acc.add(X[T // 3:].to(dev), None if valid is None else valid[T // 3:].to(dev))
    acc.add(X[:0].to(dev))                                              # empty batch: no-op
    mom2, count = acc.finalize()
    Xv = X.numpy() if valid is None else X.numpy()[valid.numpy()]
    A = orc.fc2_input(Xv, W.numpy(), b.numpy(), act, dtype=np.float64)
    assert int(count) == Xv.shape[0]
    m = mom2.cpu().numpy()
    assert np.array_equal(m, m.T)
    assert rel_fro(m, A.T @ A) < MOM2_TOL
    acc.reset()
    acc.add(X.to(dev), torch.zeros(T, dtype=torch.bool, device=dev))    # everything masked
    mom2, count = acc.finalize()
    assert int(count) == 0 and float(mom2.abs().max()) == 0.0
    acc.close()


# ------------------------------------------------------------------------------------------ layer_stats vs reference fixtures
@pytest.mark.parametrize("native", ["1", "0"])
@pytest.mark.parametrize("name", ["tiny_stats.npz", "tiny_gelu_stats.npz"])
def test_layer_stats_matches_reference_fixture(dev, golden_dir, tmp_path, monkeypatch, name, native):
    """native=1: the library's own forward (csrc/clip.cuh); native=0: HF forward with the fused kernels hooked in."""
    monkeypatch.setenv("EMCID_NATIVE_FORWARD", native)
    g = np.load(os.path.join(golden_dir, name))
    model = model_from_golden(g).to(dev)
    caps = unpack_captions(g)
    ls = _patch_ds(caps)
    ss = None if int(g["sample_size"]) < 0 else int(g["sample_size"])
    layer_name = f"text_model.encoder.layers.{int(g['layer'])}.mlp.fc2"
    stat = ls.layer_stats_text_encoder(model, None, layer_name, stats_dir=tmp_path, sample_size=ss,
                                       precision="float32", batch_tokens=int(g["batch_tokens"]), progress=None,
                                       captions_per_batch=32, num_workers=0)
    assert ls.LAST_PASS_INFO["native_forward"] == (native == "1") and ls.LAST_PASS_INFO["launches"] > 0
    assert stat.mom2.count == int(g["npz.mom2.count"])                       # bit exact
    assert stat.mom2.mom2.device.type == "cpu" and stat.mom2.mom2.dtype == torch.float32
    assert rel_fro(stat.mom2.mom2.numpy(), g["npz.mom2.mom2"]) < MOM2_TOL
    f = tmp_path / str(g["rel_path"])                                        # same file name as the reference
    dat = np.load(f)
    assert sorted(dat.files) == sorted(k[4:] for k in g.files if k.startswith("npz."))
    for k in dat.files:
        assert dat[k].dtype == g["npz." + k].dtype and dat[k].shape == g["npz." + k].shape, k
    # second call is a cache hit: no dataset access, identical numbers
    ls.get_ccs_filtered_ds = lambda tokenizer: (_ for _ in ()).throw(AssertionError("cache miss"))
    again = ls.layer_stats_text_encoder(model, None, layer_name, stats_dir=tmp_path, sample_size=ss,
                                        precision="float32", batch_tokens=int(g["batch_tokens"]), progress=None)
    assert again.mom2.count == stat.mom2.count and torch.equal(again.mom2.mom2, stat.mom2.mom2)


@pytest.mark.parametrize("native", ["1", "0"])
def test_clipl_layer_stats_matches_reference_digest(dev, golden_dir, tmp_path, monkeypatch, native):
    monkeypatch.setenv("EMCID_NATIVE_FORWARD", native)
    g = np.load(os.path.join(golden_dir, "clipl_stats_digest.npz"))
    model = rh.make_clip_text_model("clip-l", seed=0)
    if not np.allclose(weight_checksum(model), g["weight_checksum"], rtol=1e-12):
        pytest.skip("random-init CLIP-L differs from the fixture's (torch RNG stream changed)")
    model = model.to(dev)
    caps = rh.make_captions(int(g["n_caps"]), model.config.vocab_size, seed=int(g["seed_caps"]))
    ls = _patch_ds(caps)
    layer = int(g["layer"])
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (layer - 1, layer)]
    stats = ls.layer_stats_text_encoder_multi(model, None, names, stats_dir=tmp_path, sample_size=int(g["sample_size"]),
                                              precision="float32", progress=None, captions_per_batch=64, num_workers=0)
    m = stats[names[1]].mom2.mom2.numpy()
    assert stats[names[1]].mom2.count == int(g["count"]) == stats[names[0]].mom2.count
    assert rel_fro(m[:, g["cols"]], g["mom2_cols"]) < MOM2_TOL
    assert rel_fro(np.diag(m), g["diag"]) < MOM2_TOL
    assert rel_fro(m.astype(np.float64) @ g["probe_v"], g["probe_mv"]) < MOM2_TOL
    assert abs(np.linalg.norm(m.astype(np.float64)) / float(g["fro"]) - 1) < MOM2_TOL
    # the multi-layer pass equals a single-layer pass of the other layer
    single = ls.layer_stats_text_encoder(model, None, names[0], stats_dir=tmp_path / "s", sample_size=int(g["sample_size"]),
                                         precision="float32", progress=None, captions_per_batch=50, num_workers=0)
    assert single.mom2.count == stats[names[0]].mom2.count
    assert rel_fro(single.mom2.mom2.numpy(), stats[names[0]].mom2.mom2.numpy()) < 2e-6


@pytest.mark.parametrize("block_tokens", [96, 1000, 37888])
def test_token_budget_blocks_do_not_change_the_statistics(dev, golden_dir, tmp_path, block_tokens):
    """The native pass regroups loader batches into device blocks of at most `block_tokens` packed tokens
    (stat_dataset.PackedReblocker): whatever the block size — down to one or two captions per block — the count is
    bit exact and mom2 matches the reference fixture."""
    g = np.load(os.path.join(golden_dir, "tiny_stats.npz"))
    model = model_from_golden(g).to(dev)
    ls = _patch_ds(unpack_captions(g))
    ss = None if int(g["sample_size"]) < 0 else int(g["sample_size"])
    layer_name = f"text_model.encoder.layers.{int(g['layer'])}.mlp.fc2"
    stat = ls.layer_stats_text_encoder(model, None, layer_name, stats_dir=tmp_path, sample_size=ss, precision="float32",
                                       batch_tokens=int(g["batch_tokens"]), progress=None, captions_per_batch=7,
                                       num_workers=0, block_tokens=block_tokens, force_recompute=True)
    assert ls.LAST_PASS_INFO["native_forward"]
    assert stat.mom2.count == int(g["npz.mom2.count"])
    assert rel_fro(stat.mom2.mom2.numpy(), g["npz.mom2.mom2"]) < MOM2_TOL


def test_native_stats_at_bigg_width(dev, tmp_path):
    """OpenCLIP bigG layer shapes (h = 1280, d = 5120, 20 heads, erf-gelu; BASELINE configs[3]) through the native pass —
    210 pair tiles on 74 CTA pairs: two whole-tile rounds plus stream-K leftovers — against an fp64 copy of the model."""
    import copy
    model = rh.make_clip_text_model("bigg", seed=4, num_hidden_layers=3).to(dev)
    caps = rh.make_captions(400, model.config.vocab_size, seed=7, min_len=3)
    ls = _patch_ds(caps)
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (1, 2)]
    stats = ls.layer_stats_text_encoder_multi(model, None, names, stats_dir=tmp_path, sample_size=len(caps), precision="float32",
                                              progress=None, captions_per_batch=128, num_workers=0, force_recompute=True)
    assert ls.LAST_PASS_INFO["native_forward"]
    m64 = copy.deepcopy(model).double()
    batch = {k: v.to(dev) for k, v in _padded(caps).items()}
    keep = batch["attention_mask"].bool()
    feats = {}
    hooks = [m64.text_model.encoder.layers[l].mlp.fc2.register_forward_pre_hook(lambda m, a, l=l: feats.__setitem__(l, a[0]))
             for l in (1, 2)]
    with torch.no_grad():
        m64(**batch)
    for h in hooks:
        h.remove()
    for l, n in zip((1, 2), names):
        a = feats[l][keep]
        ref = (a.T @ a).cpu().numpy()
        assert stats[n].mom2.count == int(keep.sum())
        assert rel_fro(stats[n].mom2.mom2.numpy(), ref) < MOM2_TOL


def test_stats_properties_at_clipl_size(dev, tmp_path):
    """Linearity over caption sets, batching invariance, count = sum of lengths (CLIP-L, layers 7-11)."""
    model = rh.make_clip_text_model("clip-l", seed=0).to(dev)
    caps = rh.make_captions(768, model.config.vocab_size, seed=11)
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in range(7, 12)]

    def run(subset, per_batch, tag):
        ls = _patch_ds(subset)
        st = ls.layer_stats_text_encoder_multi(model, None, names, stats_dir=tmp_path / tag, sample_size=None,
                                               precision="float32", progress=None, captions_per_batch=per_batch,
                                               num_workers=0, keep_on_device=True)
        return {n: (st[n].mom2.count, st[n].mom2.mom2.double()) for n in names}

    full = run(caps, 128, "full")
    a, b = run(caps[:300], 77, "a"), run(caps[300:], 201, "b")
    total = sum(len(c) for c in caps)
    for n in names:
        assert full[n][0] == total == a[n][0] + b[n][0]
        s = a[n][1] + b[n][1]
        assert float((full[n][1] - s).norm() / s.norm()) < 3e-6
        assert float(torch.diagonal(full[n][1]).min()) > 0
        assert torch.equal(full[n][1], full[n][1].T)


# ------------------------------------------------------------------------------------------ native forward
def _padded(caps):
    B, L = len(caps), max(len(c) for c in caps)
    ids = torch.zeros(B, L, dtype=torch.long)
    pos = torch.zeros(B, L, dtype=torch.long)
    mask = torch.zeros(B, L, dtype=torch.long)
    for i, c in enumerate(caps):
        ids[i, :len(c)] = c
        pos[i, :len(c)] = torch.arange(len(c))
        mask[i, :len(c)] = 1
    return {"input_ids": ids, "position_ids": pos, "attention_mask": mask}


@pytest.mark.parametrize("kind,n_caps", [("tiny", 90), ("tiny-gelu", 41), ("clip-l", 40), ("bigg", 6)])
def test_native_forward_hidden_states_match_hf(dev, kind, n_caps):
    """Residual stream of the library's forward vs the HF forward the reference runs (fp64 copy of the same
    model as ground truth); the native fp32-class result must be as close to it as HF's own fp32 run."""
    import copy
    from emcid_b200 import clip_forward
    model = rh.make_clip_text_model(kind, seed=3).to(dev)
    caps = rh.make_captions(n_caps, model.config.vocab_size, seed=4, min_len=1)
    batch = {k: v.to(dev) for k, v in _padded(caps).items()}
    keep = batch["attention_mask"].bool()
    with torch.no_grad():
        hs64 = copy.deepcopy(model).double()(**batch, output_hidden_states=True).hidden_states
        hs32 = model(**batch, output_hidden_states=True).hidden_states
    ids, pos, cu, S, T = clip_forward.pack_batch(batch, model.config.max_position_embeddings)
    assert T == sum(len(c) for c in caps) and S == n_caps
    nat = clip_forward.NativeClipTextEncoder(model, T, S)
    try:
        for n in (0, 1, len(hs64) - 1):
            h = nat.forward_hidden(ids, pos, cu, S, T, n)
            ref = hs64[n][keep]
            err = float((h.double() - ref).norm() / ref.norm())
            err_hf = float((hs32[n][keep].double() - ref).norm() / ref.norm())
            assert err < max(3e-6, 3 * err_hf), (n, err, err_hf)
    finally:
        nat.close()


def test_short_captions_share_attention_tiles(dev):
    """Short captions (real caption data, the prompts of an edit) are grouped several to an attention tile with a block
    diagonal causal mask (csrc/attn.cuh, clip_group_captions_kernel): 300 captions of 2..12 tokens with a few full-length and
    one-token ones in between, two layers deep, against an fp64 copy of the HF model; no token may see a neighbouring
    caption (a leak would show as an error of order 1)."""
    import copy
    from emcid_b200 import clip_forward
    model = rh.make_clip_text_model("clip-l", seed=5, num_hidden_layers=2).to(dev)
    vocab = model.config.vocab_size
    g = torch.Generator().manual_seed(12)
    lens = torch.randint(2, 13, (300,), generator=g).tolist()
    for at, L in ((0, 77), (17, 1), (18, 77), (19, 76), (150, 1), (151, 1), (299, 77)):
        lens[at] = L
    caps = [torch.cat([torch.tensor([vocab - 2]), torch.randint(0, vocab - 2, (max(L - 2, 0),), generator=g),
                       torch.tensor([vocab - 1])])[:L] for L in lens]
    batch = {k: v.to(dev) for k, v in _padded(caps).items()}
    keep = batch["attention_mask"].bool()
    with torch.no_grad():
        hs64 = copy.deepcopy(model).double()(**batch, output_hidden_states=True).hidden_states
        hs32 = model(**batch, output_hidden_states=True).hidden_states
    ids, pos, cu, S, T = clip_forward.pack_batch(batch, model.config.max_position_embeddings)
    assert T == sum(lens) and S == len(lens)
    nat = clip_forward.NativeClipTextEncoder(model, T, S)
    try:
        for n in (1, 2):
            h = nat.forward_hidden(ids, pos, cu, S, T, n)
            ref = hs64[n][keep]
            err = float((h.double() - ref).norm() / ref.norm())
            err_hf = float((hs32[n][keep].double() - ref).norm() / ref.norm())
            assert err < max(3e-6, 3 * err_hf), (n, err, err_hf)
            worst = float(((h.double() - ref).norm(dim=1) / ref.norm(dim=1)).max())     # per token: no outlier rows
            assert worst < 1e-4, worst
    finally:
        nat.close()


@pytest.mark.parametrize("kind,layer,n_req", [("tiny", 1, 7), ("tiny-gelu", 0, 5), ("clip-l", 9, 40)])
def test_native_key_extraction_matches_hf(dev, monkeypatch, kind, layer, n_req):
    """Keys (fc2 input) and current outputs (fc2 output) at the last subject token: the library's forward
    (emcid_clip_forward_keys) vs the traced HF forward of the reference (compute_z.py:2252-2327), before and
    after an in-place weight edit (the insert loop writes W in place between layers, emcid_main.py:1061)."""
    from emcid_b200 import clip_forward, compute_ks
    model = rh.make_clip_text_model(kind, seed=8).to(dev)
    tok = rh.FakeTokenizer(model.config.vocab_size)
    reqs = rh.make_requests(n_req)
    name = f"text_model.encoder.layers.{layer}.mlp.fc2"

    def both():
        monkeypatch.setenv("EMCID_NATIVE_KEYS", "0")
        k_hf, z_hf = compute_ks.get_module_input_output_at_words(model, tok, reqs, name)
        assert compute_ks.LAST_PATH["native"] is False
        monkeypatch.setenv("EMCID_NATIVE_KEYS", "1")
        k, z = compute_ks.get_module_input_output_at_words(model, tok, reqs, name)
        assert compute_ks.LAST_PATH["native"] is True
        return k, z, k_hf, z_hf

    try:
        k, z, k_hf, z_hf = both()
        assert k.shape == k_hf.shape == (n_req, model.config.intermediate_size)
        assert z.shape == z_hf.shape == (n_req, model.config.hidden_size)
        assert rel_fro(k.cpu().numpy(), k_hf.cpu().numpy()) < 5e-6
        assert rel_fro(z.cpu().numpy(), z_hf.cpu().numpy()) < 5e-6
        with torch.no_grad():   # edit an earlier layer and the traced one in place
            model.text_model.encoder.layers[0].mlp.fc2.weight[...] *= 1.25
            model.text_model.encoder.layers[layer].mlp.fc2.weight[...] += 0.01
        k2, z2, k2_hf, z2_hf = both()
        assert rel_fro(z2_hf.cpu().numpy(), z_hf.cpu().numpy()) > 1e-3          # the edit is visible at all
        assert rel_fro(k2.cpu().numpy(), k2_hf.cpu().numpy()) < 5e-6
        assert rel_fro(z2.cpu().numpy(), z2_hf.cpu().numpy()) < 5e-6
    finally:
        clip_forward.release_key_encoders()


@pytest.mark.parametrize("kind,la,lb", [("tiny", 0, 1), ("clip-l", 7, 9)])
def test_native_key_extraction_resumes_after_fc2_edit(dev, monkeypatch, kind, la, lb):
    """The edit loop asks for keys at increasing layers over the same prompts and writes only fc2 of the layer it just
    solved in between (emcid_main.py:1061): the library continues from the previous call's state (finish layer la with
    the new fc2, run (la, lb]) and must agree with a from-scratch traced HF forward; any other weight change in between
    forces a fresh start."""
    from emcid_b200 import clip_forward, compute_ks
    model = rh.make_clip_text_model(kind, seed=9).to(dev)
    tok = rh.FakeTokenizer(model.config.vocab_size)
    reqs = rh.make_requests(12)
    name = "text_model.encoder.layers.{}.mlp.fc2"

    def check(k, z, layer):
        """against an fp64 copy of the model; the library must be as close to it as HF's own fp32 run"""
        import copy
        monkeypatch.setenv("EMCID_NATIVE_KEYS", "0")
        k64, z64 = compute_ks.get_module_input_output_at_words(copy.deepcopy(model).double(), tok, reqs, name.format(layer))
        k32, z32 = compute_ks.get_module_input_output_at_words(model, tok, reqs, name.format(layer))
        monkeypatch.setenv("EMCID_NATIVE_KEYS", "1")
        for got, hf32, ref in ((k, k32, k64), (z, z32, z64)):
            ref = ref.cpu().numpy()
            assert rel_fro(got.cpu().numpy(), ref) < max(3e-6, 3 * rel_fro(hf32.cpu().numpy(), ref))

    try:
        prepared = compute_ks.prepare_lookup(tok, reqs, 1, dev)
        compute_ks.get_module_input_output_at_words(model, tok, reqs, name.format(la), prepared=prepared)
        assert compute_ks.LAST_PATH["native"] and compute_ks.LAST_PATH["resumed_from"] == -1
        with torch.no_grad():
            model.text_model.encoder.layers[la].mlp.fc2.weight[...] += 0.02
        k, z = compute_ks.get_module_input_output_at_words(model, tok, reqs, name.format(lb), prepared=prepared)
        assert compute_ks.LAST_PATH["resumed_from"] == la
        check(k, z, lb)
        # a second prompt set, and a change outside fc2 of the previous layer: no continuation
        prepared2 = compute_ks.prepare_lookup(tok, reqs, 1, dev)
        compute_ks.get_module_input_output_at_words(model, tok, reqs, name.format(la), prepared=prepared2)
        assert compute_ks.LAST_PATH["resumed_from"] == -1
        with torch.no_grad():
            model.text_model.encoder.layers[la].mlp.fc2.weight[...] -= 0.01
            model.text_model.encoder.layers[0].mlp.fc1.bias[...] += 0.05
        k, z = compute_ks.get_module_input_output_at_words(model, tok, reqs, name.format(lb), prepared=prepared2)
        assert compute_ks.LAST_PATH["resumed_from"] == -1
        check(k, z, lb)
    finally:
        clip_forward.release_key_encoders()


@pytest.mark.parametrize("kind,layer", [("tiny", 1), ("clip-l", 7)])
def test_prefetched_forward_gathers_the_same_keys(dev, kind, layer):
    """The edit launches the forward of the prompts before the subject tokens are looked up (compute_ks.prefetch_keys:
    emcid_clip_forward_keys with a placeholder row) and the keys call that follows only gathers its rows from that state
    (resume_layer == layer): bit-identical to a keys call that runs the whole forward itself; a weight written between
    the two calls (anything but fc2 of that layer) voids the state."""
    from emcid_b200 import clip_forward, compute_ks
    model = rh.make_clip_text_model(kind, seed=4).to(dev)
    tok = rh.FakeTokenizer(model.config.vocab_size)
    reqs = rh.make_requests(9)
    name = f"text_model.encoder.layers.{layer}.mlp.fc2"
    try:
        plain = compute_ks.prepare_lookup(tok, reqs, 1, dev)
        k0, z0 = compute_ks.get_module_input_output_at_words(model, tok, reqs, name, prepared=plain)
        assert compute_ks.LAST_PATH["native"] and compute_ks.LAST_PATH["resumed_from"] == -1
        launched = []
        pre = compute_ks.prepare_lookup(tok, reqs, 1, dev, after_tokenise=lambda enc, serial: launched.append(
            compute_ks.prefetch_keys(model, enc, name, serial)))
        assert launched == [True]
        k1, z1 = compute_ks.get_module_input_output_at_words(model, tok, reqs, name, prepared=pre)
        assert compute_ks.LAST_PATH["resumed_from"] == layer
        assert torch.equal(k0, k1) and torch.equal(z0, z1)
        pre2 = compute_ks.prepare_lookup(tok, reqs, 1, dev, after_tokenise=lambda enc, serial: compute_ks.prefetch_keys(
            model, enc, name, serial))
        with torch.no_grad():
            model.text_model.encoder.layers[layer].mlp.fc1.bias[...] += 0.03
        k2, _ = compute_ks.get_module_input_output_at_words(model, tok, reqs, name, prepared=pre2)
        assert compute_ks.LAST_PATH["resumed_from"] == -1 and not torch.equal(k0, k2)
    finally:
        clip_forward.release_key_encoders()


@pytest.mark.parametrize("max_pos,heads", [(128, 2), (48, 2), (77, 1)])
def test_native_forward_other_caption_lengths(dev, max_pos, heads):
    """Head dim 64 with other position limits than CLIP's 77: up to 128 tokens (the 8-chunk attention instantiation, 16 KB
    tiles, two CTAs per SM), 48 tokens (3 S chunks) and a single 77-token head; hidden states against an fp64 copy."""
    import copy
    from emcid_b200 import clip_forward
    model = rh.make_clip_text_model("tiny", seed=6, hidden_size=64 * heads, num_attention_heads=heads,
                                    max_position_embeddings=max_pos, num_hidden_layers=2).to(dev)
    caps = rh.make_captions(70, model.config.vocab_size, seed=8, min_len=1, max_len=max_pos)
    caps[0] = rh.make_captions(1, model.config.vocab_size, seed=9, max_len=max_pos, full=True)[0]   # one full-length caption
    batch = {k: v.to(dev) for k, v in _padded(caps).items()}
    keep = batch["attention_mask"].bool()
    with torch.no_grad():
        hs64 = copy.deepcopy(model).double()(**batch, output_hidden_states=True).hidden_states
        hs32 = model(**batch, output_hidden_states=True).hidden_states
    ids, pos, cu, S, T = clip_forward.pack_batch(batch, max_pos)
    nat = clip_forward.NativeClipTextEncoder(model, T, S)
    try:
        for n in (1, 2):
            h = nat.forward_hidden(ids, pos, cu, S, T, n)
            ref = hs64[n][keep]
            err = float((h.double() - ref).norm() / ref.norm())
            err_hf = float((hs32[n][keep].double() - ref).norm() / ref.norm())
            assert err < max(3e-6, 3 * err_hf), (n, err, err_hf)
    finally:
        nat.close()


def test_native_forward_falls_back_on_non_right_padding(dev, tmp_path):
    """A mask that is not a right-padding mask cannot be packed: the block goes through the HF forward with the
    fused kernels hooked in, and the statistics still match a direct masked Gram."""
    from emcid_b200 import clip_forward, layer_stats
    model = rh.make_clip_text_model("tiny", seed=5).to(dev)
    caps = rh.make_captions(24, model.config.vocab_size, seed=6, min_len=6)
    batch = _padded(caps)
    batch["attention_mask"][:, 2] = 0                          # a hole in the middle of every caption
    assert clip_forward.pack_batch(batch, 77) is None
    name = "text_model.encoder.layers.1.mlp.fc2"
    runner = layer_stats.TextEncoderMom2Pass(model, [name])
    runner.run_batch(batch)
    mom2, count = runner.finalize()[name]
    assert runner._native is None
    runner.close()
    feats = {}
    hook = model.text_model.encoder.layers[1].mlp.fc2.register_forward_pre_hook(lambda m, a: feats.__setitem__("a", a[0]))
    with torch.no_grad():
        model(**{k: v.to(dev) for k, v in batch.items()})
    hook.remove()
    a = feats["a"][batch["attention_mask"].bool().to(dev)].double()
    assert int(count) == a.shape[0]
    assert float((mom2.double() - a.T @ a).norm() / (a.T @ a).norm()) < MOM2_TOL


# ------------------------------------------------------------------------------------------ closed-form update
@pytest.mark.parametrize("name", ["tiny_solve_ew05.npz", "tiny_solve_ew06.npz"])
def test_solve_matches_reference_fixture(dev, golden_dir, name):
    from emcid_b200.solve import solve_layers
    g = np.load(os.path.join(golden_dir, name))
    lam, ew = float(g["lam"]), float(g["edit_weight"])
    layers = [int(l) for l in g["layers"]]
    s = (ew / 0.5) ** 0.5
    for i, l in enumerate(layers):
        C = torch.from_numpy(orc.cov_from_state(g[f"mom2.{l}"], int(g[f"count.{l}"]))).to(dev)
        C32 = C * (1 - ew) / 0.5
        Kt = torch.from_numpy((g[f"solveK.{l}"] / s).astype(np.float32).T.copy()).to(dev)
        St = torch.from_numpy((g[f"resid.{l}"] * (len(layers) - i) / s).astype(np.float32).T.copy()).to(dev)
        adj, resid, dW = solve_layers(C32, Kt, St, lam, s, [len(layers) - i])
        ref_upd = g[f"resid.{l}"] @ g[f"adj_k.{l}"].T
        assert rel_fro(dW[0].cpu().numpy(), ref_upd) < DW_TOL
        assert rel_fro(adj[0].cpu().numpy(), g[f"adj_k.{l}"]) < DW_TOL
        assert rel_fro(resid[0].cpu().numpy(), g[f"resid.{l}"]) < 1e-6


@pytest.mark.parametrize("name", ["tiny_solve_ew05.npz", "tiny_solve_ew06.npz"])
def test_execute_and_apply_match_reference_fixture(dev, golden_dir, tmp_path, name):
    """Whole edit through the reference-shaped API on the tiny model: deltas, restored weights, applied weights."""
    from emcid_b200 import emcid_main
    from types import SimpleNamespace
    g = np.load(os.path.join(golden_dir, name))
    layers = [int(l) for l in g["layers"]]
    ss = int(g["sample_size"])
    model = rh.make_clip_text_model("tiny", seed=0)
    for l in layers:
        assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), g[f"w_before.{l}"])
    model = model.to(dev)
    tok = rh.FakeTokenizer(model.config.vocab_size)
    reqs = rh.make_requests(int(g["n_req"]))
    stats_dir = tmp_path / "stats"
    for l in layers:  # the reference's own statistics files -> identical C on both sides
        f = orc.stats_filename(str(stats_dir), "text_encoder", "ccs_filtered", f"text_model.encoder.layers.{l}.mlp.fc2",
                               "float32", ["mom2"], 3072, ss)
        os.makedirs(os.path.dirname(f), exist_ok=True)
        np.savez(f, **{"mom2.constructor": "util.runningstats.SecondMoment()", "mom2.count": int(g[f"count.{l}"]),
                       "mom2.mom2": g[f"mom2.{l}"], "sample_size": ss})
    cache = str(tmp_path / "v" / "c_")
    rh.write_vstar_cache(cache, reqs, model.config.hidden_size, seed=2)
    hp = rh.make_hparams(layers, ss, mom2_update_weight=float(g["lam"]), edit_weight=float(g["edit_weight"]))
    pipe = SimpleNamespace(text_encoder=model, tokenizer=tok, device=dev)
    emcid_main.COV_CACHE.clear()
    deltas = emcid_main.execute_emcid_text_encoder(pipe, reqs, hp, cache_name=cache, verbose=False, stat_dir=stats_dir)
    assert list(deltas) == [f"text_model.encoder.layers.{l}.mlp.fc2.weight" for l in layers]
    for l in layers:
        adj, resid = deltas[f"text_model.encoder.layers.{l}.mlp.fc2.weight"]
        assert adj.dtype == torch.float64 and adj.device.type == "cpu" and tuple(adj.shape) == g[f"adj_k.{l}"].shape
        assert rel_fro(resid.numpy(), g[f"resid.{l}"]) < 1e-5
        assert rel_fro(resid.numpy() @ adj.numpy().T, g[f"resid.{l}"] @ g[f"adj_k.{l}"].T) < DW_TOL
        w = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().cpu().numpy()
        assert np.array_equal(w, g[f"w_before.{l}"])                       # invariant: restored
    emcid_main.apply_emcid_to_text_encoder(pipe, reqs, hp, device=dev, cache_name=cache, stats_dir=stats_dir, verbose=False)
    for l in layers:
        w = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().cpu().numpy()
        dw_ref = g[f"w_after.{l}"].astype(np.float64) - g[f"w_before.{l}"]
        assert rel_fro(w.astype(np.float64) - g[f"w_before.{l}"], dw_ref) < DW_TOL


@pytest.mark.parametrize("d,h,n,B", [(3072, 768, 96, 1), (3072, 768, 1000, 2), (5120, 1280, 64, 1)])
def test_solve_full_size_matches_oracle(dev, d, h, n, B):
    """CLIP-L / bigG sized systems with cond ~1e6, bit-reproducible inputs, oracle = fp64 LU on CPU."""
    from emcid_b200.solve import solve_layers
    Cs, Ks, Ss = [], [], []
    for b in range(B):
        Cs.append(orc.exact_spd_matrix(d, d + 1024, seed=b))
        g = torch.Generator().manual_seed(100 + b)
        Ks.append((torch.randn(n, d, generator=g) * 0.3 + 0.1).numpy())
        Ss.append(torch.randn(n, h, generator=g).numpy())
    lam, ew = 4000.0, 0.5
    left = list(range(B, 0, -1))
    adj, resid, dW = solve_layers(torch.from_numpy(np.stack(Cs)).to(dev), torch.from_numpy(np.stack(Ks)).to(dev),
                                  torch.from_numpy(np.stack(Ss)).to(dev), lam, 1.0, left)
    for b in range(B):
        a_ref, r_ref, u_ref = orc.solve_layer(Cs[b], Ks[b].T, Ss[b].T, lam, ew, left[b])
        assert rel_fro(dW[b].cpu().numpy(), u_ref) < DW_TOL
        assert rel_fro(adj[b].cpu().numpy(), a_ref) < DW_TOL
        assert rel_fro(resid[b].cpu().numpy(), r_ref) < 1e-6


def test_solve_reports_breakdown(dev):
    from emcid_b200 import _lib
    from emcid_b200.solve import CachedFactor, solve_layers
    C = -torch.eye(256, device=dev)
    with pytest.raises(_lib.EmcidError):
        solve_layers(C, torch.zeros(4, 256, device=dev), torch.zeros(4, 64, device=dev), 4000.0, 1.0, [1])
    with pytest.raises(_lib.EmcidError):
        CachedFactor(C, 4000.0)


@pytest.mark.parametrize("d,h,ns", [(3072, 768, (96, 700, 1)), (5120, 1280, (64,)), (256, 64, (12, 130))])
def test_cached_factor_matches_oracle(dev, d, h, ns):
    """One factorisation of lambda * C, several edits of different widths through the push-through identity:
    same (adj_k, resid, dW) as the reference's fp64 LU of lambda * C + Ks Ks^T, and as the direct solver."""
    from emcid_b200.solve import CachedFactor, solve_layers
    C = orc.exact_spd_matrix(d, d + 1024, seed=5)
    lam, ew = 4000.0, 0.6
    s = (ew / 0.5) ** 0.5
    C32 = torch.from_numpy(C).to(dev) * (1 - ew) / 0.5
    fac = CachedFactor(C32, lam)
    for j, n in enumerate(ns):
        g = torch.Generator().manual_seed(200 + j)
        K = (torch.randn(n, d, generator=g) * 0.3 + 0.1).numpy()
        S = torch.randn(n, h, generator=g).numpy()
        adj, resid, dW = fac.solve(torch.from_numpy(K).to(dev), torch.from_numpy(S).to(dev), s, 3)
        assert tuple(adj.shape) == (d, n) and tuple(resid.shape) == (h, n) and tuple(dW.shape) == (h, d)
        a_ref, r_ref, u_ref = orc.solve_layer(C, K.T, S.T, lam, ew, 3)
        assert rel_fro(dW.cpu().numpy(), u_ref) < DW_TOL
        assert rel_fro(adj.cpu().numpy(), a_ref) < DW_TOL
        assert rel_fro(resid.cpu().numpy(), r_ref) < 1e-6
        adj_d, _, dW_d = solve_layers(C32, torch.from_numpy(K).to(dev), torch.from_numpy(S).to(dev), lam, s, [3])
        # consistency of the two paths: each refines to a predicted error of <= 2e-5 (SOLVE_ADAPT_TOL)
        assert rel_fro(dW.cpu().numpy(), dW_d[0].cpu().numpy()) < 5e-5
        assert rel_fro(adj.cpu().numpy(), adj_d[0].cpu().numpy()) < 5e-5
    fac.close()


@pytest.mark.parametrize("cached", [False, True])
def test_sequential_editing_matches_oracle(dev, tmp_path, monkeypatch, cached):
    """Config-5 pattern (experiments/sequential_editing.py): successive edits on the same model reuse the
    cached C and re-solve on the already-edited weights — with and without the cached factorisation of lambda * C."""
    from emcid_b200 import emcid_main
    from types import SimpleNamespace
    emcid_main.clear_factor_cache()
    if cached:
        monkeypatch.setattr(emcid_main, "FACTOR_CACHE_MAX_FRACTION", 1)     # tiny model: d = 256, n_pad = 128
    else:
        monkeypatch.setenv("EMCID_FACTOR_CACHE", "0")
    layers, ss = [0, 1], 60
    caps = rh.make_captions(80, 1000, seed=21)
    m_gpu = rh.make_clip_text_model("tiny", seed=3).to(dev)
    m_cpu = rh.make_clip_text_model("tiny", seed=3)
    tok = rh.FakeTokenizer(1000)
    ls = _patch_ds(caps)
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in layers]
    stats = ls.layer_stats_text_encoder_multi(m_gpu, None, names, stats_dir=tmp_path, sample_size=ss, precision="float32",
                                              progress=None, num_workers=0)
    covs = {l: orc.cov_from_state(stats[n].mom2.mom2.numpy(), stats[n].mom2.count) for l, n in zip(layers, names)}
    pipe = SimpleNamespace(text_encoder=m_gpu, tokenizer=tok, device=dev)
    emcid_main.COV_CACHE.clear()
    for e in range(3):
        reqs = [dict(r, source=f"edit{e} {r['source']}") for r in rh.make_requests(5)]
        cache = str(tmp_path / f"v{e}" / "c_")
        zs = rh.write_vstar_cache(cache, reqs, 64, seed=10 + e)
        hp = rh.make_hparams(layers, ss)
        emcid_main.apply_emcid_to_text_encoder(pipe, reqs, hp, device=dev, cache_name=cache, stats_dir=tmp_path, verbose=False)
        deltas = orc.execute_oracle(m_cpu, tok, reqs, layers, zs.numpy(), covs, 4000.0, 0.5)
        for l in layers:
            w = m_cpu.text_model.encoder.layers[l].mlp.fc2.weight
            with torch.no_grad():
                w[...] = torch.from_numpy(orc.apply_delta(w.numpy(), *deltas[l]))
    for l in layers:
        w0 = rh.make_clip_text_model("tiny", seed=3).text_model.encoder.layers[l].mlp.fc2.weight.numpy().astype(np.float64)
        got = m_gpu.text_model.encoder.layers[l].mlp.fc2.weight.detach().cpu().numpy().astype(np.float64) - w0
        want = m_cpu.text_model.encoder.layers[l].mlp.fc2.weight.numpy().astype(np.float64) - w0
        assert rel_fro(got, want) < DW_TOL
    assert len(emcid_main.FACTOR_CACHE) == (len(layers) if cached else 0)    # one factor per layer, reused by every edit
    emcid_main.clear_factor_cache()


def test_sdxl_two_encoder_edit_matches_oracle(dev, tmp_path):
    """apply_emcid_to_sdxl_text_encoders (emcid_main.py:38-106, :1085-1425): the stage-2 loop runs once per
    encoder with its own layers / lambda / statistics directory and `_2` v* files (BASELINE configs[3] shape:
    quick_gelu encoder 1, erf-gelu encoder 2)."""
    from emcid_b200 import emcid_main
    from types import SimpleNamespace
    ss = 48
    caps = rh.make_captions(64, 1000, seed=31)
    ls = _patch_ds(caps)
    tok = rh.FakeTokenizer(1000)
    enc = []
    for kind, seed, layers, lam, sub in (("tiny", 11, [0, 1], 4000.0, "text1"), ("tiny-gelu", 12, [1], 6000.0, "text2")):
        m_gpu = rh.make_clip_text_model(kind, seed=seed).to(dev)
        m_cpu = rh.make_clip_text_model(kind, seed=seed)
        names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in layers]
        sdir = tmp_path / sub
        stats = ls.layer_stats_text_encoder_multi(m_gpu, None, names, stats_dir=sdir, sample_size=ss, precision="float32",
                                                  progress=None, num_workers=0)
        covs = {l: orc.cov_from_state(stats[n].mom2.mom2.numpy(), stats[n].mom2.count) for l, n in zip(layers, names)}
        enc.append(SimpleNamespace(gpu=m_gpu, cpu=m_cpu, layers=layers, lam=lam, sdir=sdir, covs=covs))
    reqs = rh.make_requests(6)
    cache = str(tmp_path / "v" / "c_")
    z1 = rh.write_vstar_cache(cache, reqs, 64, seed=5)
    g2 = torch.Generator().manual_seed(6)
    z2 = []
    for r in reqs:   # encoder 2 reads the same names with a `_2` suffix (emcid_main.py:1157-1166)
        v = torch.randn(64, generator=g2)
        np.savez(cache + f"source_{r['source']}_dest_{r['dest']}_2.npz", v_star=v.numpy())
        z2.append(v)
    z2 = torch.stack(z2, dim=1)
    hp = rh.make_hparams(enc[0].layers, ss, mom2_update_weight=enc[0].lam)
    hp.layers_2, hp.mom2_update_weight_2 = enc[1].layers, enc[1].lam
    pipe = SimpleNamespace(text_encoder=enc[0].gpu, tokenizer=tok, text_encoder_2=enc[1].gpu, tokenizer_2=tok, device=dev)
    emcid_main.COV_CACHE.clear()
    out = emcid_main.apply_emcid_to_sdxl_text_encoders(pipe, reqs, hp, device=dev, cache_name=cache, stat_dir=enc[0].sdir,
                                                       stat_dir_2=enc[1].sdir, verbose=False, return_orig_text_encoder=True)
    assert out[0] is pipe and out[1] is not None and out[2] is not None
    for e, zs, orig in ((enc[0], z1, out[1]), (enc[1], z2, out[2])):
        deltas = orc.execute_oracle(e.cpu, tok, reqs, e.layers, zs.numpy(), e.covs, e.lam, 0.5)
        for l in e.layers:
            w0 = e.cpu.text_model.encoder.layers[l].mlp.fc2.weight.numpy().astype(np.float64)
            assert np.array_equal(orig.text_model.encoder.layers[l].mlp.fc2.weight.detach().cpu().numpy().astype(np.float64), w0)
            want = orc.apply_delta(w0.astype(np.float32), *deltas[l]).astype(np.float64) - w0
            got = e.gpu.text_model.encoder.layers[l].mlp.fc2.weight.detach().cpu().numpy().astype(np.float64) - w0
            assert rel_fro(got, want) < DW_TOL


def test_multi_token_edit_keeps_reference_semantics(dev, tmp_path):
    """num_edit_tokens > 1 (emcid_main.py:993-996): n * num_edit_tokens key columns, looked up at the last subject
    token, EOS and padding positions — served by the traced HF forward (padding rows do not exist in the packed
    library forward), solved by the library."""
    from emcid_b200 import compute_ks, emcid_main
    from types import SimpleNamespace
    model = rh.make_clip_text_model("tiny", seed=13).to(dev)
    tok = rh.FakeTokenizer(1000)
    reqs = rh.make_requests(4)
    name = "text_model.encoder.layers.1.mlp.fc2"
    k, z = compute_ks.get_module_input_output_at_words(model, tok, reqs, name, num_fact_token=3)
    assert compute_ks.LAST_PATH["native"] is False
    assert tuple(k.shape) == (4, 3, 256) and tuple(z.shape) == (4, 3, 64)
    k1, z1 = compute_ks.get_module_input_output_at_words(model, tok, reqs, name, num_fact_token=1)
    assert rel_fro(k[:, 0].cpu().numpy(), k1.cpu().numpy()) < 5e-6      # column 0 = the last subject token


def test_smoke_entry(dev):
    import __graft_entry__
    __graft_entry__.smoke()


@pytest.mark.parametrize("cached", [False, True])
def test_sequential_editing_matches_reference_fixture(dev, golden_dir, tmp_path, monkeypatch, cached):
    """BASELINE configs[4] against the unmodified reference itself: tests/golden/tiny_sequential.npz holds the fc2 weights
    after each of three successive reference edits (oracle/gen_golden.py::tiny_sequential); same statistics files, same
    v*, weights compared after every edit, with and without the cached factorisation of lambda * C."""
    from emcid_b200 import emcid_main
    from types import SimpleNamespace
    g = np.load(os.path.join(golden_dir, "tiny_sequential.npz"))
    layers, ss = [int(l) for l in g["layers"]], int(g["sample_size"])
    emcid_main.clear_factor_cache()
    if cached:
        monkeypatch.setattr(emcid_main, "FACTOR_CACHE_MAX_FRACTION", 1)     # tiny model: d = 256, n_pad = 128
    else:
        monkeypatch.setenv("EMCID_FACTOR_CACHE", "0")
    model = rh.make_clip_text_model("tiny", seed=3)
    for l in layers:
        assert np.array_equal(model.text_model.encoder.layers[l].mlp.fc2.weight.numpy(), g[f"w_before.{l}"])
    model = model.to(dev)
    tok = rh.FakeTokenizer(model.config.vocab_size)
    stats_dir = tmp_path / "stats"
    for l in layers:  # the reference's own statistics files -> identical C on both sides
        f = orc.stats_filename(str(stats_dir), "text_encoder", "ccs_filtered", f"text_model.encoder.layers.{l}.mlp.fc2",
                               "float32", ["mom2"], 3072, ss)
        os.makedirs(os.path.dirname(f), exist_ok=True)
        np.savez(f, **{"mom2.constructor": "util.runningstats.SecondMoment()", "mom2.count": int(g[f"count.{l}"]),
                       "mom2.mom2": g[f"mom2.{l}"], "sample_size": ss})
    pipe = SimpleNamespace(text_encoder=model, tokenizer=tok, device=dev)
    emcid_main.COV_CACHE.clear()
    for e in range(int(g["n_edits"])):
        reqs = [dict(r, source=f"edit{e} {r['source']}") for r in rh.make_requests(int(g["n_req"]))]
        cache = str(tmp_path / f"v{e}" / "c_")
        zs = rh.write_vstar_cache(cache, reqs, model.config.hidden_size, seed=10 + e)
        assert np.array_equal(zs.numpy(), g[f"zs.{e}"])
        hp = rh.make_hparams(layers, ss, mom2_update_weight=float(g["lam"]), edit_weight=float(g["edit_weight"]))
        emcid_main.apply_emcid_to_text_encoder(pipe, reqs, hp, device=dev, cache_name=cache, stats_dir=stats_dir, verbose=False)
        for l in layers:
            w = model.text_model.encoder.layers[l].mlp.fc2.weight.detach().cpu().numpy().astype(np.float64)
            assert rel_fro(w - g[f"w_before.{l}"], g[f"w_after.{e}.{l}"].astype(np.float64) - g[f"w_before.{l}"]) < DW_TOL, (e, l)
    assert len(emcid_main.FACTOR_CACHE) == (len(layers) if cached else 0)
    emcid_main.clear_factor_cache()
    emcid_main.COV_CACHE.clear()
