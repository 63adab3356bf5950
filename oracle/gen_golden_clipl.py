"""TEST INFRASTRUCTURE ONLY — CLIP-L-sized fixtures of the update, produced by the UNMODIFIED reference
(/root/reference, SilentView/EMCID) on CPU in the build container (several minutes on 8 cores):

    python oracle/gen_golden_clipl.py

  tests/golden/clipl_edit_digest.npz        execute_emcid_text_encoder + apply_emcid_to_text_encoder
                                            (emcid/emcid_main.py:769-815, :818-1082), 200 ICEB-style requests,
                                            shipped layer list [7, 8, 9, 10], lambda = 10000 (shipped JSON)
  tests/golden/clipl_sequential_digest.npz  BASELINE configs[4] (experiments/sequential_editing.py:98-171):
                                            10 successive 100-concept edits of layers 7-11, lambda = 4000

The [768 x 3072] fp64 updates of a CLIP-L edit are too large to commit, so the fixtures are digests: products of every
matrix with seeded probe vectors on both sides plus its Frobenius norm (a random projection preserves the relative
Frobenius error in expectation).  Inputs are rebuilt from seeds at test time: the random-init model (weight checksum
stored), the requests, the v* files, and covariances that are reproducible to the bit (oracle.exact_spd_matrix, written as
mom2 = C * 4096, count = 4096, so that mom2 / count == C exactly)."""
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from emcid_b200 import synth  # noqa: E402  (tokenizer stand-in: an INPUT of the reference run)
from oracle import emcid_oracle as orc  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle.gen_golden import checksum, write_stats_npz  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
D, H = 3072, 768
COUNT = 4096            # power of two: (C * COUNT) / COUNT == C in fp32
N_PROBE = 6


def probes(seed=0):
    rng = np.random.RandomState(seed)
    return rng.randn(D, N_PROBE), rng.randn(H, N_PROBE)


def digest(prefix, M, P, Q):
    """M [H x D] fp64."""
    M = np.asarray(M, dtype=np.float64)
    return {f"{prefix}.MP": M @ P, f"{prefix}.QtM": Q.T @ M, f"{prefix}.fro": np.linalg.norm(M)}


def write_covs(stats_dir, layers):
    for l in layers:
        C = orc.exact_spd_matrix(D, D + 1024, seed=l)
        write_stats_npz(stats_dir, l, (C * np.float32(COUNT)).astype(np.float32), COUNT, COUNT)


def fc2_weight(model, l):
    return model.text_model.encoder.layers[l].mlp.fc2.weight


def clipl_edit(out, n_req=200, layers=(7, 8, 9, 10), lam=10000, edit_weight=0.5):
    model = rh.make_clip_text_model("clip-l", seed=0)
    tok = synth.WordHashTokenizer(model.config.vocab_size)
    reqs = rh.make_requests(n_req)
    P, Q = probes()
    tmp = tempfile.mkdtemp()
    try:
        stats_dir = os.path.join(tmp, "stats")
        write_covs(stats_dir, layers)
        cache = os.path.join(tmp, "v", "c_")
        rh.write_vstar_cache(cache, reqs, H, seed=2)
        hp = rh.make_hparams(layers, COUNT, mom2_update_weight=lam, edit_weight=edit_weight)
        d = dict(n_req=n_req, layers=np.array(layers), lam=lam, edit_weight=edit_weight, count=COUNT, seed_vstar=2,
                 weight_checksum=checksum(model), n_probe=N_PROBE)
        w0 = {l: fc2_weight(model, l).detach().clone() for l in layers}
        t = time.time()
        deltas = rh.run_reference_execute(model, tok, reqs, hp, cache, stats_dir)
        print("reference execute:", round(time.time() - t, 1), "s")
        R = np.random.RandomState(1).randn(n_req, N_PROBE)
        for l in layers:
            assert torch.equal(w0[l], fc2_weight(model, l))
            adj, resid = (x.numpy() for x in deltas[f"text_model.encoder.layers.{l}.mlp.fc2.weight"])
            d.update(digest(f"upd.{l}", resid @ adj.T, P, Q))
            d[f"adj.{l}.PtA"] = P.T @ adj               # [N_PROBE x n]
            d[f"adj.{l}.AR"] = adj @ R                  # [D x N_PROBE]
            d[f"adj.{l}.fro"] = np.linalg.norm(adj)
            d[f"resid.{l}.RR"] = resid @ R              # [H x N_PROBE]
            d[f"resid.{l}.fro"] = np.linalg.norm(resid)
        t = time.time()
        rh.run_reference_apply(model, tok, reqs, hp, cache, stats_dir)
        print("reference apply:", round(time.time() - t, 1), "s")
        for l in layers:
            d.update(digest(f"applied.{l}", fc2_weight(model, l).detach().double().numpy() - w0[l].double().numpy(), P, Q))
        np.savez_compressed(os.path.join(GOLD, out), **d)
        print(out, {l: float(d[f"upd.{l}.fro"]) for l in layers})
    finally:
        shutil.rmtree(tmp)


def clipl_sequential(out, n_edits=10, n_req=100, layers=(7, 8, 9, 10, 11), lam=4000, edit_weight=0.5, keep=(0, 4, 9)):
    model = rh.make_clip_text_model("clip-l", seed=0)
    tok = synth.WordHashTokenizer(model.config.vocab_size)
    P, Q = probes()
    tmp = tempfile.mkdtemp()
    try:
        stats_dir = os.path.join(tmp, "stats")
        write_covs(stats_dir, layers)
        d = dict(n_req=n_req, n_edits=n_edits, layers=np.array(layers), lam=lam, edit_weight=edit_weight, count=COUNT,
                 seed_vstar0=40, weight_checksum=checksum(model), n_probe=N_PROBE, keep=np.array(keep))
        w0 = {l: fc2_weight(model, l).detach().double().numpy().copy() for l in layers}
        for e in range(n_edits):
            reqs = [dict(r, source=f"edit{e} {r['source']}") for r in rh.make_requests(n_req)]
            cache = os.path.join(tmp, f"v{e}", "c_")
            rh.write_vstar_cache(cache, reqs, H, seed=40 + e)
            hp = rh.make_hparams(layers, COUNT, mom2_update_weight=lam, edit_weight=edit_weight)
            t = time.time()
            rh.run_reference_apply(model, tok, reqs, hp, cache, stats_dir)
            print(f"reference edit {e}:", round(time.time() - t, 1), "s")
            if e in keep:
                for l in layers:
                    d.update(digest(f"cum.{e}.{l}", fc2_weight(model, l).detach().double().numpy() - w0[l], P, Q))
        np.savez_compressed(os.path.join(GOLD, out), **d)
        print(out, {l: float(d[f"cum.{n_edits - 1}.{l}.fro"]) for l in layers})
    finally:
        shutil.rmtree(tmp)


if __name__ == "__main__":
    assert rh.reference_available(), "run in the build container (needs /root/reference)"
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["edit", "sequential"]
    if "edit" in which:
        clipl_edit("clipl_edit_digest.npz")
    if "sequential" in which:
        clipl_sequential("clipl_sequential_digest.npz")
