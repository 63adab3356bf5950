"""world_size-2 gloo run of the caption-sharded statistics pass (host logic only: the accumulator is a
test-only oracle stand-in, the product's CUDA accumulator cannot run here)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, rh
from oracle import emcid_oracle as orc


def _worker(rank, world, port, stats_dir, q, broadcast):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cpu_accumulator import OracleAccumulator
    from emcid_b200 import layer_stats

    torch.set_num_threads(2)
    model = rh.make_clip_text_model("tiny", seed=0)
    caps = rh.make_captions(90, 1000, seed=7)
    layer_stats.get_ccs_filtered_ds = lambda tokenizer: rh.SynthTokenDataset(caps)
    names = [f"text_model.encoder.layers.{l}.mlp.fc2" for l in (0, 1)]
    stats = layer_stats.layer_stats_text_encoder_multi(
        model, None, names, stats_dir=stats_dir, sample_size=64, precision="float32", progress=None,
        captions_per_batch=16, num_workers=0, _accumulator_factory=OracleAccumulator, broadcast=broadcast)
    q.put((rank, {n: (stats[n].mom2.count, None if stats[n].mom2.mom2 is None else stats[n].mom2.mom2.numpy())
                  for n in names}))
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("broadcast", [False, True])
def test_two_rank_sharded_pass_matches_single_process(tmp_path, broadcast):
    """Layer i is reduced onto rank i mod 2, which writes its npz; the counts reach every rank; with `broadcast` so do the
    matrices (what the single-layer reference-shaped call asks for)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + int(broadcast)) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q, broadcast)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    model = rh.make_clip_text_model("tiny", seed=0)
    caps = [c.numpy() for c in rh.make_captions(90, 1000, seed=7)]
    for l in (0, 1):
        name = f"text_model.encoder.layers.{l}.mlp.fc2"
        ref = orc.layer_stats_oracle(model, caps, l, 64)
        for r in range(2):
            count, m = results[r][name]
            assert count == ref.count                         # every rank learns the job-wide count
            if broadcast or r == l % 2:                       # the layer's root (and, with broadcast, everybody) holds the matrix
                assert np.linalg.norm(m - ref.mom2) / np.linalg.norm(ref.mom2) < 1e-6
            else:
                assert m is None
        f = orc.stats_filename(str(tmp_path), "text_encoder", "ccs_filtered", name, "float32", ["mom2"], 3072, 64)
        dat = np.load(f)  # written once, by the layer's root rank
        assert int(dat["mom2.count"]) == ref.count and int(dat["sample_size"]) == 64
